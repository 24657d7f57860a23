// np2_host.h — host phases of the polish path that stay on the CPU by design (small, sequential, irregular):
// record parsing/filtering, the LQ-region state machine over sparse events, genotype rules, the phasing graph and
// Louvain clustering, consensus splicing.  Everything works on flat arrays produced/consumed by the kernels.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/np2gpu.h"

namespace np2 {

/* ---------------------------------------------------------------- ingest */
struct Ingest {
    // candidate reads = records that pass the record-level filter (main.rs:1758-1771)
    std::vector<int32_t> all_tid, all_pos;  // every record, for the sortedness assertion (main.rs:1753-1756)
    std::vector<int32_t> rec_idx;
    std::vector<uint32_t> pos, ncols, rlen;
    std::vector<uint8_t> is_clip;
    std::vector<uint64_t> seq_off;      // byte offset of SEQ in the blob
    std::vector<uint32_t> op_off;       // n + 1
    std::vector<uint32_t> op_col, op_q, op_t, op_cig;  // column-consuming ops only (M,=,X,I,D)
    std::vector<uint64_t> nib_off;      // n + 1, bytes, 16-B aligned
    std::vector<uint32_t> ck_off;       // n + 1, 32-column blocks
    uint64_t total_cols = 0;
};
// throws np2::Error(NP2_ERR_FORMAT) where the reference panics
void parse_records(const uint8_t *bam, uint64_t bam_len, uint32_t tlen, const np2_opts &opt, Ingest &out);

/* ---------------------------------------------------------------- regions */
struct Cns {  // consensus as parallel arrays (ascending)
    std::vector<uint32_t> pos;
    std::vector<uint8_t> base;
};
struct Regions {
    std::vector<uint32_t> start, end;  // reference order: descending position
};
// LQ state machine of generate_cns_from_best_score_lq (main.rs:1586-1625) driven by the sparse list of
// consensus indices whose flags != 0 (ascending); flags: bit0 qv<95, bit1 cov<2.
void find_regions(const uint32_t *cpos, const uint8_t *cbase, const uint8_t *cflags, uint64_t n, const uint32_t *events,
                  uint64_t n_events, Regions &out);

/* ---------------------------------------------------------------- candidates */
struct CandSet {
    // per region: candidates in read order (ref first), capped at 60 (main.rs:1474)
    std::vector<uint32_t> roff;  // n_regions + 1
    std::vector<uint32_t> order;
    std::vector<uint16_t> kscore;
    std::vector<uint64_t> kmer;
    std::vector<uint64_t> seq_off;  // into pool (per candidate: seq_off[i], seq_len[i])
    std::vector<uint32_t> seq_len;
    const uint8_t *pool = nullptr;
};

/* ---------------------------------------------------------------- genotype / phasing / patching */
struct RegionState {
    uint8_t lable = 0;
    std::string sudoseed;
    std::vector<uint32_t> cand;  // indices into CandSet arrays, current order
};
// mark_hete_lqseqs (main.rs:916-946): sets HETE lables and zeroes minor kscores (in cs.kscore)
void mark_hete(CandSet &cs, std::vector<RegionState> &rs);
// phase_reads_by_lqseqs (main.rs:948-1015) + louvain.rs: returns the alignseq indices to blank (sorted, unique)
std::vector<uint32_t> phase_reads(const CandSet &cs, const std::vector<RegionState> &rs, bool asref, bool use_all_reads);
// fill_seed_lqseqs (main.rs:862-914)
void fill_seed(const CandSet &cs, std::vector<RegionState> &rs, long max_indel_len);
// update_consensus_with_lqseqs (main.rs:1027-1058)
void splice(const Regions &rg, const std::vector<RegionState> &rs, uint8_t lable, const Cns &in, Cns &out);

// reupdate_consensus_with_lqseqs (main.rs:1060-1420), split around the device scoring call:
struct Reupdate {
    struct Group {
        uint32_t sj, ej;           // range in rech
        uint64_t first_string;     // index of the group's first string
    };
    std::vector<uint32_t> rech;    // RECH region indices, ascending position
    std::vector<Group> groups;
    std::vector<uint8_t> pool;     // all strings to score, concatenated
    std::vector<uint64_t> off;     // n_strings + 1
};
void reupdate_build(const Regions &rg, const CandSet &cs, const std::vector<RegionState> &rs, const Cns &cns, uint32_t k,
                    Reupdate &ru);
void reupdate_apply(const Regions &rg, CandSet &cs, std::vector<RegionState> &rs, const Reupdate &ru,
                    const uint16_t *kscores, uint32_t iter_count, const Cns &in, Cns &out);

const uint8_t LABLE_TEMP = 0x01, LABLE_SUCC = 0x80, LABLE_HETE = 0x40, LABLE_RECH = 0x20;  // main.rs:655-658

}  // namespace np2
