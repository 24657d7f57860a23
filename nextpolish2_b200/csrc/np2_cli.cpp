// np2_cli.cpp — `nextPolish2` command line on top of libnp2gpu (C ABI in include/np2gpu.h).
//
// Same surface as the reference CLI (src/utils/option.rs:45-225): positional <sorted.bam> <genome.fa[.gz]>
// <k1.yak> [k2.yak ...], options -o -u --out_pos -k -t -i -m -l -L -n -s -S -a -q -c -r --min_base_cov, same
// defaults (option.rs:267-292).  Extensions: -g/--gpus N (GPUs to use, default all visible); the sub-command
// `nextPolish2 count` (yak count on the GPU: FASTA/FASTQ[.gz] in, .yak dump out; yak/main.c:24-83).
//
// This file is the caller side of the hot path: hand-written BGZF/BAM/BAI and FASTA(.gz) readers
// (the BAM is memory-mapped; a contig's BGZF members are inflated on the device by np2_bgzf_inflate, one
// warp per member, or with --host-inflate by zlib on the host threads, in parallel and in place) that hand each contig's
// raw alignment records to np2_polish_contig, and the orchestration
// the reference does with three thread stages (main.rs:1698-1853): contigs are LPT-partitioned over the GPUs, three to
// six host threads (contexts) per GPU share one set of tables, records are printed in INPUT order (= the reference
// with -t 1).
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <future>
#include <thread>
#include <vector>

#include "../../include/np2gpu.h"

namespace {

std::string g_partial_out;  // -o file being written: removed when the run dies, so that no truncated FASTA is left behind
[[noreturn]] void die(const std::string &m) {
    fprintf(stderr, "%s\n", m.c_str());
    if (!g_partial_out.empty()) remove(g_partial_out.c_str());
    fflush(stdout);
    fflush(stderr);
    _exit(101);  // the reference aborts with a panic message; no atexit handlers: other threads may be inside CUDA calls
}

/* ---------------------------------------------------------------- FASTA (kseq semantics) */
struct Contig {
    std::string name;  // first word of the header (kseq head())
    std::string seq;
};
std::vector<Contig> read_fasta(const std::string &path) {
    gzFile f = gzopen(path.c_str(), "rb");
    if (!f) die("\"" + path + "\" does not exist!");
    gzbuffer(f, 1 << 20);
    std::vector<Contig> out;
    std::vector<char> buf(4 << 20);
    std::string header;
    bool in_header = false, bol = true;
    auto end_header = [&]() {
        size_t e = 0;
        while (e < header.size() && !isspace((unsigned char)header[e])) e++;
        out.push_back(Contig{header.substr(0, e), std::string()});
        header.clear();
    };
    for (;;) {
        const int n = gzread(f, buf.data(), (unsigned)buf.size());
        if (n <= 0) break;
        int i = 0;
        while (i < n) {
            if (in_header) {  // up to the end of the line
                const char *e = (const char *)memchr(buf.data() + i, '\n', n - i);
                const int stop = e ? (int)(e - buf.data()) : n;
                header.append(buf.data() + i, stop - i);
                i = stop;
                if (e) {
                    while (!header.empty() && header.back() == '\r') header.pop_back();
                    end_header();
                    in_header = false;
                    bol = true;
                    i++;
                }
            } else if (bol && buf[i] == '>') {
                in_header = true;
                bol = false;
                i++;
            } else {  // a stretch of sequence: everything up to the next line break goes in with one append
                const char *e = (const char *)memchr(buf.data() + i, '\n', n - i);
                int stop = e ? (int)(e - buf.data()) : n;
                int len = stop - i;
                while (len > 0 && buf[i + len - 1] == '\r') len--;
                if (len > 0) {
                    if (out.empty()) die("FASTA parsing failed!");
                    out.back().seq.append(buf.data() + i, len);
                }
                i = stop;
                bol = false;
                if (e) {
                    bol = true;
                    i++;
                }
            }
        }
    }
    if (in_header) end_header();
    gzclose(f);
    return out;
}

/* ---------------------------------------------------------------- `count`: yak count on the GPU (yak/main.c:24-83) */
// FASTA / FASTQ (optionally gzipped) streamed in batches of bases: cb(bases, offsets) per batch
template <class F>
void stream_sequences(const std::string &path, size_t batch_bases, F cb) {
    gzFile f = gzopen(path.c_str(), "rb");
    if (!f) die("\"" + path + "\" does not exist!");
    gzbuffer(f, 1 << 20);
    std::vector<uint8_t> seq;
    std::vector<uint64_t> off(1, 0);
    std::string line;
    auto getline = [&]() -> bool {  // without the line terminator
        line.clear();
        for (;;) {
            char tmp[1 << 16];
            if (!gzgets(f, tmp, sizeof tmp)) return !line.empty();
            size_t l = strlen(tmp);
            const bool eol = l && tmp[l - 1] == '\n';
            while (l && (tmp[l - 1] == '\n' || tmp[l - 1] == '\r')) l--;
            line.append(tmp, l);
            if (eol) return true;
        }
    };
    auto flush = [&](bool force) {
        if (off.size() > 1 && (force || seq.size() >= batch_bases)) {
            cb(seq, off);
            seq.clear();
            off.assign(1, 0);
        }
    };
    bool have = getline();
    while (have) {
        if (line.empty()) {
            have = getline();
            continue;
        }
        if (line[0] == '>') {  // FASTA: sequence lines until the next header
            while ((have = getline()) && (line.empty() || (line[0] != '>' && line[0] != '@')))
                seq.insert(seq.end(), line.begin(), line.end());
            off.push_back(seq.size());
        } else if (line[0] == '@') {  // FASTQ: sequence lines until '+', then as many quality characters
            size_t n = 0;
            while ((have = getline()) && !(line.size() && line[0] == '+')) {
                seq.insert(seq.end(), line.begin(), line.end());
                n += line.size();
            }
            off.push_back(seq.size());
            size_t q = 0;
            while (q < n && (have = getline())) q += line.size();
            have = getline();
        } else die("sequence parsing failed: " + path);
        flush(false);
    }
    flush(true);
    gzclose(f);
}
int main_count(int argc, char **argv) {
    int k = 31, bloom = 0, pre = 10, gpu = 0;
    std::string out;
    std::vector<std::string> in;
    for (int i = 2; i < argc; i++) {
        const std::string a = argv[i];
        auto val = [&]() -> std::string {
            if (i + 1 >= argc) die("option " + a + " needs a value");
            return argv[++i];
        };
        if (a == "-k") k = atoi(val().c_str());
        else if (a == "-b") bloom = atoi(val().c_str());
        else if (a == "-p") pre = atoi(val().c_str());
        else if (a == "-o") out = val();
        else if (a == "-g") gpu = atoi(val().c_str());
        else if (a == "-t" || a == "-K" || a == "-H") val();  // accepted for compatibility with `yak count`
        else in.push_back(a);
    }
    if (in.empty() || out.empty()) {
        fprintf(stderr,
                "Usage: nextPolish2 count [options] -o <out.yak> <in.fa|fq[.gz]> [in.fa]\n"
                "  -k INT   k-mer size [31]\n"
                "  -b INT   as in `yak count`: when > 0 only k-mers seen at least twice are kept (what yak's Bloom-filter\n"
                "           pass + second pass + shrink leave, without the Bloom filter's false positives); a second\n"
                "           input must be the same file as the first\n"
                "  -o FILE  dump the counts in yak's format\n"
                "  -g INT   GPU to use [0]\n");
        return 1;
    }
    // yak with -b and two DIFFERENT files keeps the k-mers of file 2 that passed the Bloom pass over file 1 (main.c:65-71):
    // that is not what "count the second file, keep counts >= 2" computes, so it is refused rather than mis-answered.
    // The usual invocation (the same reads twice, as test/hh.sh does) is the supported one.  yak's Bloom false positives
    // (singletons that survive) are not reproduced either way: this dump holds exactly the k-mers seen twice or more.
    if (bloom > 0 && in.size() >= 2 && in[0] != in[1])
        die("ERROR: count -b with two different input files is not supported (give the same file twice, as yak's own "
            "two-pass recipe does)");
    if (pre != 10) die("ERROR: -p must be 10 (NextPolish2 compares hash >> 10, kmer.rs:52-54)");
    if (k >= 64) die("ERROR: -k must be smaller than 64");
    np2_ctx *ctx = nullptr;
    if (np2_ctx_create(gpu, &ctx) != NP2_OK) die(np2_last_error());
    np2_counter *c = nullptr;
    if (np2_count_create(ctx, (uint32_t)k, &c) != NP2_OK) die(np2_last_error());
    // yak main.c:65-71: without -b the first file is counted; with -b the table is rebuilt from the second one
    const std::string &src = (bloom > 0 && in.size() >= 2) ? in[1] : in[0];
    uint64_t n_seq = 0;
    stream_sequences(src, 256u << 20, [&](const std::vector<uint8_t> &seq, const std::vector<uint64_t> &off) {
        if (np2_count_add(c, seq.data(), off.data(), off.size() - 1) != NP2_OK) die(np2_last_error());
        n_seq += off.size() - 1;
    });
    uint64_t n_kmers = 0;
    const uint64_t distinct = np2_count_distinct(c, &n_kmers);
    if (np2_count_finish(c, bloom > 0 ? 2 : 1, out.c_str(), nullptr) != NP2_OK) die(np2_last_error());
    fprintf(stderr, "[M::count] %llu sequences, %llu k-mers, %llu distinct; dumped to '%s'\n", (unsigned long long)n_seq,
            (unsigned long long)n_kmers, (unsigned long long)distinct, out.c_str());
    np2_count_destroy(c);
    np2_ctx_destroy(ctx);
    return 0;
}

/* ---------------------------------------------------------------- BGZF / BAM / BAI */
// The file is memory-mapped.  A contig's records are the bytes between two virtual offsets the index gives (first
// chunk begin, last chunk end of the reference): the BGZF members in between are located by a serial walk over their
// headers (a few thousand per contig) and inflated IN PARALLEL straight into their final place in the record buffer,
// so no byte is copied twice and no lock is held while a contig is decoded.
struct BamFile {
    const uint8_t *map = nullptr;
    size_t size = 0;
    std::vector<std::string> ref_names;
    std::vector<uint32_t> ref_lens;
    std::vector<uint64_t> ref_voff;  // virtual offset of the first record of each reference (UINT64_MAX = none / unknown)
    std::vector<uint64_t> ref_vend;  // virtual offset just behind its last record
    uint64_t first_rec_voff = 0;
    bool have_index = false;
    int threads = 1;
};
struct Member {   // one BGZF member (gzip member with the BC extra field)
    uint64_t coff;   // file offset of the member
    uint64_t data;   // file offset of the deflate payload
    uint32_t clen;   // payload bytes
    uint32_t isize;  // inflated bytes
    uint32_t total;  // member bytes
};
// parses the member header at file offset o; false at EOF
bool member_at(const BamFile &bf, uint64_t o, Member &m) {
    if (o >= bf.size) return false;
    const uint8_t *h = bf.map + o;
    if (o + 18 > bf.size || h[0] != 31 || h[1] != 139 || h[2] != 8 || !(h[3] & 4)) die("BAM/SAM parsing failed!");
    const uint32_t xlen = h[10] | h[11] << 8;
    if (o + 12 + xlen > bf.size) die("BAM/SAM parsing failed!");
    uint32_t bsize = 0;
    for (uint32_t x = 0; x + 4 <= xlen;) {
        const uint8_t *e = h + 12 + x;
        const uint32_t slen = e[2] | e[3] << 8;
        if (x + 4 + slen > xlen) die("BAM/SAM parsing failed!");  // the subfield runs past the extra field
        if (e[0] == 'B' && e[1] == 'C' && slen == 2) bsize = (e[4] | e[5] << 8) + 1;
        x += 4 + slen;
    }
    if (!bsize || bsize < 12 + xlen + 8 || o + bsize > bf.size) die("BAM/SAM parsing failed!");
    m.coff = o;
    m.data = o + 12 + xlen;
    m.clen = bsize - 12 - xlen - 8;
    m.total = bsize;
    memcpy(&m.isize, h + bsize - 4, 4);
    if (m.isize > 65536) die("BAM/SAM parsing failed!");  // a BGZF member inflates to at most 64 KiB
    return true;
}
void inflate_member(const BamFile &bf, const Member &m, uint8_t *out) {
    if (!m.isize) return;
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (inflateInit2(&zs, -15) != Z_OK) die("BAM/SAM parsing failed!");
    zs.next_in = (Bytef *)(bf.map + m.data);
    zs.avail_in = m.clen;
    zs.next_out = out;
    zs.avail_out = m.isize;
    const int rc = inflate(&zs, Z_FINISH);
    inflateEnd(&zs);
    if (rc != Z_STREAM_END || zs.total_out != m.isize) die("BAM/SAM parsing failed!");
}
// growable byte buffer that is never zero-filled (a std::vector::resize would touch every page first).  The polish
// lanes use page-locked memory (np2_host_alloc): the device then pulls the SEQ fields straight out of the inflated
// records and the library's host-side compaction pass is skipped.
struct Blob {
    uint8_t *p = nullptr;
    size_t n = 0, cap = 0;
    bool pinned = false;
    explicit Blob(bool page_locked = false) : pinned(page_locked) {}
    Blob(const Blob &) = delete;
    Blob &operator=(const Blob &) = delete;
    ~Blob() { release(); }
    void release() {
        if (p) {
            if (pinned) np2_host_free(p);
            else free(p);
        }
        p = nullptr;
        cap = 0;
    }
    void resize(size_t want) {
        if (want > cap) {
            release();
            cap = want + want / 4 + 4096;
            if (pinned) {
                void *q = nullptr;
                if (np2_host_alloc(cap, &q) != NP2_OK) {  // fall back to pageable memory
                    pinned = false;
                    q = malloc(cap);
                }
                p = static_cast<uint8_t *>(q);
            } else {
                p = static_cast<uint8_t *>(malloc(cap));
            }
            if (!p) die("out of memory");
        }
        n = want;
    }
    void swap(Blob &o) {
        std::swap(p, o.p);
        std::swap(n, o.n);
        std::swap(cap, o.cap);
        std::swap(pinned, o.pinned);
    }
    const uint8_t *data() const { return p; }
    size_t size() const { return n; }
};

void open_bam(const std::string &path, BamFile &bf) {
    const int fd = open(path.c_str(), O_RDONLY);
    if (fd < 0) die("\"" + path + "\" does not exist!");
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size <= 0) die("BAM/SAM parsing failed!");
    bf.size = (size_t)st.st_size;
    void *mp = mmap(nullptr, bf.size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (mp == MAP_FAILED) die("BAM/SAM parsing failed!");
    bf.map = static_cast<const uint8_t *>(mp);
    // header: sequential members until the reference dictionary is complete
    std::vector<uint8_t> u;
    std::vector<uint64_t> blk_coff, blk_uoff;
    uint64_t next = 0;
    auto need = [&](size_t n) {
        while (u.size() < n) {
            Member m;
            if (!member_at(bf, next, m)) die("BAM/SAM parsing failed!");
            blk_coff.push_back(m.coff);
            blk_uoff.push_back(u.size());
            const size_t o = u.size();
            u.resize(o + m.isize);
            inflate_member(bf, m, u.data() + o);
            next = m.coff + m.total;
        }
    };
    need(12);
    if (memcmp(u.data(), "BAM\1", 4) != 0) die("BAM/SAM parsing failed!");
    int32_t l_text, n_ref;
    memcpy(&l_text, u.data() + 4, 4);
    if (l_text < 0) die("BAM/SAM parsing failed!");
    need(12 + (size_t)l_text);
    memcpy(&n_ref, u.data() + 8 + l_text, 4);
    if (n_ref < 0) die("BAM/SAM parsing failed!");
    size_t o = 12 + (size_t)l_text;
    for (int32_t i = 0; i < n_ref; i++) {
        need(o + 4);
        int32_t l_name;
        memcpy(&l_name, u.data() + o, 4);
        if (l_name <= 0 || l_name > (1 << 20)) die("BAM/SAM parsing failed!");
        need(o + 4 + l_name + 4);
        bf.ref_names.emplace_back((const char *)u.data() + o + 4, strnlen((const char *)u.data() + o + 4, (size_t)l_name));
        uint32_t l_ref;
        memcpy(&l_ref, u.data() + o + 4 + l_name, 4);
        bf.ref_lens.push_back(l_ref);
        o += 8 + l_name;
    }
    // virtual offset of the first alignment record
    size_t bi = blk_uoff.size() - 1;
    while (blk_uoff[bi] > o) bi--;
    if (o == u.size()) bf.first_rec_voff = next << 16;
    else bf.first_rec_voff = blk_coff[bi] << 16 | (o - blk_uoff[bi]);
    bf.ref_voff.assign(n_ref, UINT64_MAX);
    bf.ref_vend.assign(n_ref, 0);
    // BAI: smallest chunk begin and largest chunk end of each reference
    for (const std::string &ip : {path + ".bai", path.substr(0, path.size() > 4 ? path.size() - 4 : 0) + ".bai"}) {
        FILE *fi = fopen(ip.c_str(), "rb");
        if (!fi) continue;
        char magic[4];
        int32_t nr;
        if (fread(magic, 1, 4, fi) != 4 || memcmp(magic, "BAI\1", 4) != 0 || fread(&nr, 4, 1, fi) != 1) die("bad BAI");
        for (int32_t r = 0; r < nr && r < n_ref; r++) {
            int32_t n_bin;
            if (fread(&n_bin, 4, 1, fi) != 1) die("bad BAI");
            uint64_t best = UINT64_MAX, last = 0;
            for (int32_t b = 0; b < n_bin; b++) {
                uint32_t bin;
                int32_t n_chunk;
                if (fread(&bin, 4, 1, fi) != 1 || fread(&n_chunk, 4, 1, fi) != 1) die("bad BAI");
                for (int32_t c = 0; c < n_chunk; c++) {
                    uint64_t be[2];
                    if (fread(be, 8, 2, fi) != 2) die("bad BAI");
                    if (bin != 37450) {
                        best = std::min(best, be[0]);
                        last = std::max(last, be[1]);
                    }
                }
            }
            int32_t n_intv;
            if (fread(&n_intv, 4, 1, fi) != 1) die("bad BAI");
            fseeko(fi, (off_t)n_intv * 8, SEEK_CUR);
            bf.ref_voff[r] = best;
            bf.ref_vend[r] = last;
        }
        fclose(fi);
        bf.have_index = true;
        break;
    }
    if (!bf.have_index) die("Faield random access BAM/SAM!");  // IndexedReader needs the index (main.rs:1745-1747)
}

// IndexedReader::fetch((tid, 0, len)) + read loop (main.rs:1745-1751): every record of that reference, file order.
// Thread-safe (the map is read-only).  With a context the members are inflated ON THE DEVICE (np2_bgzf_inflate: the
// compressed span goes up, one warp inflates each member, the records come back into the blob — page-locked in the polish
// lanes, so both copies are DMA transfers); without one, `threads` host workers call zlib (--host-inflate, and the
// passes that run before any context exists).
// the BGZF members that hold the records of reference `tid`, and which bytes of their inflated concatenation the records
// are: [u0, u0 + n).  false = the reference has no records.
bool locate_members(const BamFile &bf, int tid, std::vector<Member> &ms, uint64_t &u0_out, uint64_t &n_out) {
    ms.clear();
    u0_out = n_out = 0;
    if (tid < 0 || bf.ref_voff[tid] == UINT64_MAX) return false;
    const uint64_t v0 = bf.ref_voff[tid], v1 = bf.ref_vend[tid];
    if (v1 <= v0) return false;
    const uint64_t c0 = v0 >> 16, c1 = v1 >> 16;
    const uint32_t u0 = (uint32_t)(v0 & 0xFFFF), u1 = (uint32_t)(v1 & 0xFFFF);
    // members [c0, c1]; the one at c1 only contributes its first u1 bytes (none when u1 == 0)
    uint64_t total = 0;
    for (uint64_t o = c0; o <= c1;) {
        Member m;
        if (!member_at(bf, o, m)) {
            if (o == c1 && u1 == 0) break;  // the end offset may point at EOF
            die("BAM/SAM parsing failed!");
        }
        if (o == c1 && u1 == 0) break;
        ms.push_back(m);
        total += m.isize;
        o += m.total;
    }
    if (ms.empty()) return false;
    const uint64_t cut_tail = (ms.back().coff == c1) ? ms.back().isize - std::min<uint32_t>(u1, ms.back().isize) : 0;
    if ((uint64_t)u0 + cut_tail > total) die("BAM/SAM parsing failed!");
    u0_out = u0;
    n_out = total - u0 - cut_tail;
    return true;
}
void member_table(const std::vector<Member> &ms, std::vector<uint64_t> &off, std::vector<uint32_t> &clen, std::vector<uint32_t> &isz) {
    off.resize(ms.size());
    clen.resize(ms.size());
    isz.resize(ms.size());
    for (size_t i = 0; i < ms.size(); i++) {
        off[i] = ms[i].data;
        clen[i] = ms[i].clen;
        isz[i] = ms[i].isize;
    }
}
void fetch_records(const BamFile &bf, int tid, Blob &blob, int threads, np2_ctx *gpu = nullptr) {
    blob.resize(0);
    std::vector<Member> ms;
    uint64_t u0 = 0, n = 0;
    if (!locate_members(bf, tid, ms, u0, n)) return;
    std::vector<uint64_t> uoff(1, 0);
    for (const Member &m : ms) uoff.push_back(uoff.back() + m.isize);
    const uint64_t total = uoff.back(), cut_tail = total - u0 - n;
    blob.resize(n);
    if (gpu) {
        std::vector<uint64_t> off;
        std::vector<uint32_t> clen, isz;
        member_table(ms, off, clen, isz);
        if (np2_bgzf_inflate(gpu, bf.map, bf.size, off.data(), clen.data(), isz.data(), (uint32_t)ms.size(), u0, n, blob.p,
                             nullptr) != NP2_OK)
            die(np2_last_error());
    }
    std::atomic<size_t> next{gpu ? ms.size() : 0};
    auto work = [&]() {
        std::vector<uint8_t> tmp;
        for (size_t i; (i = next.fetch_add(1)) < ms.size();) {
            const uint64_t b = uoff[i], e = uoff[i + 1];          // position of the member in the untrimmed stream
            const uint64_t lo = std::max<uint64_t>(b, u0), hi = std::min<uint64_t>(e, total - cut_tail);
            if (lo >= hi) continue;
            if (lo == b && hi == e) {
                inflate_member(bf, ms[i], blob.p + (b - u0));       // whole member: straight into place
            } else {                                               // first / last member: through a scratch buffer
                tmp.resize(ms[i].isize);
                inflate_member(bf, ms[i], tmp.data());
                memcpy(blob.p + (lo - u0), tmp.data() + (lo - b), hi - lo);
            }
        }
    };
    const int T = gpu ? 1 : std::max(1, std::min<int>(threads, (int)ms.size()));
    std::vector<std::thread> th;
    for (int t = 1; t < T; t++) th.emplace_back(work);
    work();
    for (auto &t : th) t.join();
    // the index is trusted for the range, the records are not: every one must belong to this reference
    for (uint64_t p = 0; p < n;) {
        int32_t bs, ref_id;
        if (p + 8 > n) die("BAM/SAM parsing failed!");
        memcpy(&bs, blob.p + p, 4);
        memcpy(&ref_id, blob.p + p + 4, 4);
        if (bs < 32 || p + 4 + (uint64_t)bs > n || ref_id != tid) die("BAM/SAM parsing failed!");
        p += 4 + (uint64_t)bs;
    }
}

/* ---------------------------------------------------------------- options (option.rs:45-292) */
struct Cli {
    std::string bam, fa, out = "stdout";
    std::vector<std::string> yaks;
    np2_opts o;
    int threads = 0, gpus = 0;  // threads 0 = not given: BGZF inflate and record parsing use up to 16 hardware threads
    bool host_inflate = false;  // --host-inflate: zlib on the host threads instead of the device inflate kernel
};
void usage() {
    fprintf(stderr,
            "Usage: nextPolish2 [OPTIONS] <HiFi.map.bam> <genome.fa[.gz]> <short.read.yak>...\n"
            "  -o, --out <FILE>            output file [default: stdout]\n"
            "  -u, --uppercase             output in uppercase sequences\n"
            "      --out_pos               output each base and its position\n"
            "  -k, --min_kmer_count <INT>  filter kmers in k-mer dataset with count <= INT [default: 5]\n"
            "  -t, --thread <INT>          host threads for BGZF inflate and record parsing [default: min(16, hardware threads)]\n"
            "  -i, --iter_count <INT>      number of iterations to attempt phasing [default: 2]\n"
            "  -m, --model <ref|len>       phasing model [default: ref]\n"
            "  -l, --min_read_len <INT>    filter reads with length <= INT [default: 1000]\n"
            "  -L, --min_ctg_len <INT>     don't correct reference sequences with length <= INT [default: 1000000]\n"
            "  -n, --max_indel_len <INT>   ignore indel errors with length > INT [default: 20]\n"
            "  -s, --use_supplementary     use supplementary alignments\n"
            "  -S, --use_secondary         use secondary alignments\n"
            "  -a, --min_map_len <FLOAT>   filter alignments with alignment length <= min(INT, FLOAT * read_length) [default: 500.5]\n"
            "  -q, --min_map_qual <INT>    filter alignments with mapping quality <= INT [default: 1]\n"
            "  -c, --max_clip_len <INT>    filter alignments with unaligned length >= INT [default: 100]\n"
            "  -r, --use_all_reads         use all unfiltered reads\n"
            "      --min_base_cov <INT>    accepted for compatibility (unused by the reference as well)\n"
            "  -g, --gpus <INT>            GPUs to use [default: all visible]\n"
            "      --host-inflate          inflate the BGZF members with zlib on the host threads instead of on the GPU\n");
}
Cli parse_args(int argc, char **argv) {
    Cli c;
    np2_opts_default(&c.o);
    std::vector<std::string> pos;
    auto val = [&](int &i) -> std::string {
        if (i + 1 >= argc) {
            usage();
            exit(2);
        }
        return argv[++i];
    };
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        if (a == "-h" || a == "--help") {
            usage();
            exit(0);
        } else if (a == "-o" || a == "--out") c.out = val(i);
        else if (a == "-u" || a == "--uppercase") c.o.uppercase = 1;
        else if (a == "--out_pos") c.o.out_pos = 1;
        else if (a == "-k" || a == "--min_kmer_count") c.o.min_kmer_count = (uint32_t)atoi(val(i).c_str());
        else if (a == "-t" || a == "--thread") c.threads = std::max(1, atoi(val(i).c_str()));
        else if (a == "-i" || a == "--iter_count") c.o.iter_count = (uint32_t)atoi(val(i).c_str());
        else if (a == "-m" || a == "--model") {
            std::string m = val(i);
            if (m != "ref" && m != "len") die("invalid value for --model");
            c.o.model = m == "len";
        } else if (a == "-l" || a == "--min_read_len") c.o.min_read_len = (uint32_t)atoi(val(i).c_str());
        else if (a == "-L" || a == "--min_ctg_len") c.o.min_ctg_len = (uint64_t)atoll(val(i).c_str());
        else if (a == "-n" || a == "--max_indel_len") c.o.max_indel_len = atoi(val(i).c_str());
        else if (a == "-s" || a == "--use_supplementary") c.o.use_supplementary = 1;
        else if (a == "-S" || a == "--use_secondary") c.o.use_secondary = 1;
        else if (a == "-a" || a == "--min_map_len") {
            const float f = (float)atof(val(i).c_str());  // option.rs:232,258-259
            c.o.min_map_len = (uint32_t)f;
            c.o.min_map_fra = f - (float)(uint32_t)f;
        } else if (a == "-q" || a == "--min_map_qual") c.o.min_map_qual = atoi(val(i).c_str());
        else if (a == "-c" || a == "--max_clip_len") c.o.max_clip_len = (uint32_t)atoi(val(i).c_str());
        else if (a == "-r" || a == "--use_all_reads") c.o.use_all_reads = 1;
        else if (a == "--min_base_cov") val(i);
        else if (a == "-g" || a == "--gpus") c.gpus = atoi(val(i).c_str());
        else if (a == "--host-inflate") c.host_inflate = true;
        else if (!a.empty() && a[0] == '-' && a.size() > 1) {
            fprintf(stderr, "error: unexpected argument '%s'\n", a.c_str());
            usage();
            exit(2);
        } else pos.push_back(a);
    }
    if (pos.size() < 3) {
        usage();
        exit(2);
    }
    c.bam = pos[0];
    c.fa = pos[1];
    c.yaks.assign(pos.begin() + 2, pos.end());
    return c;
}

}  // namespace

// `nextPolish2 records <in.bam> <reference name> [threads | gpu]`: the raw alignment records of one reference on stdout
// (what IndexedReader::fetch + read hand the worker closure) - a seam for the BGZF / BAI reader tests: host-only with a
// thread count, through the device inflate kernel with `gpu`
int main_records(int argc, char **argv) {
    if (argc < 4) {
        fprintf(stderr, "Usage: nextPolish2 records <in.bam> <reference name> [threads | gpu]\n");
        return 1;
    }
    BamFile bf;
    const bool on_gpu = argc > 4 && std::string(argv[4]) == "gpu";
    bf.threads = argc > 4 && !on_gpu ? std::max(1, atoi(argv[4])) : 4;
    open_bam(argv[2], bf);
    int tid = -1;
    for (size_t r = 0; r < bf.ref_names.size(); r++)
        if (bf.ref_names[r] == argv[3]) tid = (int)r;
    if (tid < 0) die("Faield random access BAM/SAM!");
    np2_ctx *ctx = nullptr;
    if (on_gpu && np2_ctx_create(0, &ctx) != NP2_OK) die(np2_last_error());
    {
        Blob blob(on_gpu);
        fetch_records(bf, tid, blob, bf.threads, ctx);
        fwrite(blob.data(), 1, blob.size(), stdout);
    }
    np2_ctx_destroy(ctx);
    return 0;
}

int main(int argc, char **argv) {
    if (argc >= 2 && std::string(argv[1]) == "count") return main_count(argc, argv);
    if (argc >= 2 && std::string(argv[1]) == "records") return main_records(argc, argv);
    const auto t_start = std::chrono::steady_clock::now();
    auto since = [](std::chrono::steady_clock::time_point t0) {
        return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    };
    std::atomic<uint64_t> us_fetch{0}, us_polish{0}, us_tables{0}, us_lib_total{0}, us_lib_wait{0}, us_destroy{0}, us_call{0};
    const int timing = getenv("NP2_CLI_TIMING") ? std::max(1, atoi(getenv("NP2_CLI_TIMING"))) : 0;  // 2: a line per contig
    Cli cli = parse_args(argc, argv);
    FILE *out = stdout;
    if (cli.out != "stdout") {  // option.rs:308-328: refuse to overwrite
        if (FILE *t = fopen(cli.out.c_str(), "rb")) {
            fclose(t);
            die("\"" + cli.out + "\" already exists!");
        }
        out = fopen(cli.out.c_str(), "wb");
        if (!out) die("Failed to freopen: \"" + cli.out + "\"");
        g_partial_out = cli.out;
    }
    std::vector<Contig> contigs = read_fasta(cli.fa);
    const double s_fasta = since(t_start);
    double s_setup = 0, s_workers = 0;
    const size_t n = contigs.size();
    for (auto &c : contigs)
        if (c.seq.size() >= 0xFFFFFFFFull) die(c.name + " is too long!");  // main.rs:1707-1711
    std::vector<std::vector<uint8_t>> results(n);
    std::vector<uint8_t> done(n, 0);
    // Records leave in input order as soon as every contig before them is done (the reference's writer thread streams
    // them as they finish, main.rs:1845-1851): the finished prefix is written and freed, not held until the end.
    std::mutex write_mu;
    size_t next_write = 0;
    auto flush_ready = [&]() {
        std::lock_guard<std::mutex> lk(write_mu);
        while (next_write < n && done[next_write]) {
            std::vector<uint8_t> &r = results[next_write];
            if (fwrite(r.data(), 1, r.size(), out) != r.size()) die("Failed to write the output!");
            std::vector<uint8_t>().swap(r);
            next_write++;
        }
    };
    std::vector<size_t> todo;
    for (size_t i = 0; i < n; i++) {
        if (contigs[i].seq.size() < cli.o.min_ctg_len) {  // main.rs:1727-1730, no GPU involved
            std::vector<uint32_t> pos(contigs[i].seq.size());
            for (size_t p = 0; p < pos.size(); p++) pos[p] = (uint32_t)p;
            const uint8_t *b = (const uint8_t *)contigs[i].seq.data();
            uint64_t need = np2_format_fasta(contigs[i].name.c_str(), pos.data(), b, pos.size(), cli.o.uppercase,
                                             cli.o.out_pos, nullptr, 0);
            results[i].resize(need);
            np2_format_fasta(contigs[i].name.c_str(), pos.data(), b, pos.size(), cli.o.uppercase, cli.o.out_pos,
                             results[i].data(), need);
            done[i] = 1;
        } else todo.push_back(i);
    }
    if (!todo.empty()) {
        BamFile bf;
        bf.threads = cli.threads > 0 ? cli.threads : (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
        np2_set_host_threads((uint32_t)bf.threads);
        np2_set_stage_timing(timing == 1 ? 1 : 0);  // the per-stage event timers only serve NP2_CLI_TIMING
        open_bam(cli.bam, bf);
        int n_gpu = cli.gpus;
        if (n_gpu <= 0) {
            n_gpu = np2_device_count();
            if (n_gpu == 0) die("no usable GPU (libnp2gpu has no CPU fallback)");
        }
        n_gpu = (int)std::min<size_t>(n_gpu, todo.size());
        // LPT partition by contig length (weight ~ length x depth)
        std::vector<size_t> order = todo;
        std::sort(order.begin(), order.end(), [&](size_t a, size_t b) {
            return contigs[a].seq.size() != contigs[b].seq.size() ? contigs[a].seq.size() > contigs[b].seq.size() : a < b;
        });
        std::vector<std::vector<size_t>> share(n_gpu);
        std::vector<uint64_t> load(n_gpu, 0);
        for (size_t i : order) {
            int g = (int)(std::min_element(load.begin(), load.end()) - load.begin());
            share[g].push_back(i);
            load[g] += contigs[i].seq.size();
        }
        std::mutex bam_mu, err_mu;
        std::string first_err;
        // -S: two passes over every reference's records, like retrieve_secondary_seq_from_bam (secondary.rs:85-150)
        np2_secmap *secmap = nullptr;
        if (cli.o.use_secondary) {
            if (np2_secmap_create(&secmap) != NP2_OK) die(np2_last_error());
            Blob blob;
            for (int pass = 0; pass < 2; pass++)
                for (size_t r = 0; r < bf.ref_names.size(); r++) {
                    fetch_records(bf, (int)r, blob, bf.threads);
                    const int rc = pass == 0 ? np2_secmap_scan_ids(secmap, blob.data(), blob.size())
                                             : np2_secmap_scan_seqs(secmap, blob.data(), blob.size());
                    if (rc != NP2_OK) die(np2_last_error());
                }
        }
        // Per GPU: the tables are staged once, then three to six host threads (one context + stream each, tables shared) take
        // that GPU's contigs in input order, so BGZF decoding / record parsing / upload of one contig overlap the
        // kernels of the other.
        auto fail = [&](const std::string &m) {
            std::lock_guard<std::mutex> lk(err_mu);
            if (first_err.empty()) first_err = m;
        };
        const bool records_on_device = !cli.host_inflate && !(getenv("NP2_CLI_RECORDS_ON_HOST") && atoi(getenv("NP2_CLI_RECORDS_ON_HOST")));
        auto polish_one = [&](np2_ctx *ctx, std::vector<np2_table *> &tabs, std::shared_future<bool> &tabs_ready, size_t i,
                              Blob &blob) -> bool {
            double ms_fetch = 0;
            int tid = -1;
            for (size_t r = 0; r < bf.ref_names.size(); r++)
                if (bf.ref_names[r] == contigs[i].name) tid = (int)r;
            if (tid < 0) return fail("Faield random access BAM/SAM!"), false;
            // Default: the contig's BGZF members go to the device, are inflated there and the records stay there
            // (np2_job_create_bgzf).  -S rewrites the records on the host and --host-inflate asks for zlib: both need the
            // records in host memory (NP2_CLI_RECORDS_ON_HOST=1 does the same with the device inflate, for comparisons).
            const bool on_device = records_on_device && !secmap;
            if (!on_device) {
                // the host inflate uses every thread, so the lanes take turns; the device inflate of a lane runs on the
                // lane's own context next to the kernels of the others
                std::unique_lock<std::mutex> lk(bam_mu, std::defer_lock);
                if (cli.host_inflate) lk.lock();
                const auto t0 = std::chrono::steady_clock::now();
                fetch_records(bf, tid, blob, bf.threads, cli.host_inflate ? nullptr : ctx);
                ms_fetch = since(t0) * 1e3;
                us_fetch += (uint64_t)(ms_fetch * 1e3);
            }
            const auto t_polish = std::chrono::steady_clock::now();
            if (secmap) {  // secondary records get their SEQ (main.rs:1775-1783)
                uint64_t need = 0;
                if (np2_secmap_fill(secmap, blob.data(), blob.size(), nullptr, 0, &need) != NP2_OK)
                    return fail(np2_last_error()), false;
                Blob filled(false);
                filled.resize(need);
                if (np2_secmap_fill(secmap, blob.data(), blob.size(), filled.p, need, &need) != NP2_OK)
                    return fail(np2_last_error()), false;
                blob.swap(filled);
            }
            if (!tabs_ready.get()) return false;  // the staging thread has reported the error
            np2_job *job = nullptr;
            const auto t_call = std::chrono::steady_clock::now();
            // = np2_polish_contig, in its three steps so that NP2_CLI_TIMING=2 can clock them
            double ms_create = 0, ms_upload = 0, ms_run = 0;
            {
                int rc;
                if (on_device) {
                    std::vector<Member> ms;
                    uint64_t u0 = 0, nrec = 0;
                    std::vector<uint64_t> off;
                    std::vector<uint32_t> clen, isz;
                    locate_members(bf, tid, ms, u0, nrec);
                    member_table(ms, off, clen, isz);
                    rc = np2_job_create_bgzf(ctx, (const uint8_t *)contigs[i].seq.data(), (uint32_t)contigs[i].seq.size(), bf.map,
                                             bf.size, off.data(), clen.data(), isz.data(), (uint32_t)ms.size(), u0, nrec,
                                             tabs.data(), (uint32_t)tabs.size(), &cli.o, &job);
                } else {
                    rc = np2_job_create(ctx, (const uint8_t *)contigs[i].seq.data(), (uint32_t)contigs[i].seq.size(), blob.data(),
                                        blob.size(), tabs.data(), (uint32_t)tabs.size(), &cli.o, &job);
                }
                ms_create = since(t_call) * 1e3;
                if (rc == NP2_OK) rc = np2_job_upload(job);
                ms_upload = since(t_call) * 1e3 - ms_create;
                if (rc == NP2_OK) rc = np2_job_run(job, -1);
                ms_run = since(t_call) * 1e3 - ms_create - ms_upload;
                if (rc != NP2_OK) {
                    np2_job_destroy(job);
                    return fail(np2_last_error()), false;
                }
            }
            us_call += (uint64_t)(since(t_call) * 1e6);
            const uint32_t *pos;
            const uint8_t *base;
            uint64_t nb = np2_job_get_consensus(job, cli.o.out_pos ? &pos : nullptr, &base);
            if (cli.o.out_pos) {
                uint64_t need = np2_format_fasta(contigs[i].name.c_str(), pos, base, nb, cli.o.uppercase, 1, nullptr, 0);
                results[i].resize(need);
                np2_format_fasta(contigs[i].name.c_str(), pos, base, nb, cli.o.uppercase, 1, results[i].data(), need);
            } else {
                uint32_t span[2];
                np2_job_get_span(job, &span[0], &span[1]);
                // header needs the first / last position only: format with a two-entry position view
                std::string hdr = ">" + contigs[i].name + " start:" + std::to_string(span[0]) +
                                  " end:" + std::to_string(span[1]) + "\n";
                results[i].assign(hdr.begin(), hdr.end());
                size_t o = results[i].size();
                results[i].resize(o + nb + 1);
                if (cli.o.uppercase)
                    for (uint64_t x = 0; x < nb; x++) results[i][o + x] = (uint8_t)toupper(base[x]);
                else memcpy(results[i].data() + o, base, nb);
                results[i][o + nb] = '\n';
            }
            if (timing) {  // the library's own clocks: device-side step and the upload stages
                const char *names;
                const float *ms;
                const uint32_t *ln;
                const uint32_t ns = np2_job_get_timings(job, &names, &ms, &ln);
                for (uint32_t x = 0; x < ns; x++) {
                    if (!strcmp(names, "total")) us_lib_total += (uint64_t)(ms[x] * 1e3);
                    if (!strcmp(names, "upload:host_wait")) us_lib_wait += (uint64_t)(ms[x] * 1e3);
                    names += strlen(names) + 1;
                }
            }
            const auto t_destroy = std::chrono::steady_clock::now();
            np2_job_destroy(job);
            us_destroy += (uint64_t)(since(t_destroy) * 1e6);
            {
                std::lock_guard<std::mutex> lk(write_mu);  // done[] is read under this lock
                done[i] = 1;
            }
            flush_ready();
            us_polish += (uint64_t)(since(t_polish) * 1e6);
            if (timing > 1)
                fprintf(stderr, "[np2 contig] %s ctx %p: done at %.3f s; fetch %.1f ms, create %.1f, upload %.1f, run %.1f, rest %.1f ms\n",
                        contigs[i].name.c_str(), (void *)ctx, since(t_start), ms_fetch, ms_create, ms_upload, ms_run,
                        since(t_polish) * 1e3 - ms_create - ms_upload - ms_run);
            return true;
        };
        auto worker = [&](int g) {
            np2_ctx *ctx = nullptr;
            if (np2_ctx_create(g, &ctx) != NP2_OK) return fail(np2_last_error());
            // The tables are staged first, on a context of their own (they serve the jobs of every context of this GPU).
            // Staging them next to the lanes' first fetches was tried and gained nothing: page-locking the record
            // buffers, growing the pools and loading the tables all queue on the same driver locks
            // (profiles/r02aw_cli_e2e.txt).
            std::vector<np2_table *> tabs;
            np2_ctx *tab_ctx = nullptr;
            if (np2_ctx_create(g, &tab_ctx) != NP2_OK) return fail(np2_last_error());
            std::promise<bool> tab_promise;
            std::shared_future<bool> tabs_ready = tab_promise.get_future().share();
            {
                const auto t_tab = std::chrono::steady_clock::now();
                bool ok = true;
                for (auto &y : cli.yaks) {
                    np2_table *t = nullptr;
                    if (np2_yak_load(tab_ctx, y.c_str(), &t) != NP2_OK) {
                        fail(np2_last_error());
                        ok = false;
                        break;
                    }
                    tabs.push_back(t);
                }
                us_tables += (uint64_t)(since(t_tab) * 1e6);
                tab_promise.set_value(ok);
                if (!ok) return;
            }
            std::sort(share[g].begin(), share[g].end());
            std::atomic<size_t> next{0};
            // Record buffers, one per lane.  Device inflate: page-locked, so the records come back by DMA and the device
            // gathers the SEQ fields out of them itself.  Host inflate: pageable on purpose — page-locking ~0.5 GB per lane
            // costs more than the library's compaction pass saves on anything but very large inputs (measured,
            // profiles/r01_cli_e2e.txt).  They, the contexts and the tables are NOT released when a lane or a GPU is done:
            // freeing page-locked memory or a pool synchronises the device and stalls the lanes still at work (100 ms
            // and more, profiles/r02av_cli_e2e.txt), and the process ends right after the last record is written.
            // Contigs in flight per GPU.  With the records on the device a lane costs a context and its pool, no page-locked
            // buffer, and its host share is small: five lanes on a 16-core host polish 10 Mbp contigs every 9.2 ms where
            // three need 11.4 (profiles/r02bf_cli_lanes.txt); a third of the cores a GPU has, between 3 and 6.  The paths
            // that bring the records to the host keep three (more only page-lock more memory, profiles/r02ay_cli_e2e.txt).
            const size_t cores_per_gpu = std::max<size_t>(1, std::thread::hardware_concurrency() / (size_t)std::max(1, n_gpu));
            const size_t auto_lanes = records_on_device ? std::min<size_t>(6, std::max<size_t>(3, cores_per_gpu / 3)) : 3;
            const size_t want_lanes = getenv("NP2_CLI_LANES") ? (size_t)std::max(1, atoi(getenv("NP2_CLI_LANES"))) : auto_lanes;
            const size_t n_lanes = std::max<size_t>(1, std::min<size_t>(want_lanes, share[g].size()));
            std::vector<Blob *> blobs;
            for (size_t x = 0; x < n_lanes; x++) blobs.push_back(new Blob(!cli.host_inflate && !records_on_device));
            auto lane = [&](np2_ctx *c, Blob *blob) {
                for (;;) {
                    const size_t x = next.fetch_add(1);
                    if (x >= share[g].size()) break;
                    {
                        std::lock_guard<std::mutex> lk(err_mu);
                        if (!first_err.empty()) break;
                    }
                    if (!polish_one(c, tabs, tabs_ready, share[g][x], *blob)) break;
                }
            };
            std::vector<np2_ctx *> extra;
            std::vector<std::thread> more;
            for (size_t x = 1; x < n_lanes; x++) {
                np2_ctx *c2 = nullptr;
                if (np2_ctx_create(g, &c2) != NP2_OK) break;
                extra.push_back(c2);
                more.emplace_back(lane, c2, blobs[x]);
            }
            lane(ctx, blobs[0]);
            for (auto &t : more) t.join();
        };
        s_setup = since(t_start);
        std::vector<std::thread> th;
        for (int g = 0; g < n_gpu; g++) th.emplace_back(worker, g);
        for (auto &t : th) t.join();
        s_workers = since(t_start);
        np2_secmap_destroy(secmap);
        if (!first_err.empty()) die(first_err);
    }
    const auto t_write = std::chrono::steady_clock::now();
    flush_ready();  // whatever is left (contigs below -L behind the last polished one, or a run without any)
    if (next_write != n) die("internal error: a contig was not written");
    if (out != stdout) fclose(out);
    g_partial_out.clear();
    if (timing) {
        uint64_t bp = 0;
        for (size_t i : todo) bp += contigs[i].seq.size();
        const double wall = since(t_start);
        fprintf(stderr,
                "[np2 timing] wall %.3f s (FASTA read until %.3f, BAM open + GPU probe until %.3f, workers until %.3f), "
                "%.1f Mbp polished -> %.1f Mbp/s end to end; summed over worker threads: table "
                "staging %.3f s, BGZF inflate + record fetch %.3f s, parse + upload + GPU + result %.3f s; "
                "FASTA write %.3f s; inside it: np2_polish_contig %.3f s (of which the library's run step %.3f s, waiting "
                "for the upload %.3f s), job destroy %.3f s\n",
                wall, s_fasta, s_setup, s_workers, bp / 1e6, bp / 1e6 / wall, us_tables / 1e6, us_fetch / 1e6, us_polish / 1e6, since(t_write),
                us_call / 1e6, us_lib_total / 1e6, us_lib_wait / 1e6, us_destroy / 1e6);
    }
    // Every record is written and the file is closed.  Leave without the static destructors and the CUDA runtime's own
    // teardown of contexts, pools and page-locked buffers (more than a second for three lanes; the kernel reclaims the
    // same resources when the process is gone).
    fflush(stdout);
    fflush(stderr);
    if (!todo.empty()) _exit(0);
    return 0;
}
