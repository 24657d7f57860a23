"""Synthetic input generator (ctypes wrapper over libnp2synth.so).

Host-only tooling for tests and bench.py: deterministic genomes, truth
haplotypes, HiFi reads with by-construction alignments as raw BAM records,
yak table synthesis, and BAM/BAI/yak/FASTA writers.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libnp2synth.so")
_lib = None


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("np2_synth.cpp", "np2_align.cpp")]
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", _LIB_PATH] + srcs + ["-lz", "-lpthread"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        L.np2s_genome.argtypes = [C.c_uint64, C.c_uint32, C.c_double, C.c_double, C.c_void_p]
        L.np2s_contig_make.restype = C.c_void_p
        L.np2s_contig_make.argtypes = [C.c_uint64, C.c_void_p, C.c_uint32, C.c_int32] + [C.c_double] * 11 + [C.c_int, C.c_int]
        L.np2s_contig_sizes.argtypes = [C.c_void_p] + [C.POINTER(C.c_uint64)] * 4
        L.np2s_contig_copy.argtypes = [C.c_void_p] * 4
        L.np2s_contig_free.argtypes = [C.c_void_p]
        L.np2s_table.restype = C.c_uint64
        L.np2s_table.argtypes = [C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_double, C.c_uint32,
                                 C.c_void_p, C.c_void_p, C.c_uint64]
        L.np2s_table_mt.restype = C.c_uint64
        L.np2s_table_mt.argtypes = [C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_double, C.c_uint32,
                                    C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32]
        L.np2s_write_yak.argtypes = [C.c_char_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64]
        L.np2s_write_short_reads.argtypes = [C.c_char_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint32, C.c_double,
                                             C.c_uint32, C.c_double]
        L.np2s_write_bam.argtypes = [C.c_char_p, C.c_uint32, C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.np2s_align.restype = C.c_void_p
        L.np2s_align.argtypes = [C.c_void_p, C.c_uint32, C.c_int32, C.c_void_p, C.c_void_p, C.c_char_p, C.c_uint64, C.c_int]
        L.np2s_aln_size.restype = C.c_uint64
        L.np2s_aln_size.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        L.np2s_aln_copy.argtypes = [C.c_void_p, C.c_void_p]
        L.np2s_aln_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def read_fasta(path):
    """[(name, bytes)] from a FASTA/FASTQ file, .gz or not (first word of the header, like kseq)."""
    import gzip
    op = gzip.open if path.endswith(".gz") else open
    out, name, parts = [], None, []
    with op(path, "rb") as f:
        data = f.read()
    lines = data.split(b"\n")
    if data[:1] == b">":
        for line in lines:
            if line[:1] == b">":
                if name is not None:
                    out.append((name, b"".join(parts)))
                name, parts = line[1:].split()[0].decode(), []
            elif line:
                parts.append(line.strip())
        if name is not None:
            out.append((name, b"".join(parts)))
    else:
        for i in range(0, len(lines) - 3, 4):
            out.append((lines[i][1:].split()[0].decode(), lines[i + 1].strip()))
    return out


def align_reads(contig, reads, ref_id=0, threads=8):
    """Mini-aligner (np2_align.cpp): reads = [(name, bytes)] -> (BAM records blob sorted by position, #aligned)."""
    T = np.ascontiguousarray(contig, dtype=np.uint8)
    seqs = np.frombuffer(b"".join(r[1] for r in reads), np.uint8)
    off = np.zeros(len(reads) + 1, np.uint64)
    off[1:] = np.cumsum([len(r[1]) for r in reads])
    names = b"".join(r[0].encode() + b"\0" for r in reads)
    h = lib().np2s_align(T.ctypes.data, len(T), ref_id, seqs.ctypes.data, off.ctypes.data, names, len(reads), threads)
    n = C.c_uint64()
    size = lib().np2s_aln_size(h, C.byref(n))
    bam = np.empty(size, np.uint8)
    lib().np2s_aln_copy(h, bam.ctypes.data)
    lib().np2s_aln_free(h)
    return bam, n.value


def genome(seed, length, gc=0.41, tandem_frac=0.0):
    out = np.empty(length, dtype=np.uint8)
    lib().np2s_genome(seed, length, gc, tandem_frac, out.ctypes.data)
    return out


def make_contig(seed, A, ref_id=0, depth=30.0, asm_err=2e-5, het=0.0, read_err=0.002, mean_len=15000.0, sd_len=2000.0,
                min_len=5000.0, max_len=25000.0, frac_clip=0.0, frac_lowq=0.0, frac_supp=0.0, eqx=False, threads=8):
    A = np.ascontiguousarray(A, dtype=np.uint8)
    h = lib().np2s_contig_make(seed, A.ctypes.data, len(A), ref_id, depth, asm_err, het, read_err, mean_len, sd_len,
                               min_len, max_len, frac_clip, frac_lowq, frac_supp, int(eqx), threads)
    s = [C.c_uint64() for _ in range(4)]
    lib().np2s_contig_sizes(h, *[C.byref(x) for x in s])
    hap1 = np.empty(s[0].value, np.uint8)
    hap2 = np.empty(s[1].value, np.uint8)
    bam = np.empty(s[2].value, np.uint8)
    lib().np2s_contig_copy(h, hap1.ctypes.data, hap2.ctypes.data if len(hap2) else None, bam.ctypes.data)
    lib().np2s_contig_free(h)
    return {"contig": A, "hap1": hap1, "hap2": hap2, "bam": bam, "n_reads": s[3].value}


def _seq_ptrs(seqs):
    seqs = [np.ascontiguousarray(x, dtype=np.uint8) for x in seqs if len(x)]
    ptrs = (C.c_void_p * len(seqs))(*[x.ctypes.data for x in seqs])
    lens = (C.c_uint64 * len(seqs))(*[len(x) for x in seqs])
    return seqs, ptrs, lens


def make_table(seed, k, seqs, mean_count=40.0, keep_min=2):
    """All canonical k-mers of seqs -> (hashes u64, counts u16)."""
    seqs, ptrs, lens = _seq_ptrs(seqs)
    n = lib().np2s_table(seed, k, ptrs, lens, len(seqs), mean_count, keep_min, None, None, 0)
    h = np.empty(n, np.uint64)
    c = np.empty(n, np.uint16)
    n2 = lib().np2s_table(seed, k, ptrs, lens, len(seqs), mean_count, keep_min, h.ctypes.data, c.ctypes.data, n)
    assert n2 == n
    return h, c


def make_table_mt(seed, k, seqs, mean_count=40.0, keep_min=2, threads=8):
    """make_table built by `threads` workers in one call; counts are a function of (seed, k, hash, multiplicity), so the
    table is independent of the thread count (but not identical to make_table's sequential draw)."""
    seqs, ptrs, lens = _seq_ptrs(seqs)
    cap = int(sum(len(x) for x in seqs))
    h = np.empty(cap, np.uint64)
    c = np.empty(cap, np.uint16)
    n = lib().np2s_table_mt(seed, k, ptrs, lens, len(seqs), mean_count, keep_min, h.ctypes.data, c.ctypes.data, cap, threads)
    assert n <= cap
    return h[:n].copy(), c[:n].copy()


def write_yak(path, k, hashes, counts):
    hashes = np.ascontiguousarray(hashes, np.uint64)
    counts = np.ascontiguousarray(counts, np.uint16)
    if lib().np2s_write_yak(path.encode(), k, hashes.ctypes.data, counts.ctypes.data, len(hashes)) != 0:
        raise OSError("cannot write " + path)


def write_short_reads(path, seed, seqs, depth_each=30.0, rlen=150, sub_rate=0.001):
    seqs, ptrs, lens = _seq_ptrs(seqs)
    if lib().np2s_write_short_reads(path.encode(), seed, ptrs, lens, len(seqs), depth_each, rlen, sub_rate) != 0:
        raise OSError("cannot write " + path)


def write_bam(path, names, ref_lens, blobs, level=1):
    """Coordinate-sorted BAM + .bai from per-contig raw record blobs (refID of blob i must be i)."""
    blobs = [np.ascontiguousarray(b, np.uint8) for b in blobs]
    nm = b"".join(n.encode() + b"\0" for n in names)
    rl = (C.c_uint32 * len(names))(*ref_lens)
    ptrs = (C.c_void_p * len(blobs))(*[b.ctypes.data for b in blobs])
    lens = (C.c_uint64 * len(blobs))(*[len(b) for b in blobs])
    if lib().np2s_write_bam(path.encode(), len(names), nm, rl, ptrs, lens, level) != 0:
        raise OSError("cannot write " + path)


def write_fasta(path, names, seqs, width=0):
    import gzip
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "wb") as f:
        for n, s in zip(names, seqs):
            f.write(b">" + n.encode() + b"\n")
            b = bytes(np.asarray(s, np.uint8))
            if width:
                for i in range(0, len(b), width):
                    f.write(b[i:i + width] + b"\n")
            else:
                f.write(b + b"\n")


def bam_record(ref_id, pos, cigar, seq, flag=0, mapq=60, name="x"):
    """Hand-crafted BAM record for edge-case tests. cigar: [(op_char, len)], seq: str."""
    ops = "MIDNSHP=X"
    enc = {c: i for i, c in enumerate("=ACMGRSVTWYHKDBN")}
    nm = name.encode() + b"\0"
    cg = np.array([(l << 4) | ops.index(o) for o, l in cigar], dtype="<u4").tobytes()
    l_seq = len(seq)
    sq = bytearray((l_seq + 1) // 2)
    for i, ch in enumerate(seq):
        sq[i >> 1] |= enc[ch.upper()] << (4 if i % 2 == 0 else 0)
    body = (np.array([ref_id, pos], "<i4").tobytes() + bytes([len(nm), mapq]) + np.array([4680, len(cigar), flag], "<u2").tobytes()
            + np.array([l_seq, -1, -1, 0], "<i4").tobytes() + nm + cg + bytes(sq) + b"\xff" * l_seq)
    return np.frombuffer(np.array([len(body)], "<i4").tobytes() + body, np.uint8)
