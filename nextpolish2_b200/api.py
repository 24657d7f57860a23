"""ctypes mirror of include/np2gpu.h.

Names follow the reference's seams: Table ~ KmerInfo (src/utils/kmer.rs:62-221),
polish_contig ~ the worker closure (src/main.rs:1726-1838), Opts ~ Option (src/utils/option.rs:15-41).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib_path():
    return os.path.join(_HERE, "libnp2gpu.so")


class Np2Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("np2gpu error %d: %s" % (code, msg))
        self.code = code


class Opts(C.Structure):
    """np2_opts; defaults = reference src/utils/option.rs:267-292."""
    _fields_ = [("min_kmer_count", C.c_uint32), ("iter_count", C.c_uint32), ("model", C.c_uint32),
                ("min_read_len", C.c_uint32), ("min_ctg_len", C.c_uint64), ("max_indel_len", C.c_int32),
                ("use_supplementary", C.c_uint32), ("use_secondary", C.c_uint32), ("use_all_reads", C.c_uint32),
                ("min_map_len", C.c_uint32), ("min_map_fra", C.c_float), ("min_map_qual", C.c_int32),
                ("max_clip_len", C.c_uint32), ("uppercase", C.c_uint32), ("out_pos", C.c_uint32),
                ("reserved", C.c_uint32)]

    def __init__(self, **kw):
        super().__init__(min_kmer_count=5, iter_count=2, model=0, min_read_len=1000, min_ctg_len=1000000,
                         max_indel_len=20, use_supplementary=0, use_secondary=0, use_all_reads=0, min_map_len=500,
                         min_map_fra=0.5, min_map_qual=1, max_clip_len=100, uppercase=0, out_pos=0, reserved=0)
        for k, v in kw.items():
            setattr(self, k, v)


EXPORTS = [
    "np2_last_error", "np2_opts_default", "np2_ctx_create", "np2_ctx_destroy", "np2_yak_load", "np2_yak_from_arrays",
    "np2_yak_free", "np2_yak_clone", "np2_yak_image", "np2_yak_adopt", "np2_yak_k", "np2_yak_size", "np2_yak_device_bytes", "np2_yak_lookup", "np2_yak_lookup_device",
    "np2_seq_kscore", "np2_bench_gather32", "np2_bench_gather", "np2_l2_fetch_granularity", "np2_polish_contig", "np2_job_create", "np2_job_upload", "np2_job_run", "np2_job_destroy",
    "np2_job_get_consensus", "np2_job_get_span", "np2_job_get_reads", "np2_job_get_msa", "np2_job_get_dp_consensus", "np2_job_get_regions",
    "np2_job_get_candidates", "np2_job_get_dropped", "np2_job_get_pair_weights", "np2_job_get_timings", "np2_job_get_traffic", "np2_job_get_stats", "np2_format_fasta",
    "np2_host_alloc", "np2_host_free", "np2_job_ingest_path", "np2_debug_parse",
    "np2_secmap_create", "np2_secmap_destroy", "np2_secmap_scan_ids", "np2_secmap_scan_seqs", "np2_secmap_fill", "np2_secmap_size",
    "np2_debug_phase", "np2_set_host_threads", "np2_set_stage_timing",
    "np2_bgzf_inflate", "np2_job_create_bgzf",
    "np2_device_count", "np2_count_create", "np2_count_add", "np2_count_distinct", "np2_count_finish", "np2_count_destroy",
]


def load_library():
    """Loads libnp2gpu.so (no CUDA call is made).  Raises if the extension has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    p = lib_path()
    if not os.path.exists(p):
        raise ImportError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nextpolish2_b200 has no CPU fallback)" % p)
    L = C.CDLL(p)
    vp, u64, u32 = C.c_void_p, C.c_uint64, C.c_uint32
    L.np2_last_error.restype = C.c_char_p
    L.np2_opts_default.argtypes = [vp]
    L.np2_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.np2_ctx_destroy.argtypes = [vp]
    L.np2_yak_load.argtypes = [vp, C.c_char_p, C.POINTER(vp)]
    L.np2_yak_from_arrays.argtypes = [vp, u32, vp, vp, u64, C.POINTER(vp)]
    L.np2_yak_free.argtypes = [vp]
    L.np2_yak_clone.argtypes = [vp, vp, C.POINTER(vp)]
    L.np2_yak_image.argtypes = [vp, C.POINTER(vp), C.POINTER(u64), C.POINTER(u32)]
    L.np2_yak_adopt.argtypes = [vp, u32, u64, u32, vp, u64, C.POINTER(vp)]
    L.np2_yak_k.restype = u32
    L.np2_yak_k.argtypes = [vp]
    L.np2_yak_size.restype = u64
    L.np2_yak_size.argtypes = [vp]
    L.np2_yak_device_bytes.restype = u64
    L.np2_yak_device_bytes.argtypes = [vp]
    L.np2_yak_lookup.argtypes = [vp, vp, vp, u64, u32, vp]
    L.np2_yak_lookup_device.argtypes = [vp, vp, vp, u64, u32, vp, u32, C.POINTER(C.c_float)]
    L.np2_seq_kscore.argtypes = [vp, vp, vp, vp, u64, u32, vp]
    L.np2_bench_gather32.argtypes = [vp, u64, u64, u32, C.POINTER(C.c_float)]
    L.np2_bench_gather.argtypes = [vp, u64, u64, u32, u32, C.POINTER(C.c_float)]
    L.np2_l2_fetch_granularity.argtypes = [vp, u32, C.POINTER(u32)]
    L.np2_polish_contig.argtypes = [vp, vp, u32, vp, u64, vp, u32, vp, C.POINTER(vp)]
    L.np2_job_create.argtypes = [vp, vp, u32, vp, u64, vp, u32, vp, C.POINTER(vp)]
    L.np2_job_upload.argtypes = [vp]
    L.np2_host_alloc.argtypes = [u64, C.POINTER(vp)]
    L.np2_bgzf_inflate.argtypes = [vp, vp, u64, vp, vp, vp, u32, u64, u64, vp, C.POINTER(C.c_float)]
    L.np2_job_create_bgzf.argtypes = [vp, vp, u32, vp, u64, vp, vp, vp, u32, u64, u64, vp, u32, vp, C.POINTER(vp)]
    L.np2_host_free.argtypes = [vp]
    L.np2_job_ingest_path.argtypes = [vp]
    L.np2_debug_parse.argtypes = [vp, u64, u32, vp, u32, vp]
    L.np2_job_run.argtypes = [vp, C.c_int32]
    L.np2_job_destroy.argtypes = [vp]
    for name, n in [("np2_job_get_consensus", 2), ("np2_job_get_reads", 6), ("np2_job_get_msa", 5),
                    ("np2_job_get_dp_consensus", 3), ("np2_job_get_regions", 3), ("np2_job_get_candidates", 6),
                    ("np2_job_get_dropped", 1), ("np2_job_get_pair_weights", 2)]:
        f = getattr(L, name)
        f.restype = u64
        f.argtypes = [vp] + [C.POINTER(vp)] * n
    L.np2_job_get_span.restype = u64
    L.np2_job_get_span.argtypes = [vp, C.POINTER(u32), C.POINTER(u32)]
    L.np2_job_get_timings.restype = u32
    L.np2_job_get_timings.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    L.np2_job_get_traffic.argtypes = [vp] + [C.POINTER(u64)] * 5
    L.np2_job_get_stats.argtypes = [vp, vp]
    L.np2_job_get_stats.restype = None
    L.np2_format_fasta.restype = u64
    L.np2_format_fasta.argtypes = [C.c_char_p, vp, vp, u64, C.c_int, C.c_int, vp, u64]
    L.np2_device_count.restype = C.c_int
    L.np2_count_create.argtypes = [vp, u32, C.POINTER(vp)]
    L.np2_count_add.argtypes = [vp, vp, vp, u64]
    L.np2_count_distinct.restype = u64
    L.np2_count_distinct.argtypes = [vp, C.POINTER(u64)]
    L.np2_count_finish.argtypes = [vp, u32, C.c_char_p, C.POINTER(vp)]
    L.np2_count_destroy.argtypes = [vp]
    L.np2_set_host_threads.argtypes = [u32]
    L.np2_set_host_threads.restype = None
    L.np2_set_stage_timing.argtypes = [C.c_int]
    L.np2_set_stage_timing.restype = None
    L.np2_debug_phase.argtypes = [vp, vp, u64, u32, u32, vp, u64, C.POINTER(u64), C.POINTER(u32)]
    L.np2_secmap_create.argtypes = [C.POINTER(vp)]
    L.np2_secmap_destroy.argtypes = [vp]
    L.np2_secmap_scan_ids.argtypes = [vp, vp, u64]
    L.np2_secmap_scan_seqs.argtypes = [vp, vp, u64]
    L.np2_secmap_fill.argtypes = [vp, vp, u64, vp, u64, C.POINTER(u64)]
    L.np2_secmap_size.restype = u64
    L.np2_secmap_size.argtypes = [vp, C.POINTER(u64)]
    _LIB = L
    return L


def _check(rc):
    if rc != 0:
        raise Np2Error(rc, load_library().np2_last_error().decode())


def _arr(ptr, n, dtype):
    dt = np.dtype(dtype)
    if n == 0 or not ptr.value:
        return np.empty(0, dt)
    src = (C.c_char * (n * dt.itemsize)).from_address(ptr.value)
    return np.frombuffer(src, dtype=dt).copy()  # one copy out of library-owned memory


class Context:
    """One GPU + one stream (np2_ctx)."""

    def __init__(self, device=0):
        self.h = C.c_void_p()
        _check(load_library().np2_ctx_create(device, C.byref(self.h)))
        self.device = device

    def close(self):
        if self.h:
            load_library().np2_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PinnedBuffer:
    """Page-locked host bytes from np2_host_alloc: records placed here are gathered by the device itself (K0)."""

    def __init__(self, data):
        data = np.ascontiguousarray(data, np.uint8)
        self.p = C.c_void_p()
        _check(load_library().np2_host_alloc(max(len(data), 1), C.byref(self.p)))
        self.array = np.frombuffer((C.c_char * max(len(data), 1)).from_address(self.p.value), dtype=np.uint8)[:len(data)]
        self.array[:] = data

    def free(self):
        if self.p:
            self.array = None
            load_library().np2_host_free(self.p)
            self.p = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def bgzf_members(buf):
    """Walks the BGZF member headers of a byte buffer (SAM spec 4.1): -> (payload_off u64[], payload_len u32[],
    isize u32[]) of the raw DEFLATE payloads, in file order.  Host only."""
    b = np.ascontiguousarray(buf, np.uint8)
    off, ln, isz = [], [], []
    o, n = 0, len(b)
    while o < n:
        if o + 18 > n or b[o] != 31 or b[o + 1] != 139 or b[o + 2] != 8 or not (b[o + 3] & 4):
            raise ValueError("not a BGZF member at byte %d" % o)
        xlen = int(b[o + 10]) | int(b[o + 11]) << 8
        x, end, bsize = o + 12, o + 12 + xlen, None
        while x + 4 <= end:
            slen = int(b[x + 2]) | int(b[x + 3]) << 8
            if b[x] == 66 and b[x + 1] == 67 and slen == 2:
                bsize = int(b[x + 4]) | int(b[x + 5]) << 8
            x += 4 + slen
        if bsize is None or o + bsize + 1 > n or bsize + 1 < 12 + xlen + 8:
            raise ValueError("bad BGZF member at byte %d" % o)
        total = bsize + 1
        off.append(o + 12 + xlen)
        ln.append(total - 12 - xlen - 8)
        isz.append(int(b[o + total - 4]) | int(b[o + total - 3]) << 8 | int(b[o + total - 2]) << 16 | int(b[o + total - 1]) << 24)
        o += total
    return np.array(off, np.uint64), np.array(ln, np.uint32), np.array(isz, np.uint32)


def bgzf_inflate(ctx, comp, payload_off, payload_len, isize, skip=0, out_len=None, out=None):
    """np2_bgzf_inflate: the members inflated on the device, bytes [skip, skip + out_len) of their concatenation.
    comp: np.uint8 array or PinnedBuffer; out: optional PinnedBuffer / np.uint8 array to receive the bytes.
    -> (np.uint8 array, kernel milliseconds)"""
    carr = comp.array if isinstance(comp, PinnedBuffer) else np.ascontiguousarray(comp, np.uint8)
    po = np.ascontiguousarray(payload_off, np.uint64)
    pl = np.ascontiguousarray(payload_len, np.uint32)
    iz = np.ascontiguousarray(isize, np.uint32)
    total = int(iz.astype(np.uint64).sum())
    if out_len is None:
        out_len = total - skip
    if out is None:
        out = np.empty(max(out_len, 1), np.uint8)
    oarr = out.array if isinstance(out, PinnedBuffer) else out
    ms = C.c_float(0)
    _check(load_library().np2_bgzf_inflate(ctx.h, carr.ctypes.data, len(carr), po.ctypes.data, pl.ctypes.data, iz.ctypes.data,
                                           len(po), skip, out_len, oarr.ctypes.data, C.byref(ms)))
    return oarr[:out_len], ms.value


def debug_phase(keys, vals, model=0, use_all_reads=False, with_path=False):
    """Host-only test seam: reduced agreement edges -> read orders that phasing drops (np2_debug_phase)."""
    keys = np.ascontiguousarray(keys, np.uint64)
    vals = np.ascontiguousarray(vals, np.int64)
    n, path = C.c_uint64(), C.c_uint32()
    out = np.empty(max(len(keys) * 2 + 1, 1), np.uint32)
    _check(load_library().np2_debug_phase(keys.ctypes.data, vals.ctypes.data, len(keys), model, int(use_all_reads),
                                          out.ctypes.data, len(out), C.byref(n), C.byref(path)))
    res = out[:n.value].copy()
    return (res, path.value) if with_path else res


def debug_parse(bam, tlen, opts=None, threads=0):
    """Host-only test seam: record parse + filter over `threads` speculative byte ranges -> counts and a digest."""
    bam = np.ascontiguousarray(bam, np.uint8)
    opts = opts or Opts()
    out = (C.c_uint64 * 6)()
    _check(load_library().np2_debug_parse(bam.ctypes.data, len(bam), tlen, C.byref(opts), threads, out))
    return dict(zip(["records", "reads", "ops", "columns", "fallback", "digest"], [int(x) for x in out]))


def bench_gather(ctx, buf_bytes, n_loads, block_bytes=32, repeat=5):
    """Mean ms of n_loads random aligned block reads of block_bytes (32 / 64 / 128) over a buf_bytes scratch buffer."""
    ms = C.c_float()
    _check(load_library().np2_bench_gather(ctx.h, buf_bytes, n_loads, block_bytes, repeat, C.byref(ms)))
    return ms.value


def l2_fetch_granularity(ctx, nbytes=0):
    """Sets (nbytes = 32 / 64 / 128) or queries (0) the device's L2 fetch granularity hint; returns the previous value."""
    prev = C.c_uint32()
    _check(load_library().np2_l2_fetch_granularity(ctx.h, nbytes, C.byref(prev)))
    return prev.value


def bench_gather32(ctx, buf_bytes, n_loads, repeat=5):
    """Measured random 32-byte-sector read rate (ms per n_loads loads): K5's roofline denominator."""
    ms = C.c_float()
    _check(load_library().np2_bench_gather32(ctx.h, buf_bytes, n_loads, repeat, C.byref(ms)))
    return ms.value


class Table:
    """A yak k-mer count table staged in HBM (KmerInfo, src/utils/kmer.rs:62-221)."""

    def __init__(self, ctx, handle):
        self.ctx = ctx
        self.h = handle

    @classmethod
    def load(cls, ctx, path):
        h = C.c_void_p()
        _check(load_library().np2_yak_load(ctx.h, path.encode(), C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def from_arrays(cls, ctx, k, hashes, counts):
        hashes = np.ascontiguousarray(hashes, np.uint64)
        counts = np.ascontiguousarray(counts, np.uint16)
        h = C.c_void_p()
        _check(load_library().np2_yak_from_arrays(ctx.h, k, hashes.ctypes.data, counts.ctypes.data, len(hashes), C.byref(h)))
        return cls(ctx, h)

    def clone(self, ctx):
        """A replica on another GPU of this process: one peer copy of the device image (np2_yak_clone)."""
        h = C.c_void_p()
        _check(load_library().np2_yak_clone(ctx.h, self.h, C.byref(h)))
        return Table(ctx, h)

    def image(self):
        """(device pointer, bytes, buckets per sub-table) of the staged table: what a broadcast moves."""
        p, n, nb = C.c_void_p(), C.c_uint64(), C.c_uint32()
        _check(load_library().np2_yak_image(self.h, C.byref(p), C.byref(n), C.byref(nb)))
        return p.value, n.value, nb.value

    @classmethod
    def adopt(cls, ctx, k, n_keys, buckets_per_subtable, d_image_ptr, nbytes):
        """A table from a received device image (np2_yak_adopt copies it)."""
        h = C.c_void_p()
        _check(load_library().np2_yak_adopt(ctx.h, k, n_keys, buckets_per_subtable, d_image_ptr, nbytes, C.byref(h)))
        return cls(ctx, h)

    @property
    def k(self):
        return load_library().np2_yak_k(self.h)

    def __len__(self):
        return load_library().np2_yak_size(self.h)

    @property
    def device_bytes(self):
        return load_library().np2_yak_device_bytes(self.h)

    def lookup(self, hashes, min_count=5):
        """insert + retrieve_kmers + get (kmer.rs:113-170) for a batch: count if present and >= min_count else 0."""
        hashes = np.ascontiguousarray(hashes, np.uint64)
        out = np.empty(len(hashes), np.uint16)
        _check(load_library().np2_yak_lookup(self.ctx.h, self.h, hashes.ctypes.data, len(hashes), min_count, out.ctypes.data))
        return out

    def lookup_device(self, d_hashes_ptr, n, d_out_ptr, min_count=5, repeat=1):
        """Device-resident batch; returns the mean kernel time in ms (CUDA events on the library stream)."""
        ms = C.c_float()
        _check(load_library().np2_yak_lookup_device(self.ctx.h, self.h, d_hashes_ptr, n, min_count, d_out_ptr, repeat, C.byref(ms)))
        return ms.value

    def seq_kscore(self, seqs, min_count=5):
        """kscore of byte strings: min filtered count over their canonical k-mers (main.rs:761-769)."""
        off = np.zeros(len(seqs) + 1, np.uint64)
        off[1:] = np.cumsum([len(s) for s in seqs])
        pool = np.frombuffer(b"".join(bytes(s) for s in seqs) + b"\0", np.uint8).copy()
        out = np.empty(len(seqs), np.uint16)
        _check(load_library().np2_seq_kscore(self.ctx.h, self.h, pool.ctypes.data, off.ctypes.data, len(seqs), min_count, out.ctypes.data))
        return out

    def free(self):
        if self.h:
            load_library().np2_yak_free(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Job:
    """One contig (np2_job): create -> upload -> run, with stage dumps for the parity tests."""

    def __init__(self, ctx, contig, bam, tables, opts=None):
        self.ctx = ctx
        self.contig = np.ascontiguousarray(contig, np.uint8)
        self.bam = np.ascontiguousarray(bam, np.uint8)
        self.tables = list(tables)
        self.opts = opts or Opts()
        tp = (C.c_void_p * len(self.tables))(*[t.h for t in self.tables])
        self.h = C.c_void_p()
        _check(load_library().np2_job_create(ctx.h, self.contig.ctypes.data, len(self.contig), self.bam.ctypes.data,
                                             len(self.bam), tp, len(self.tables), C.byref(self.opts), C.byref(self.h)))

    @classmethod
    def from_bgzf(cls, ctx, contig, comp, payload_off, payload_len, isize, skip, rec_len, tables, opts=None):
        """np2_job_create_bgzf: the contig's BGZF members in (comp = the compressed file bytes), the records never
        leave the device."""
        self = cls.__new__(cls)
        self.ctx = ctx
        self.contig = np.ascontiguousarray(contig, np.uint8)
        self.bam = comp.array if isinstance(comp, PinnedBuffer) else np.ascontiguousarray(comp, np.uint8)
        self.tables = list(tables)
        self.opts = opts or Opts()
        po = np.ascontiguousarray(payload_off, np.uint64)
        pl = np.ascontiguousarray(payload_len, np.uint32)
        iz = np.ascontiguousarray(isize, np.uint32)
        tp = (C.c_void_p * len(self.tables))(*[t.h for t in self.tables])
        self.h = C.c_void_p()
        _check(load_library().np2_job_create_bgzf(ctx.h, self.contig.ctypes.data, len(self.contig), self.bam.ctypes.data,
                                                  len(self.bam), po.ctypes.data, pl.ctypes.data, iz.ctypes.data, len(po), skip,
                                                  rec_len, tp, len(self.tables), C.byref(self.opts), C.byref(self.h)))
        return self

    def upload(self):
        _check(load_library().np2_job_upload(self.h))
        return self

    @property
    def ingest_path(self):
        """1 = SEQ gathered by the device from page-locked records, 2 = compacted on the host, 0 = passthrough."""
        return load_library().np2_job_ingest_path(self.h)

    def run(self, dump_iter=-1):
        _check(load_library().np2_job_run(self.h, dump_iter))
        return self

    def _get(self, name, n):
        ptrs = [C.c_void_p() for _ in range(n)]
        cnt = getattr(load_library(), name)(self.h, *[C.byref(p) for p in ptrs])
        return cnt, ptrs

    def consensus(self):
        n, p = self._get("np2_job_get_consensus", 2)
        return _arr(p[0], n, np.uint32), _arr(p[1], n, np.uint8)

    def bases(self, copy=True):
        """(first_pos, last_pos, bases): all the FASTA record needs, without materialising per-base positions.
        copy=False returns a view of the library's page-locked result buffer (valid until destroy())."""
        base = C.c_void_p()
        n = load_library().np2_job_get_consensus(self.h, None, C.byref(base))
        f, l = C.c_uint32(), C.c_uint32()
        load_library().np2_job_get_span(self.h, C.byref(f), C.byref(l))
        if copy or n == 0:
            return f.value, l.value, _arr(base, n, np.uint8)
        return f.value, l.value, np.frombuffer((C.c_char * n).from_address(base.value), dtype=np.uint8)

    def reads(self):
        n, p = self._get("np2_job_get_reads", 6)
        nib_off = _arr(p[3], n + 1 if n else 0, np.uint64)
        return {"rec_idx": _arr(p[0], n, np.int32), "t_s": _arr(p[1], n, np.uint32), "t_e": _arr(p[2], n, np.uint32),
                "nib_off": nib_off, "nib": _arr(p[4], int(nib_off[-1]) if n else 0, np.uint8),
                "blank": _arr(p[5], n, np.uint8)}

    def msa(self):
        n, p = self._get("np2_job_get_msa", 5)
        return {"off": _arr(p[0], len(self.contig) + 1 if n else 0, np.uint64), "bases": _arr(p[1], n, np.uint16),
                "delta": _arr(p[2], n, np.uint16), "count": _arr(p[3], n, np.uint32), "besti": _arr(p[4], n, np.uint32)}

    def dp_consensus(self):
        n, p = self._get("np2_job_get_dp_consensus", 3)
        return {"pos": _arr(p[0], n, np.uint32), "base": _arr(p[1], n, np.uint8), "flags": _arr(p[2], n, np.uint8)}

    def regions(self):
        n, p = self._get("np2_job_get_regions", 3)
        lab = _arr(p[2], n, np.uint8)
        return {"start": _arr(p[0], n, np.uint32), "end": _arr(p[1], n, np.uint32), "lable": lab}

    def candidates(self):
        nreg = len(self.regions()["start"])
        n, p = self._get("np2_job_get_candidates", 6)
        seq_off = _arr(p[4], n + 1 if nreg else 0, np.uint64)
        return {"roff": _arr(p[0], nreg + 1 if nreg else 0, np.uint64), "order": _arr(p[1], n, np.uint32),
                "kscore": _arr(p[2], n, np.uint16), "kmer": _arr(p[3], n, np.uint64), "seq_off": seq_off,
                "seq": _arr(p[5], int(seq_off[-1]) if len(seq_off) else 0, np.uint8)}

    def dropped(self):
        n, p = self._get("np2_job_get_dropped", 1)
        return _arr(p[0], n, np.uint32)

    def pair_weights(self):
        """(keys a << 32 | b ascending, vals #agree + #differ * (2^32 - 1)) of the dumped iteration (main.rs:953-992)."""
        n, p = self._get("np2_job_get_pair_weights", 2)
        return _arr(p[0], n, np.uint64), _arr(p[1], n, np.int64)

    def timings(self):
        names, ms, ln = C.c_void_p(), C.c_void_p(), C.c_void_p()
        n = load_library().np2_job_get_timings(self.h, C.byref(names), C.byref(ms), C.byref(ln))
        if n == 0:
            return {}
        msv = _arr(ms, n, np.float32)
        lnv = _arr(ln, n, np.uint32)
        out, off = {}, 0
        raw = C.string_at(names.value, 4096)
        for i in range(n):
            end = raw.index(b"\0", off)
            out[raw[off:end].decode()] = (float(msv[i]), int(lnv[i]))
            off = end + 1
        return out

    def traffic(self):
        v = [C.c_uint64() for _ in range(5)]
        load_library().np2_job_get_traffic(self.h, *[C.byref(x) for x in v])
        return dict(zip(["h2d_bytes", "d2h_bytes", "kernel_launches", "alignment_columns", "probes"], [x.value for x in v]))

    def stats(self):
        v = (C.c_uint64 * 12)()
        load_library().np2_job_get_stats(self.h, v)
        return dict(zip(["records", "groups", "runs", "dp_bases", "regions", "read_region_pairs", "pair_edges",
                         "iterations_built", "speculative_passes", "repeated_passes", "host_syncs"], [int(x) for x in v]))

    def destroy(self):
        if self.h:
            load_library().np2_job_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def polish_contig(ctx, contig, bam, tables, opts=None):
    """The worker closure (main.rs:1726-1838): host buffers in, consensus (pos, base) out."""
    j = Job(ctx, contig, bam, tables, opts)
    try:
        j.upload().run(-1)
        return j.consensus()
    finally:
        j.destroy()


class Counter:
    """`yak count` on the device (np2_count_*): add reads, then dump a .yak file and / or stage a Table."""

    def __init__(self, ctx, k):
        self.ctx = ctx
        self.h = C.c_void_p()
        _check(load_library().np2_count_create(ctx.h, k, C.byref(self.h)))

    def add(self, seqs):
        """seqs: list of bytes / uint8 arrays"""
        arrs = [np.frombuffer(x, np.uint8) if isinstance(x, (bytes, bytearray)) else np.ascontiguousarray(x, np.uint8) for x in seqs]
        off = np.zeros(len(arrs) + 1, np.uint64)
        off[1:] = np.cumsum([len(a) for a in arrs])
        cat = np.concatenate(arrs) if arrs else np.empty(0, np.uint8)
        _check(load_library().np2_count_add(self.h, cat.ctypes.data, off.ctypes.data, len(arrs)))
        return self

    @property
    def distinct(self):
        n = C.c_uint64()
        d = load_library().np2_count_distinct(self.h, C.byref(n))
        return d, n.value

    def finish(self, min_count=1, dump_path=None, table=False):
        t = C.c_void_p()
        _check(load_library().np2_count_finish(self.h, min_count, dump_path.encode() if dump_path else None,
                                               C.byref(t) if table else None))
        return Table(self.ctx, t) if table else None

    def close(self):
        if self.h:
            load_library().np2_count_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def set_host_threads(n):
    """Host threads one call may use for record parsing / SEQ compaction (np2_set_host_threads)."""
    load_library().np2_set_host_threads(int(n))


def set_stage_timing(on):
    """Per-stage CUDA-event timers of Job.timings() on / off (np2_set_stage_timing); "total" and host phases stay."""
    load_library().np2_set_stage_timing(1 if on else 0)


class SecondarySeqs:
    """-S / --use_secondary: the reference's sec_seqs map (src/utils/secondary.rs:85-150) over record blobs.

    scan_ids(blob) for every contig, then scan_seqs(blob) for every contig, then fill(blob) per contig gives the
    record blob np2_polish_contig takes with Opts(use_secondary=1).  Host only."""

    def __init__(self):
        self.h = C.c_void_p()
        _check(load_library().np2_secmap_create(C.byref(self.h)))

    def scan_ids(self, bam):
        bam = np.ascontiguousarray(bam, np.uint8)
        _check(load_library().np2_secmap_scan_ids(self.h, bam.ctypes.data, len(bam)))
        return self

    def scan_seqs(self, bam):
        bam = np.ascontiguousarray(bam, np.uint8)
        _check(load_library().np2_secmap_scan_seqs(self.h, bam.ctypes.data, len(bam)))
        return self

    def fill(self, bam):
        bam = np.ascontiguousarray(bam, np.uint8)
        need = C.c_uint64()
        _check(load_library().np2_secmap_fill(self.h, bam.ctypes.data, len(bam), None, 0, C.byref(need)))
        out = np.empty(need.value, np.uint8)
        _check(load_library().np2_secmap_fill(self.h, bam.ctypes.data, len(bam), out.ctypes.data, len(out), C.byref(need)))
        return out

    @property
    def counts(self):
        n = C.c_uint64()
        ids = load_library().np2_secmap_size(self.h, C.byref(n))
        return ids, n.value

    def close(self):
        if self.h:
            load_library().np2_secmap_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def format_fasta(tid, pos, base, uppercase=False, out_pos=False):
    """display_consensusbase_vec (main.rs:607-645)."""
    pos = np.ascontiguousarray(pos, np.uint32)
    base = np.ascontiguousarray(base, np.uint8)
    L = load_library()
    n = L.np2_format_fasta(tid.encode(), pos.ctypes.data, base.ctypes.data, len(base), int(uppercase), int(out_pos), None, 0)
    out = np.empty(n, np.uint8)
    L.np2_format_fasta(tid.encode(), pos.ctypes.data, base.ctypes.data, len(base), int(uppercase), int(out_pos), out.ctypes.data, n)
    return bytes(out)
