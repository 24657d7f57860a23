"""ctypes wrapper over the CPU oracle (oracle/libnp2oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  Never imported by the
nextpolish2_b200 package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libnp2oracle.so")
_lib = None


class Opts(C.Structure):
    """np2o_opts == np2_opts (defaults: reference src/utils/option.rs:267-292)."""
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


def build(force=False):
    src = os.path.join(_HERE, "np2_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        L.np2o_last_error.restype = C.c_char_p
        L.np2o_yak_hash64.restype = C.c_uint64
        L.np2o_yak_hash64.argtypes = [C.c_uint64, C.c_uint64]
        L.np2o_yak_hash64_64.restype = C.c_uint64
        L.np2o_yak_hash64_64.argtypes = [C.c_uint64]
        L.np2o_yak_hash_long.restype = C.c_uint64
        L.np2o_yak_hash_long.argtypes = [C.c_void_p]
        L.np2o_seq_hashes.restype = C.c_int64
        L.np2o_seq_hashes.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_uint64]
        L.np2o_table_load.restype = C.c_void_p
        L.np2o_table_load.argtypes = [C.c_char_p]
        L.np2o_table_from_arrays.restype = C.c_void_p
        L.np2o_table_from_arrays.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64]
        L.np2o_table_destroy.argtypes = [C.c_void_p]
        L.np2o_table_k.restype = C.c_uint32
        L.np2o_table_k.argtypes = [C.c_void_p]
        L.np2o_table_size.restype = C.c_uint64
        L.np2o_table_size.argtypes = [C.c_void_p]
        L.np2o_table_lookup.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p]
        L.np2o_table_set_stream_scan.argtypes = [C.c_void_p, C.c_int]
        L.np2o_job_create.restype = C.c_void_p
        L.np2o_job_create.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint32]
        L.np2o_job_run.argtypes = [C.c_void_p, C.c_int32]
        L.np2o_job_destroy.argtypes = [C.c_void_p]
        L.np2o_get_seconds.restype = C.c_double
        L.np2o_get_seconds.argtypes = [C.c_void_p]
        for name, n in [("np2o_get_reads", 6), ("np2o_get_msa", 5), ("np2o_get_dp_consensus", 3), ("np2o_get_regions", 3),
                        ("np2o_get_candidates", 6), ("np2o_get_dropped", 1), ("np2o_get_consensus", 2),
                        ("np2o_get_pair_weights", 2)]:
            f = getattr(L, name)
            f.restype = C.c_uint64
            f.argtypes = [C.c_void_p] + [C.POINTER(C.c_void_p)] * n
        L.np2o_order_exposure.argtypes = [C.c_void_p, C.c_int]
        L.np2o_debug_phase.restype = C.c_int64
        L.np2o_debug_phase.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint64]
        L.np2o_format_fasta.restype = C.c_uint64
        L.np2o_format_fasta.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_uint64]
        _lib = L
    return _lib


class OracleError(RuntimeError):
    pass


def _arr(ptr, n, dtype):
    if n == 0 or not ptr.value:
        return np.empty(0, dtype)
    dt = np.dtype(dtype)
    buf = C.string_at(ptr.value, n * dt.itemsize)
    return np.frombuffer(buf, dtype=dt).copy()


def yak_hash64(key, mask):
    return lib().np2o_yak_hash64(key, mask)


def seq_hashes(seq, k):
    seq = np.ascontiguousarray(np.frombuffer(seq, np.uint8) if isinstance(seq, (bytes, bytearray)) else seq, np.uint8)
    out = np.empty(max(len(seq), 1), np.uint64)
    n = lib().np2o_seq_hashes(seq.ctypes.data, len(seq), k, out.ctypes.data, len(out))
    if n < 0:
        raise OracleError(lib().np2o_last_error().decode())
    return out[:n].copy()


class Table:
    def __init__(self, handle):
        if not handle:
            raise OracleError(lib().np2o_last_error().decode())
        self.h = handle

    @classmethod
    def load(cls, path):
        return cls(lib().np2o_table_load(path.encode()))

    @classmethod
    def from_arrays(cls, k, hashes, counts):
        hashes = np.ascontiguousarray(hashes, np.uint64)
        counts = np.ascontiguousarray(counts, np.uint16)
        return cls(lib().np2o_table_from_arrays(k, hashes.ctypes.data, counts.ctypes.data, len(hashes)))

    @property
    def k(self):
        return lib().np2o_table_k(self.h)

    def __len__(self):
        return lib().np2o_table_size(self.h)

    def set_stream_scan(self, on):
        lib().np2o_table_set_stream_scan(self.h, int(on))

    def lookup(self, hashes, min_count=5):
        hashes = np.ascontiguousarray(hashes, np.uint64)
        out = np.empty(len(hashes), np.uint16)
        lib().np2o_table_lookup(self.h, hashes.ctypes.data, len(hashes), min_count, out.ctypes.data)
        return out

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.np2o_table_destroy(self.h)
            self.h = None


class Job:
    """One contig through the oracle; keeps stage dumps of iteration `dump_iter`."""

    def __init__(self, contig, bam, tables, opts=None, dump_iter=0):
        self.contig = np.ascontiguousarray(contig, np.uint8)
        self.bam = np.ascontiguousarray(bam, np.uint8)
        self.opts = opts or Opts()
        self.tables = tables
        tp = (C.c_void_p * len(tables))(*[t.h for t in tables])
        self.h = lib().np2o_job_create(self.contig.ctypes.data, len(self.contig), self.bam.ctypes.data, len(self.bam),
                                       C.byref(self.opts), tp, len(tables))
        rc = lib().np2o_job_run(self.h, dump_iter)
        if rc != 0:
            msg = lib().np2o_last_error().decode()
            lib().np2o_job_destroy(self.h)
            self.h = None
            raise OracleError(msg)

    def _get(self, name, dtypes):
        ptrs = [C.c_void_p() for _ in dtypes]
        n = getattr(lib(), name)(self.h, *[C.byref(p) for p in ptrs])
        return n, ptrs

    @property
    def seconds(self):
        return lib().np2o_get_seconds(self.h)

    def reads(self):
        n, p = self._get("np2o_get_reads", range(6))
        nib_off = _arr(p[3], n + 1, np.uint64)
        return {"rec_idx": _arr(p[0], n, np.int32), "t_s": _arr(p[1], n, np.uint32), "t_e": _arr(p[2], n, np.uint32),
                "nib_off": nib_off, "nib": _arr(p[4], int(nib_off[-1]) if n else 0, np.uint8),
                "blank": _arr(p[5], n, np.uint8)}

    def msa(self):
        n, p = self._get("np2o_get_msa", range(5))
        return {"off": _arr(p[0], len(self.contig) + 1 if n else 0, np.uint64), "bases": _arr(p[1], n, np.uint16),
                "delta": _arr(p[2], n, np.uint16), "count": _arr(p[3], n, np.uint32), "besti": _arr(p[4], n, np.uint32)}

    def dp_consensus(self):
        n, p = self._get("np2o_get_dp_consensus", range(3))
        return {"pos": _arr(p[0], n, np.uint32), "base": _arr(p[1], n, np.uint8), "flags": _arr(p[2], n, np.uint8)}

    def regions(self):
        n, p = self._get("np2o_get_regions", range(3))
        return {"start": _arr(p[0], n, np.uint32), "end": _arr(p[1], n, np.uint32), "lable": _arr(p[2], n, np.uint8)}

    def candidates(self):
        nreg = len(self.regions()["start"])
        n, p = self._get("np2o_get_candidates", range(6))
        seq_off = _arr(p[4], n + 1 if nreg else 0, np.uint64)
        return {"roff": _arr(p[0], nreg + 1 if nreg else 0, np.uint64), "order": _arr(p[1], n, np.uint32),
                "kscore": _arr(p[2], n, np.uint16), "kmer": _arr(p[3], n, np.uint64), "seq_off": seq_off,
                "seq": _arr(p[5], int(seq_off[-1]) if len(seq_off) else 0, np.uint8)}

    def dropped(self):
        n, p = self._get("np2o_get_dropped", range(1))
        return _arr(p[0], n, np.uint32)

    def pair_weights(self):
        """(keys a << 32 | b ascending, vals #agree + #differ * (2^32 - 1)) of the dumped iteration (main.rs:953-992)."""
        n, p = self._get("np2o_get_pair_weights", range(2))
        return _arr(p[0], n, np.uint64), _arr(p[1], n, np.int64)

    def consensus(self):
        n, p = self._get("np2o_get_consensus", range(2))
        return _arr(p[0], n, np.uint32), _arr(p[1], n, np.uint8)

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.np2o_job_destroy(self.h)
            self.h = None


def order_exposure(reset=False):
    """Counters of the places where the reference's result depends on FxHashMap iteration order (np2_oracle.h)."""
    v = (C.c_uint64 * 4)()
    lib().np2o_order_exposure(v, int(reset))
    return {"phasing_calls": int(v[0]), "declustered_communities": int(v[1]), "tied_conflicting_pairs": int(v[2]),
            "reads_in_tied_pairs": int(v[3])}


def debug_phase(keys, vals, model=0, use_all_reads=False):
    """Phasing (main.rs:994-1015 + louvain.rs) on pre-summed pair weights -> sorted read orders to drop."""
    keys = np.ascontiguousarray(keys, np.uint64)
    vals = np.ascontiguousarray(vals, np.int64)
    out = np.empty(max(len(keys) * 2 + 1, 1), np.uint32)
    n = lib().np2o_debug_phase(keys.ctypes.data, vals.ctypes.data, len(keys), model, int(use_all_reads), out.ctypes.data, len(out))
    if n < 0:
        raise OracleError(lib().np2o_last_error().decode())
    return out[:n].copy()


def format_fasta(tid, pos, base, uppercase=False, out_pos=False):
    pos = np.ascontiguousarray(pos, np.uint32)
    base = np.ascontiguousarray(base, np.uint8)
    cap = len(base) * (1 if not out_pos else len(tid) + 16) + len(tid) + 64
    out = np.empty(cap, np.uint8)
    n = lib().np2o_format_fasta(tid.encode(), pos.ctypes.data, base.ctypes.data, len(base), int(uppercase), int(out_pos),
                                out.ctypes.data, cap)
    assert n <= cap
    return bytes(out[:n])
