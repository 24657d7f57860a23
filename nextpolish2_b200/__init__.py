"""nextpolish2_b200 — B200-native (sm_100a) implementation of the NextPolish2 per-contig polish path.

Python here is only the host-side mirror of the C ABI in include/np2gpu.h (ctypes): the product is
libnp2gpu.so (hand-written CUDA kernels + C++ host phases).  There is no CPU fallback: loading fails
loudly when the library is missing and every call fails when no CUDA device is present.
"""
from .api import (Context, Table, Job, Opts, Np2Error, PinnedBuffer, SecondarySeqs, Counter, polish_contig, format_fasta, lib_path,
                  load_library, bgzf_members, bgzf_inflate)  # noqa: F401

__all__ = ["Context", "Table", "Job", "Opts", "Np2Error", "PinnedBuffer", "SecondarySeqs", "Counter", "polish_contig", "format_fasta", "lib_path", "load_library", "bgzf_members", "bgzf_inflate"]
