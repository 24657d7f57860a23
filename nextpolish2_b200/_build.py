"""Build recipe for the in-tree native libraries (used by __graft_entry__.build())."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libnp2gpu.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "-diag-suppress", "177", "-Xcompiler", "-Wno-deprecated-declarations", "-shared"]
SOURCES = ["np2_kernels.cu", "np2_geno.cu", "np2_regions.cu", "np2_count.cu", "np2_inflate.cu", "np2_api.cu", "np2_host.cpp", "np2_secondary.cpp", "np2_phase.cpp"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_gpu_lib(force=False, verbose=False):
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "np2gpu.h")]
    if force or _stale(LIB, deps):
        cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + \
              [os.path.join(CSRC, s) for s in SOURCES]
        subprocess.check_call(cmd)
    return LIB


CLI = os.path.join(HERE, "nextPolish2")


def build_cli(force=False):
    """The nextPolish2-compatible command line (C++ host + libnp2gpu)."""
    src = os.path.join(CSRC, "np2_cli.cpp")
    if force or _stale(CLI, [src, LIB]):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", CLI, src, "-L" + HERE, "-lnp2gpu", "-lz", "-lpthread",
                               "-Wl,-rpath,$ORIGIN", "-Wl,-rpath-link," + "/usr/local/cuda/lib64"])
    return CLI
