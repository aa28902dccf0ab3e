"""In-tree build of the native libraries (no JIT cache: the .so files travel with the repo snapshot).

  librustfhe_b200.so : CUDA kernels (sm_100a) + C ABI (include/tfhe_b200.h) + host keygen   -- the product
  libhostemul.so     : CPU execution of the kernels' per-lane arithmetic -- CPU tests only, never a fallback
  examples/homnand_bench : the reference's homnand-bench example through the C++ host side (include/tfhe_b200.hpp)

Staleness is decided by a content hash of the sources stored beside each artefact (mtimes do not survive a checkout or
a snapshot push).  Every build writes to a temporary file and os.replace()s it under an exclusive file lock, so that several
ranks importing the package at once (bench.py runs one process per GPU) never see a half-written library.  lib() (the
import path) only ever builds the product library; the emulator and the example are built by the tests that use them.
"""
import fcntl
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "librustfhe_b200.so")
EMUL = os.path.join(CSRC, "libhostemul.so")
ROOT = os.path.dirname(HERE)
EXAMPLE = os.path.join(ROOT, "examples", "homnand_bench")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC,-O3,-pthread",
              "-shared", "-cudart", "static"]
NCCL_LIBS = ["-lnccl"]   # the multi-GPU group API replicates the keys with ncclBroadcast inside the library


def _digest(paths, extra=""):
    h = hashlib.sha256(extra.encode())
    for p in paths:
        h.update(os.path.basename(p).encode())
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _stale(target, paths, extra=""):
    stamp = target + ".srchash"
    if not os.path.exists(target) or not os.path.exists(stamp):
        return True
    with open(stamp) as f:
        return f.read().strip() != _digest(paths, extra)


def _locked_build(target, paths, make, extra="", force=False):
    """make(tmp_path) must produce the artefact at tmp_path; it is moved into place atomically under a lock."""
    lock = target + ".lock"
    with open(lock, "w") as lk:
        fcntl.flock(lk, fcntl.LOCK_EX)
        try:
            if not force and not _stale(target, paths, extra):
                return target   # another process built it while we waited
            tmp = f"{target}.tmp.{os.getpid()}"
            try:
                make(tmp)
                os.replace(tmp, target)
            finally:
                if os.path.exists(tmp):
                    os.unlink(tmp)
            with open(target + ".srchash.tmp", "w") as f:
                f.write(_digest(paths, extra))
            os.replace(target + ".srchash.tmp", target + ".srchash")
        finally:
            fcntl.flock(lk, fcntl.LOCK_UN)
    return target


def sources():
    names = ("engine.cu", "group.cu", "blind_rotate.cuh", "blind_rotate_t2.cuh", "t2_steps.cuh", "blind_rotate_f64.cuh", "blind_rotate_f64t.cuh", "blind_rotate_f64l2.cuh", "blind_rotate_f64w2.cuh", "fft64.cuh", "fft64_tables.h", "keyswitch.cuh", "aux_kernels.cuh", "hostkeys.cpp",
             "wire.cpp", "csprng.hpp", "ntt32.cuh", "cmux_steps.cuh", "ntt_tables.h", "tfhe_rng.cuh")
    return [os.path.join(CSRC, f) for f in names if os.path.exists(os.path.join(CSRC, f))] + [os.path.join(ROOT, "include", "tfhe_b200.h")]


def _gxx():
    return "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def build(force=False, verbose=False):
    """The product library.  Called on import (rustfhe_b200._capi.lib): builds only when the sources changed."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    src = sources()
    flags = " ".join(NVCC_FLAGS + NCCL_LIBS)
    if not force and not _stale(LIB, src, flags):
        return LIB
    if not os.path.exists(nvcc):
        if os.path.exists(LIB):
            print("rustfhe_b200: WARNING: librustfhe_b200.so does not match its sources and there is no nvcc to rebuild it; "
                  "using the prebuilt library", file=sys.stderr)
            return LIB
        raise RuntimeError("nvcc not found and librustfhe_b200.so is not built")

    def make(tmp):
        cus = [os.path.join(CSRC, f) for f in ("engine.cu", "group.cu", "hostkeys.cpp", "wire.cpp") if os.path.exists(os.path.join(CSRC, f))]
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + cus + NCCL_LIBS + ["-o", tmp]
        subprocess.check_call(cmd)
    return _locked_build(LIB, src, make, flags, force)


def build_emul(force=False):
    """CPU emulation of the kernels' per-lane arithmetic (tests/test_host_logic.py only)."""
    src = [os.path.join(CSRC, f) for f in ("host_emul.cpp", "ntt32.cuh", "cmux_steps.cuh", "t2_steps.cuh", "fft64.cuh", "fft64_tables.h", "ntt_tables.h", "tfhe_rng.cuh")]
    if not force and not _stale(EMUL, src):
        return EMUL

    def make(tmp):
        subprocess.check_call([_gxx(), "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-I/usr/local/cuda/include",
                               os.path.join(CSRC, "host_emul.cpp"), "-o", tmp])
    return _locked_build(EMUL, src, make, "", force)


def build_example(force=False):
    """examples/homnand_bench: the reference's example on the C++ host side, linked against the product library."""
    build()
    src = [EXAMPLE + ".cpp", os.path.join(ROOT, "include", "tfhe_b200.hpp"), os.path.join(ROOT, "include", "tfhe_b200.h")] + sources()
    if not force and not _stale(EXAMPLE, src):
        return EXAMPLE

    def make(tmp):
        subprocess.check_call([_gxx(), "-O2", "-std=c++17", "-Wall", "-I" + os.path.join(ROOT, "include"), EXAMPLE + ".cpp", "-L" + HERE,
                               "-lrustfhe_b200", "-Wl,-rpath,$ORIGIN/../rustfhe_b200", "-o", tmp])
    return _locked_build(EXAMPLE, src, make, "", force)


def build_all(force=False, verbose=False):
    build(force, verbose)
    build_emul(force)
    build_example(force)
    return LIB


if __name__ == "__main__":
    print(build_all(force=True, verbose=True))
