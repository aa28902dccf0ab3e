"""In-tree build of the native libraries (no JIT cache: the .so files travel with the repo snapshot).

  librustfhe_b200.so : CUDA kernels (sm_100a) + C ABI (include/tfhe_b200.h) + host keygen   -- the product
  libhostemul.so     : CPU execution of the kernels' per-lane arithmetic -- CPU tests only, never a fallback
  examples/homnand_bench : the reference's homnand-bench example through the C++ host side (include/tfhe_b200.hpp)
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "librustfhe_b200.so")
EMUL = os.path.join(CSRC, "libhostemul.so")
ROOT = os.path.dirname(HERE)
EXAMPLE = os.path.join(ROOT, "examples", "homnand_bench")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC,-O3,-pthread",
              "-shared", "-cudart", "static"]


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def sources():
    return [os.path.join(CSRC, f) for f in ("engine.cu", "blind_rotate.cuh", "keyswitch.cuh", "aux_kernels.cuh", "hostkeys.cpp", "wire.cpp",
                                            "ntt32.cuh", "cmux_steps.cuh", "ntt_tables.h", "tfhe_rng.cuh")] + [
        os.path.join(HERE, "..", "include", "tfhe_b200.h")]


def build(force=False, verbose=False):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if force or not _newer(LIB, sources()):
        if not os.path.exists(nvcc):
            if os.path.exists(LIB):
                return LIB  # GPU box without a toolkit: use the prebuilt library that travelled with the snapshot
            raise RuntimeError("nvcc not found and librustfhe_b200.so is not built")
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [
            os.path.join(CSRC, "engine.cu"), os.path.join(CSRC, "hostkeys.cpp"), os.path.join(CSRC, "wire.cpp"), "-o", LIB]
        subprocess.check_call(cmd)
    emul_src = [os.path.join(CSRC, f) for f in ("host_emul.cpp", "ntt32.cuh", "cmux_steps.cuh", "ntt_tables.h")]
    if force or not _newer(EMUL, emul_src):
        gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([gxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-I/usr/local/cuda/include",
                               os.path.join(CSRC, "host_emul.cpp"), "-o", EMUL])
    ex_src = [EXAMPLE + ".cpp", os.path.join(ROOT, "include", "tfhe_b200.hpp"), os.path.join(ROOT, "include", "tfhe_b200.h"), LIB]
    if force or not _newer(EXAMPLE, ex_src):
        gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([gxx, "-O2", "-std=c++17", "-Wall", "-I" + os.path.join(ROOT, "include"), EXAMPLE + ".cpp", "-L" + HERE,
                               "-lrustfhe_b200", "-Wl,-rpath,$ORIGIN/../rustfhe_b200", "-o", EXAMPLE])
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
