"""Build libffmp_b200.so (hand-written sm_100a CUDA + the C-ABI) in-tree with nvcc.

nvcc cross-compiles without a GPU; the resulting .so travels to the GPU box with the repo snapshot.
Every .cu is compiled to an object in parallel (csrc/build/), then linked: a rebuild after touching one file takes seconds.
"""
import glob
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(CSRC, "libffmp_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-fmad=false",                 # SPEC.md: one rounding per written fp32 operation
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _headers():
    return glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(HERE, "..", "include", "*.h"))


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in sources() + _headers())


def _compile(src: str, force: bool, verbose: bool) -> str:
    obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
    deps = [src] + _headers()
    if not force and os.path.exists(obj) and all(os.path.getmtime(d) <= os.path.getmtime(obj) for d in deps):
        return obj
    nvcc = os.environ.get("NVCC", "nvcc")
    res = subprocess.run([nvcc, *NVCC_FLAGS, "-c", "-o", obj, src], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed compiling " + src)
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    if force or needs_build():
        os.makedirs(OBJ, exist_ok=True)
        with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
            objs = list(ex.map(lambda s: _compile(s, force, verbose), sources()))
        nvcc = os.environ.get("NVCC", "nvcc")
        res = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB, *objs],
                             capture_output=True, text=True)
        if res.returncode != 0:
            print(res.stdout)
            print(res.stderr)
            raise RuntimeError("nvcc failed linking libffmp_b200.so")
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose=True))
