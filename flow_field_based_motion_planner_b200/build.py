"""Build libffmp_b200.so (hand-written sm_100a CUDA + the C-ABI) in-tree with nvcc.

nvcc cross-compiles without a GPU; the resulting .so travels to the GPU box with the repo snapshot.
"""
import glob
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libffmp_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-fmad=false",                 # SPEC.md: one rounding per written fp32 operation
    "-Xcompiler", "-fPIC", "-shared",
    "-Xptxas", "-v",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if force or needs_build():
        nvcc = os.environ.get("NVCC", "nvcc")
        cmd = [nvcc, *NVCC_FLAGS, "-o", LIB, *sources()]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or res.returncode != 0:
            print(res.stdout)
            print(res.stderr)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed building libffmp_b200.so")
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
