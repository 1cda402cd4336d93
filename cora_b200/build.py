"""Build libcora_b200.so (sm_100a) in-tree with nvcc.  Used by __graft_entry__.build()."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib", "libcora_b200.so")


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build(force=False, verbose=False):
    srcs = [os.path.join(SRC, f) for f in sorted(os.listdir(SRC))]
    srcs.append(os.path.join(HERE, "..", "include", "cora_b200.h"))
    if not force and not _newer(LIB, srcs):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function", "-shared", "--use_fast_math=false",
           "-o", LIB, os.path.join(SRC, "unity.cu"), "-lcudart"]
    cmd = [c for c in cmd if c != "--use_fast_math=false"]
    if os.environ.get("CORA_B200_NVCC_DEFS"):
        cmd[1:1] = os.environ["CORA_B200_NVCC_DEFS"].split()
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libcora_b200.so")
    if verbose:
        sys.stderr.write(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
