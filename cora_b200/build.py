"""Build libcora_b200.so (sm_100a) in-tree with nvcc.  Used by __graft_entry__.build().

The library is a handful of objects compiled in parallel: `unity.cu` (host logic, C-ABI, the small kernels)
and one object per persistent-kernel instantiation (`pk_instance.cu` with -DPK_D=<d> -DPK_R=<rank>):
rank 0 is the any-rank tile-pipeline kernel, rank > 0 the rank-specialised streaming kernel.
"""
import concurrent.futures
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(LIBDIR, "obj")
LIB = os.path.join(LIBDIR, "libcora_b200.so")

# (d, rank) pairs the persistent kernels are compiled for; ranks outside the list run on the (d, 0) kernels.
# SE(2): refine at rank 2, staircase 3..6; SE(3): refine at rank 3, staircase 4..8 (BASELINE cfg3: 5 -> 7).
PK_LIST = [(2, 0), (3, 0), (2, 2), (2, 3), (2, 4), (2, 5), (2, 6), (3, 3), (3, 4), (3, 5), (3, 6), (3, 7), (3, 8)]


def pk_list():
    env = os.environ.get("CORA_B200_PK")  # development: "3:5,3:0" compiles only these pairs
    if env:
        return [tuple(int(x) for x in item.split(":")) for item in env.split(",")]
    return PK_LIST


def _nvcc_base():
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function"]
    if os.environ.get("CORA_B200_NVCC_DEFS"):
        cmd += os.environ["CORA_B200_NVCC_DEFS"].split()
    return cmd


def _deps_stamp(extra):
    h = hashlib.sha1()
    files = [os.path.join(SRC, f) for f in sorted(os.listdir(SRC)) if f.endswith((".cu", ".cuh", ".hpp", ".h"))]
    files.append(os.path.join(HERE, "..", "include", "cora_b200.h"))
    files.append(os.path.abspath(__file__))
    for f in files:
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(extra.encode())
    return h.hexdigest()


def _compile(job):
    name, src, defs, verbose = job
    obj = os.path.join(OBJDIR, name + ".o")
    stamp_file = obj + ".stamp"
    stamp = _deps_stamp(" ".join(defs) + os.environ.get("CORA_B200_NVCC_DEFS", ""))
    if os.path.exists(obj) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return obj, ""
    cmd = _nvcc_base() + defs + (["-Xptxas=-v"] if verbose else []) + ["-c", src, "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed on %s:\n%s%s" % (name, res.stdout, res.stderr))
    with open(stamp_file, "w") as fh:
        fh.write(stamp)
    return obj, res.stderr


def build(force=False, verbose=False):
    pairs = pk_list()
    # one stamp for the whole library: sources + instantiation list + extra flags.  A matching stamp beside the
    # .so means it is current even when the objects are absent (they do not travel to the GPU box).
    lib_stamp = _deps_stamp(repr(pairs) + os.environ.get("CORA_B200_NVCC_DEFS", ""))
    stamp_file = LIB + ".stamp"
    if not force and not verbose and os.path.exists(LIB) and os.path.exists(stamp_file) \
            and open(stamp_file).read() == lib_stamp:
        return LIB
    os.makedirs(OBJDIR, exist_ok=True)
    # the compiled pairs as an X-macro list for pk_registry.cuh (a generated header: nvcc splits -D values at commas)
    xlist = "#define CORA_PK_LIST " + " ".join("X(%d, %d)" % p for p in pairs) + "\n"
    gen = os.path.join(OBJDIR, "pk_list.gen.h")
    if not os.path.exists(gen) or open(gen).read() != xlist:
        with open(gen, "w") as fh:
            fh.write(xlist)
    jobs = [("unity", os.path.join(SRC, "unity.cu"), ["-I" + OBJDIR, "-DCORA_PK_LIST_HEADER", "-DPKS=" + "_".join("%d%d" % p for p in pairs)], verbose)]
    for d, r in pairs:
        jobs.append(("pk_%d_%d" % (d, r), os.path.join(SRC, "pk_instance.cu"), ["-DPK_D=%d" % d, "-DPK_R=%d" % r], verbose))
    if force:
        for name, *_ in jobs:
            for suffix in (".o", ".o.stamp"):
                try:
                    os.remove(os.path.join(OBJDIR, name + suffix))
                except FileNotFoundError:
                    pass
    workers = int(os.environ.get("CORA_B200_BUILD_JOBS", str(os.cpu_count() or 4)))
    with concurrent.futures.ThreadPoolExecutor(max_workers=workers) as ex:
        results = list(ex.map(_compile, jobs))
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            sys.stderr.write(log)
    cmd = _nvcc_base() + ["-shared", "-o", LIB] + objs + ["-lcudart"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed linking libcora_b200.so")
    with open(stamp_file, "w") as fh:
        fh.write(lib_stamp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
