"""Build libsvslam.so (sm_100a only) in-tree with nvcc.  Usage: python build.py [--force]"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
OUT = os.path.join(HERE, "libsvslam.so")
OBJ = os.path.join(HERE, "build")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-fvisibility=hidden,-fopenmp", "--fmad=false"]
# --fmad=false for the image kernels: they must round exactly where OpenCV rounds.  The FP64 solver files may
# contract to FMA (fewer instructions, better accuracy; parity there is a 1e-9 tolerance, not bit-exactness).
FMA_OK = {"ba.cu", "geom.cu", "ba_shard.cu"}


def sources():
    out = []
    for d in (CSRC, HOST):
        if os.path.isdir(d):
            for f in sorted(os.listdir(d)):
                if f.endswith((".cu", ".cpp")):
                    out.append(os.path.join(d, f))
    return out


def _newest_header():
    t = 0
    for d in (CSRC, HOST, os.path.join(HERE, "..", "include")):
        if os.path.isdir(d):
            for f in os.listdir(d):
                if f.endswith((".h", ".cuh", ".hpp")):
                    t = max(t, os.path.getmtime(os.path.join(d, f)))
    return t


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = sources()
    hdr_t = _newest_header()
    objs, jobs = [], []
    for s in srcs:
        o = os.path.join(OBJ, os.path.basename(s) + ".o")
        objs.append(o)
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), hdr_t):
            flags = list(COMMON)
            if os.path.basename(s) in FMA_OK:
                flags.remove("--fmad=false")
            cmd = [NVCC] + ARCH + flags + (["-Xptxas", "-v"] if verbose else []) + ["-x", "cu", "-c", s, "-o", o]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        return cmd, r

    with ThreadPoolExecutor(max_workers=8) as ex:
        for cmd, r in ex.map(run, jobs):
            if verbose or r.returncode != 0:
                sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError("nvcc failed for " + cmd[-3])
    if jobs or not os.path.exists(OUT):
        cmd = [NVCC] + ARCH + ["-shared", "-Xcompiler", "-fopenmp", "-o", OUT] + objs + ["-lgomp"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
