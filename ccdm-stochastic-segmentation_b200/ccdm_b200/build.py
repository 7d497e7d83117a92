"""Builds ``csrc/*.cu`` into ``libccdm_b200.so`` (in-tree, sm_100a only).

nvcc cross-compiles without a GPU; the built library is git-ignored but travels to
the GPU box with the repo snapshot.  ``python -m ccdm_b200.build`` or
``__graft_entry__.build()``.
"""
import glob
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libccdm_b200.so")
STAMP = os.path.join(HERE, "libccdm_b200.stamp")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "--expt-relaxed-constexpr",
         "-Xcompiler", "-fPIC,-fvisibility=default", "-Xptxas", "-v"] + (["-DCCDM_TRACE"] if os.environ.get("CCDM_TRACE") == "1" else []) + (["-DCCDM_ABLATE"] if os.environ.get("CCDM_ABLATE_BUILD") == "1" else []) + (
             ["-DCCDM_SILU_EXPERIMENTS"] if os.environ.get("CCDM_SILU_EXPERIMENTS") == "1" else [])


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _digest():
    h = hashlib.sha256()
    for f in _sources() + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [
            os.path.join(HERE, "..", "..", "include", "ccdm_b200.h")]:
        with open(f, "rb") as fh:
            h.update(f.encode() + b"\0" + fh.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == digest:
        return LIB
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for src in _sources():
        obj = os.path.join(HERE, "build", os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [NVCC, *FLAGS, "-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f"== {os.path.basename(src)}\n{out}")
        failed |= p.returncode != 0
    with open(os.path.join(HERE, "build", "ptxas.log"), "w") as fh:
        fh.write("\n".join(log))
    if failed or verbose:
        sys.stderr.write("\n".join(log))
    if failed:
        raise RuntimeError("nvcc failed; see output above")
    subprocess.check_call([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB, *objs, "-lcudart_static",
                           "-lpthread", "-ldl", "-lrt"])
    with open(STAMP, "w") as fh:
        fh.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
