"""slq_build.py -- compiles csrc/*.cu into libslq_b200.so (in-tree, next to this file) for sm_100a.

nvcc cross-compiles without a GPU, so this runs in the build container; the resulting .so travels
to the GPU box with the repo snapshot (it is git-ignored, not gpurun-ignored).
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libslq_b200.so")
STAMP = os.path.join(HERE, "build", "sources.sha256")
SOURCES = ["quantizer.cu", "layers.cu", "conv_umma.cu", "stem_umma.cu", "eval_tail.cu", "probe.cu", "tail_umma.cu",
           "block_tail.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "-I", os.path.join(ROOT, "include"), "-I", CSRC,
]


def _digest():
    h = hashlib.sha256()
    names = sorted(os.listdir(CSRC)) + ["../../include/slq.h"]
    for n in names:
        p = os.path.join(CSRC, n)
        if os.path.isfile(p):
            h.update(n.encode())
            h.update(open(p, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def needs_build():
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    return open(STAMP).read().strip() != _digest()


def build(force=False, verbose=False, debug=False):
    """Builds libslq_b200.so if the sources changed.  Returns the library path.
    debug=True builds libslq_b200_dbg.so instead: the same kernels with the timeline tracer / wait
    statistics compiled in (-DSLQ_DEBUG_TRACE=1), used by tools/trace_*.py and tools/wait_stats.py."""
    if debug:
        return _build_debug(verbose)
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(HERE, "build", src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd))
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out.decode())
            raise RuntimeError("nvcc failed on %s" % src)
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    subprocess.check_call(cmd)
    with open(STAMP, "w") as f:
        f.write(_digest())
    return LIB


LIB_DBG = os.path.join(HERE, "libslq_b200_dbg.so")


def _build_debug(verbose=False):
    stamp = os.path.join(HERE, "build", "dbg", "sources.sha256")
    if os.path.exists(LIB_DBG) and os.path.exists(stamp) and open(stamp).read().strip() == _digest():
        return LIB_DBG
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(os.path.join(HERE, "build", "dbg"), exist_ok=True)
    procs, objs = [], []
    for src in SOURCES:
        obj = os.path.join(HERE, "build", "dbg", src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + ["-DSLQ_DEBUG_TRACE=1", "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd))
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out.decode())
            raise RuntimeError("nvcc failed on %s" % src)
    subprocess.check_call([nvcc, "-shared", "-o", LIB_DBG] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"])
    with open(stamp, "w") as f:
        f.write(_digest())
    return LIB_DBG


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True, debug="--debug" in sys.argv))
