"""Build libconcentus_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

Two translation units (decoder half, encoder half) are compiled in parallel and linked into one shared object."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
SRCS = [os.path.join(HERE, "csrc", "opus_capi.cu"), os.path.join(HERE, "csrc", "opus_enc_capi.cu"), os.path.join(HERE, "csrc", "opus_enc_pipe.cu"),
        os.path.join(HERE, "csrc", "opus_repacketizer.cu")]
OBJDIR = os.path.join(HERE, "build")
OUT = os.path.join(HERE, "libconcentus_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
EXTRA = os.environ.get("CB200_DEFS", "").split()   # e.g. "-DCB_ENC_WPB=7 -DCB_PHASE_SYNC=1" (A/B builds only)
if os.environ.get("CB200_OUT"):
    OUT = os.environ["CB200_OUT"]
    OBJDIR = OUT + ".obj"
FLAGS = EXTRA + ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-diag-suppress", "550"]


def _newest_src():
    t = 0.0
    for root in (os.path.join(HERE, "csrc"), os.path.join(HERE, "..", "include")):
        for f in os.listdir(root):
            t = max(t, os.path.getmtime(os.path.join(root, f)))
    return t


def _compile(src, verbose):
    obj = os.path.join(OBJDIR, os.path.basename(src)[:-3] + ".o")
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return obj, r


def build(force=False, verbose=False):
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= _newest_src():
        return OUT
    os.makedirs(OBJDIR, exist_ok=True)
    with ThreadPoolExecutor(len(SRCS)) as ex:
        res = list(ex.map(lambda s: _compile(s, verbose), SRCS))
    for (obj, r), src in zip(res, SRCS):
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("nvcc failed for %s" % src)
        if verbose:
            sys.stderr.write(r.stderr)
    r = subprocess.run([NVCC, "-shared", "-o", OUT] + [o for o, _ in res], capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
