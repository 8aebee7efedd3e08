"""Build libcmflow_b200.so in-tree with nvcc for sm_100a (no torch headers, no JIT cache).

  python -m cmflow_b200.build [-v] [--force] [--ptxas]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "libcmflow_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(os.path.dirname(HERE), "include", "cmflow_b200.h"))
    return hdrs


def build(verbose=False, force=False, ptxas_v=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = sources()
    hdr_t = max(os.path.getmtime(h) for h in _deps())
    objs = [os.path.join(OBJ, os.path.basename(s)[:-3] + ".o") for s in srcs]

    def cc(pair):
        s, o = pair
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), hdr_t):
            cmd = ["nvcc"] + ARCH + FLAGS + (["-Xptxas", "-v"] if ptxas_v else []) + ["-c", s, "-o", o]
            if verbose:
                print(" ".join(cmd), flush=True)
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0 or ptxas_v or (verbose and (r.stdout or r.stderr)):
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError("nvcc failed for " + s)
            return True
        return False

    with ThreadPoolExecutor(max_workers=8) as ex:
        rebuilt = list(ex.map(cc, zip(srcs, objs)))
    if force or any(rebuilt) or not os.path.exists(LIB):
        cmd = ["nvcc", "-shared"] + ARCH + ["-o", LIB] + objs
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return LIB


def build_variant(tag, defines, files):
    """Developer A/B builds: libcmflow_b200_<tag>.so with `files` (basenames under csrc/) recompiled under extra -D defines, every other
    object shared with the main build.  Select it at run time with CMF_LIB=<path> (see _lib.py)."""
    build()
    vobj = os.path.join(HERE, "_obj_" + tag)
    os.makedirs(vobj, exist_ok=True)
    objs = []
    for s in sources():
        base = os.path.basename(s)
        if base in files:
            o = os.path.join(vobj, base[:-3] + ".o")
            subprocess.check_call(["nvcc"] + ARCH + FLAGS + ["-D" + d for d in defines] + ["-c", s, "-o", o])
        else:
            o = os.path.join(OBJ, base[:-3] + ".o")
        objs.append(o)
    out = os.path.join(HERE, "libcmflow_b200_%s.so" % tag)
    subprocess.check_call(["nvcc", "-shared"] + ARCH + ["-o", out] + objs)
    return out


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="--force" in sys.argv, ptxas_v="--ptxas" in sys.argv))
