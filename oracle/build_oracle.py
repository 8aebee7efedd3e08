"""Build recipe for the oracle (TEST INFRASTRUCTURE).

  python -m oracle.build_oracle          # C restatement -> oracle/_build/liboracle.so
  python -m oracle.build_oracle --ref    # + the reference's own lib/src CUDA kernels -> oracle/_ref/

`--ref` compiles the four reference .cu files *where they lie* under /root/reference/lib/src
(never copied into this repo) together with oracle/ref_shim.cu, a C-ABI shim written here that
forwards to the reference's `*_kernel_launcher_fast` entry points.  Output goes only to
oracle/_ref/ (git-ignored, NOT gpurun-ignored, so the .so travels to the GPU box).  The reference's
.cpp wrappers are not compiled: they include <THC/THC.h>, which torch 2.11 no longer ships
(SURVEY.md 8c); the kernels and launchers themselves are used unmodified.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")
REF_OUT = os.path.join(HERE, "_ref")
REF_SRC = "/root/reference/lib/src"
REF_CU = ["ball_query_gpu.cu", "group_points_gpu.cu", "interpolate_gpu.cu", "sampling_gpu.cu"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_c(verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    src = os.path.join(HERE, "pointops_oracle.c")
    out = os.path.join(BUILD, "liboracle.so")
    if _newer(out, [src]):
        cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", out, src, "-lm"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return out


def build_ref(verbose=False):
    """nvcc the reference's own kernels (sm_100a) + our shim. Returns path or None if /root/reference is absent."""
    out = os.path.join(REF_OUT, "libpointnet2_ref.so")
    if not os.path.isdir(REF_SRC):
        return out if os.path.exists(out) else None
    os.makedirs(REF_OUT, exist_ok=True)
    import torch  # headers only: the reference *_gpu.h files include <torch/serialize/tensor.h>

    ti = os.path.join(os.path.dirname(torch.__file__), "include")
    inc = ["-I" + ti, "-I" + os.path.join(ti, "torch", "csrc", "api", "include"), "-I" + REF_SRC]
    arch = ["-gencode", "arch=compute_100a,code=sm_100a"]
    shim = os.path.join(HERE, "ref_shim.cu")
    srcs = [os.path.join(REF_SRC, f) for f in REF_CU] + [shim]
    objs = [os.path.join(REF_OUT, os.path.basename(s) + ".o") for s in srcs]

    def cc(pair):
        s, o = pair
        if _newer(o, [s]):
            # -O2 and default -fmad: the reference's own flags (lib/setup.py:18-19)
            cmd = ["nvcc", "-O2"] + arch + inc + ["-Xcompiler", "-fPIC", "-c", s, "-o", o]
            if verbose:
                print(" ".join(cmd))
            subprocess.check_call(cmd)

    with ThreadPoolExecutor(max_workers=5) as ex:
        list(ex.map(cc, zip(srcs, objs)))
    if _newer(out, objs):
        cmd = ["nvcc", "-shared"] + arch + ["-o", out] + objs
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    print(build_c(verbose=True))
    if "--ref" in sys.argv:
        print(build_ref(verbose=True))
