"""Build the CPU oracle shared libraries (test infrastructure, not product).

    python oracle/build.py            # builds oracle/_build/liboracle_{f32,f64}.so

The reference's rasterizer source (graphdeco-inria/diff-gaussian-rasterization) and
sampler source (princeton-vl/RAFT-Stereo/sampler) are not under /root/reference, so
there is no `oracle/_ref/` build: nothing of the reference compiles from its own
sources (it is pure Python; see DESIGN.md "Oracle").
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build")
SRCS = ["gpsg_oracle.c", "corr_oracle.c", "taichi_splat_oracle.c"]


def _newer(dst, srcs):
    if not os.path.exists(dst):
        return True
    t = os.path.getmtime(dst)
    return any(os.path.getmtime(s) > t for s in srcs)


def build(force=False, verbose=False):
    os.makedirs(OUT, exist_ok=True)
    srcs = [os.path.join(HERE, s) for s in SRCS if os.path.exists(os.path.join(HERE, s))]
    outs = {}
    for tag, defs in (("f32", []), ("f64", ["-DORACLE_F64"])):
        dst = os.path.join(OUT, f"liboracle_{tag}.so")
        outs[tag] = dst
        if not force and not _newer(dst, srcs + [os.path.abspath(__file__)]):
            continue
        # -ffp-contract=off: no FMA contraction, so fp32 results follow the written op order
        cmd = ["gcc", "-O2", "-fPIC", "-shared", "-std=c11", "-fopenmp", "-ffp-contract=off", "-fno-fast-math",
               "-Wall", "-Wno-unused-function"] + defs + srcs + ["-lm", "-o", dst]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return outs


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
