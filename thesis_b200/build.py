"""Compile the CUDA kernels + C ABI into thesis_b200/_librbpf.so (sm_100a only).

nvcc cross-compiles without a GPU, so this runs in the build container; the
resulting .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "_librbpf.so")
SOURCES = ["ctx.cu", "k_motion.cu", "k_match.cu", "k_weight.cu", "k_raycast.cu", "k_resample.cu", "k_migrate.cu", "k_misc.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",                       # cell indices replay the reference's float64 expressions
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off",
    "--shared",
]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "rbpf_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, defines=(), out=None):
    """defines/out: tuning variants, e.g. build(defines=["RC_INFLIGHT=2"], out="_librbpf_if2.so")."""
    target = os.path.join(HERE, out) if out else OUT
    if not force and not out and not needs_build():
        return OUT
    cmd = ([NVCC] + FLAGS + ["-D" + d for d in defines] + (["-Xptxas", "-v"] if verbose else []) +
           [os.path.join(CSRC, s) for s in SOURCES] + ["-o", target])
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    subprocess.check_call(cmd, env=env)
    return target


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
