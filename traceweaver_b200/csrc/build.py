"""Build libtw_b200.so (hand-written sm_100a CUDA + C ABI) in-tree with nvcc.

    python -m traceweaver_b200.csrc.build [--force] [--verbose]

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT = os.path.join(PKG, "libtw_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "-Xptxas", "-v",
         "-Wno-deprecated-gpu-targets"]


def sources():
    return sorted(glob.glob(os.path.join(HERE, "*.cu")))


def stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = sources() + glob.glob(os.path.join(HERE, "*.cuh")) + \
        [os.path.join(PKG, "..", "include", "traceweaver_b200.h"), os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not stale():
        return OUT
    cmd = [NVCC] + FLAGS + ["-o", OUT] + sources()
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log = os.path.join(HERE, "build.log")
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + res.stdout)
    if verbose or res.returncode != 0:
        print(res.stdout)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed ({res.returncode}); see {log}")
    return OUT


def build_profiling():
    """libtw_b200_prof.so: the same sources with -DTW_PROFILE_PHASES (clock64 phase timers inside the
    scoring / stitch / refit kernels; scripts/*_phase_profile.py read them).  Never used by the product."""
    out = os.path.join(PKG, "libtw_b200_prof.so")
    cmd = [NVCC] + FLAGS + ["-DTW_PROFILE_PHASES", "-o", out] + sources()
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        print(res.stdout)
        raise RuntimeError("nvcc failed (profiling build)")
    return out


def build_variant(name, defines):
    """libtw_b200_<name>.so with extra -D flags: measurement variants (scripts/*), never used by the product."""
    out = os.path.join(PKG, f"libtw_b200_{name}.so")
    cmd = [NVCC] + FLAGS + [f"-D{d}" for d in defines] + ["-o", out] + sources()
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        print(res.stdout)
        raise RuntimeError(f"nvcc failed (variant {name})")
    return out


if __name__ == "__main__":
    if "--prof" in sys.argv:
        print(build_profiling())
    else:
        print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
