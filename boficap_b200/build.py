"""Builds libbofi_b200.so (sm_100a only) in-tree with nvcc.  `python -m boficap_b200.build`."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "engine.cu")
OUT = os.path.join(HERE, "lib", "libbofi_b200.so")
DEPS = [os.path.join(HERE, "csrc", f) for f in ("engine.cu", "common.cuh", "gemm_simt.cuh", "gemm_tc.cuh", "gemm_ln_tc.cuh", "gemm_tc2.cuh", "gemm_tc2_ln.cuh", "kernels.cuh", "attention_mma.cuh", "train_kernels.cuh", "attention_bwd_mma.cuh", "bound_loop.cuh", "train.inl", "train_abi.inl")] + \
       [os.path.join(os.path.dirname(HERE), "include", "bofi_b200.h")]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in DEPS)


PROF_OUT = os.path.join(HERE, "lib", "libbofi_b200_prof.so")


def build(force=False, verbose=False, prof=False, out=None):
    """prof=True builds lib/libbofi_b200_prof.so with the GEMM stall counters compiled in (-DBOFI_GEMM_PROF);
    it is only ever loaded when BOFI_LIB_PATH points at it (tools/gemm_stalls.py).  `out` (or BOFI_BUILD_OUT): build to
    another file, e.g. a candidate that is moved over the shipped library only once it is known to be good."""
    out = out or os.environ.get("BOFI_BUILD_OUT")
    if not prof and not force and not out and not needs_build():
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    out = out or (PROF_OUT if prof else OUT)
    cmd = [nvcc_path(), "-std=c++17", "-O3", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
           "-shared", "-Xcompiler", "-fPIC"] + (["-DBOFI_GEMM_PROF"] if prof else []) + ["-o", out, SRC]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed")
    if verbose:
        print(r.stderr)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, prof="--prof" in sys.argv))
