"""In-tree build of the CUDA library (liblnaphylodyn_b200.so) for sm_100a with nvcc.

The built .so sits next to this file (git-ignored, but it travels to the GPU box with the repo
snapshot).  `python -m lnaphylodyn_b200.build` or `__graft_entry__.build()` runs this.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "liblnaphylodyn_b200.so")
SOURCES = ["lna_capi.cu"]
DEPS = ["lna_capi.cu", "lna_capi_eslice.inl", "lna_comm.inl", "lna_device.cuh", "lna_kernels.cuh", "lna_eslice.cuh", "lna_split.cuh",
        "lna_sirs.cuh", "lna_legacy.cuh", "lna_thin.cuh", "lna_async.cuh", "lna_seg.cuh", "lna_capi_legacy.inl",
        os.path.join("..", "..", "include", "lnaphylodyn_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--shared", "-Xcompiler", "-fPIC", "-ldl",
    "-Xptxas", "-v",
]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA toolkit is required to build lnaphylodyn_b200")


def stale():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force=False, verbose=False):
    if not force and not stale():
        return SO
    try:
        nvcc = nvcc_path()
    except RuntimeError:
        if os.path.exists(SO) and not force:  # prebuilt library shipped with the tree, no toolkit on this box
            return SO
        raise
    cmd = [nvcc] + NVCC_FLAGS + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", SO]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if res.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building liblnaphylodyn_b200.so")
    if verbose:
        print(log)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
