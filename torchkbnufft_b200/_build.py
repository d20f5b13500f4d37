"""In-tree build of ``libb200nufft.so`` (nvcc, sm_100a only).

``python -m torchkbnufft_b200._build`` or ``__graft_entry__.build()``.  nvcc
cross-compiles without a GPU; the built library sits next to the sources
(git-ignored) so it travels with a repository snapshot to the GPU host.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB_NAME = "libb200nufft.so"
LIB_PATH = os.path.join(CSRC, LIB_NAME)
SOURCES = ["b2n_points.cu", "b2n_interp.cu", "b2n_interp_tiled.cu", "b2n_interp_tiled_cl.cu", "b2n_interp_own.cu", "b2n_interp_tiled3d.cu", "b2n_fftops.cu", "b2n_fft.cu", "b2n_peer.cu",
           "b2n_fft_plans_a.cu", "b2n_fft_plans_b.cu", "b2n_fft_plans_c.cu", "b2n_fft_plans_d.cu", "b2n_fft_plans_e.cu", "b2n_fft_plans_f.cu", "b2n_fft_plans_g.cu", "b2n_fft_plans_h.cu"]
HEADERS = ["b2n_common.cuh", "b2n_math.cuh", "b2n_tiling.cuh", "b2n_interp.cuh", "b2n_fft_core.cuh", "b2n_fft_fast.cuh",
           "b2n_fft_args.cuh", "b2n_fft_fast_kernels.cuh", "b2n_tiled_common.cuh", "b2n_peer.cuh", os.path.join("..", "..", "include", "b200nufft.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    # index arithmetic must not be contracted into FMAs on the host side either; device
    # code uses explicit __f*_rn intrinsics where rounding matters (b2n_math.cuh)
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the B200 NUFFT engine needs the CUDA toolkit to build")


def _stamp() -> str:
    h = hashlib.sha256()
    for name in SOURCES + HEADERS:
        with open(os.path.join(CSRC, name), "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu for sm_100a and link the shared library.  Returns its path."""
    stamp_file = os.path.join(CSRC, ".build_stamp")
    stamp = _stamp()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(stamp_file):
        with open(stamp_file) as f:
            if f.read().strip() == stamp:
                return LIB_PATH
    nvcc = _nvcc()
    objdir = os.path.join(CSRC, "build")
    os.makedirs(objdir, exist_ok=True)

    def compile_one(src: str) -> str:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), file=sys.stderr)
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{res.stdout}\n{res.stderr}")
        if verbose:
            print(res.stderr, file=sys.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        objs = list(pool.map(compile_one, SOURCES))
    link = [nvcc, "-shared", "-o", LIB_PATH] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    with open(stamp_file, "w") as f:
        f.write(stamp)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
