"""Case table shared by oracle/make_golden.py (which runs the REFERENCE on these
inputs and stores its outputs in tests/golden/ref_cases.npz) and the tests
(which rebuild the same seeded inputs and compare).  numpy only."""
from __future__ import annotations

import zlib

import numpy as np

# geometries whose module buffers are dumped from the reference (ref_buffers.npz)
BUFFER_GEOMS = [
    dict(im_size=(8,), c64=False, table_oversamp=64),
    dict(im_size=(12, 10), c64=False),  # defaults: grid 2N, J=6, L=1024
    dict(im_size=(12, 10), c64=True),  # c64: scaling_coef from the float32-rounded alpha
    dict(im_size=(6, 7, 5), grid_size=(9, 12, 8), numpoints=(2, 5, 6), table_oversamp=(32, 64, 128), c64=False),
    dict(im_size=(16, 16), grid_size=(20, 24), numpoints=4, n_shift=(3, 5), kbwidth=2.0, order=0.0,
         table_oversamp=128, c64=True),
    dict(im_size=(7,), numpoints=1, table_oversamp=16, c64=False),  # J=1: scaling == 1
    dict(im_size=(9, 8), numpoints=(3, 6), table_oversamp=(7, 16), c64=False),  # odd J, odd L
]

# small seeded cases run through the reference in complex64 and complex128
CASES = {
    # 1-D, non-2x grid
    "d1": dict(im_size=(19,), grid_size=(57,), B=2, C=3, M=61, omega="uniform"),
    # 2-D defaults, random points
    "d2": dict(im_size=(16, 12), B=2, C=3, M=150, omega="uniform"),
    # 2-D, trajectory beyond [-pi, pi] and exact grid hits (rounding ties, wrap-around)
    "d2_edge": dict(im_size=(10, 14), grid_size=(16, 20), B=1, C=2, M=200, omega="edge"),
    # 2-D, mixed neighbours / small odd table, custom n_shift
    "d2_mixed": dict(im_size=(9, 8), grid_size=(13, 16), numpoints=(3, 6), table_oversamp=(7, 16),
                     n_shift=(2, 5), B=1, C=2, M=90, omega="uniform"),
    # 3-D, non-2x grid
    "d3": dict(im_size=(8, 7, 6), grid_size=(12, 11, 10), B=1, C=2, M=120, omega="uniform"),
    # 3-D mixed J incl. J=2
    "d3_mixed": dict(im_size=(6, 7, 5), grid_size=(9, 12, 8), numpoints=(2, 5, 6), table_oversamp=(32, 64, 128),
                     B=1, C=2, M=70, omega="uniform", toep=False),
    # batched trajectories (B, d, M)
    "d2_batched": dict(im_size=(12, 10), B=3, C=2, M=80, omega="batched"),
    # radial 2-D: dense centre (many points per cell)
    "d2_radial": dict(im_size=(16, 16), B=1, C=4, M=24 * 32, omega="radial"),
}


def _seed(name: str) -> int:
    return zlib.crc32(name.encode()) & 0x7FFFFFFF


def _cnormal(rng, shape, cdtype):
    x = rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
    return np.ascontiguousarray(x.astype(cdtype))


def case_inputs(case: dict, cdtype) -> dict:
    """Seeded inputs of a case.  The float64 draws are identical for both
    precisions; complex64 inputs are the rounded complex128 ones."""
    name = [k for k, v in CASES.items() if v is case][0]
    rng = np.random.default_rng(_seed(name))
    rdtype = np.float32 if np.dtype(cdtype) == np.complex64 else np.float64
    N = tuple(case["im_size"])
    K = tuple(case.get("grid_size") or tuple(2 * n for n in N))
    d, B, C, M = len(N), case["B"], case["C"], case["M"]
    kind = case["omega"]
    if kind == "uniform":
        omega = rng.uniform(-np.pi, np.pi, size=(d, M))
    elif kind == "batched":
        omega = rng.uniform(-np.pi, np.pi, size=(B, d, M))
    elif kind == "edge":
        omega = rng.uniform(-1.7 * np.pi, 1.7 * np.pi, size=(d, M))
        for dim in range(d):  # exact multiples of the grid spacing and of half a table step
            gam = 2 * np.pi / K[dim]
            omega[dim, : M // 4] = gam * rng.integers(-K[dim], K[dim], size=M // 4)
            omega[dim, M // 4: M // 2] = gam * (rng.integers(-K[dim], K[dim], size=M // 4) + 0.5 / 1024
                                                * rng.integers(0, 2048, size=M // 4))
        omega[:, 0] = 0.0
        omega[:, 1] = np.pi
        omega[:, 2] = -np.pi
    elif kind == "radial":
        n_read = 32
        n_spokes = M // n_read
        theta = np.arange(n_spokes) * np.pi / n_spokes
        r = np.linspace(-np.pi, np.pi, n_read, endpoint=False)
        omega = np.stack([np.outer(np.sin(theta), r).reshape(-1), np.outer(np.cos(theta), r).reshape(-1)])
    else:
        raise KeyError(kind)
    omega = np.ascontiguousarray(omega.astype(rdtype))
    return dict(
        omega=omega,
        grid=_cnormal(rng, (B, C) + K, cdtype),
        kdata=_cnormal(rng, (B, C, M), cdtype),
        image=_cnormal(rng, (B, 1) + N, cdtype),
        image_multi=_cnormal(rng, (B, C) + N, cdtype),
        smaps=_cnormal(rng, (B if kind == "batched" else 1, C) + N, cdtype),
        weights=np.ascontiguousarray(rng.uniform(0.5, 1.5, size=(B, 1, M) if kind == "batched" else (1, M))
                                     .astype(rdtype)),
        im_size=N,
        grid_size=K,
    )


# ---- sparse-matrix interpolation shim (oracle/make_golden_spmat.py, tests/test_spmat.py) -------
SPMAT_CASES = {
    "s2d_f64": dict(im_size=(8, 8), M=30, dtype="float64"),
    "s2d_f32": dict(im_size=(12, 10), M=41, dtype="float32"),
    "s1d_f64": dict(im_size=(16,), M=25, dtype="float64", numpoints=4),
    "s3d_f64": dict(im_size=(6, 8, 5), M=37, dtype="float64", grid_size=(10, 12, 9), numpoints=(4, 5, 6),
                    n_shift=(1, 3, 2)),
}


def spmat_case_inputs(name):
    """(omega (d, M), image (2, 1, *N), kdata (2, 3, M), smaps (1, 3, *N)) regenerated from the case's seed."""
    import zlib

    import numpy as np
    cfg = SPMAT_CASES[name]
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    real = np.dtype(cfg["dtype"])
    cplx = np.complex64 if real == np.float32 else np.complex128
    N, M, d = cfg["im_size"], cfg["M"], len(cfg["im_size"])
    omega = rng.uniform(-np.pi, np.pi, size=(d, M)).astype(real)

    def cn(shape):
        return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(cplx)

    return omega, cn((2, 1) + tuple(N)), cn((2, 3, M)), cn((1, 3) + tuple(N))
