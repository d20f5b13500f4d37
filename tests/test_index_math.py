"""The engine's coordinate/index arithmetic (csrc/b2n_math.cuh), compiled for the
host, against the reference's integer indices (bit-exact) -- no GPU needed.  The same
header is what the CUDA kernels include; the device build swaps the plain operators
for the correctly-rounded __f*_rn intrinsics."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import kbnufft_oracle as orc
from conftest import ROOT
from golden_cases import CASES, case_inputs
from torchkbnufft_b200 import workloads


@pytest.fixture(scope="module")
def hostlib(tmp_path_factory):
    out = tmp_path_factory.mktemp("hostmath") / "libindexmath.so"
    src = os.path.join(ROOT, "tests", "index_math_host.cpp")
    cmd = ["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
           "-I/usr/local/cuda/include", src, "-o", str(out)]
    subprocess.run(cmd, check=True)
    return ctypes.CDLL(str(out))


def run_host(hostlib, omega, K, J, L, n_shift):
    d, M = omega.shape
    W = int(np.prod(J))
    arr_ind = np.empty((W, M), np.int64)
    tab_idx = np.empty((W, d, M), np.int64)
    parg = np.empty(M, omega.dtype)
    K, J, L = (np.ascontiguousarray(np.asarray(v, np.int64)) for v in (K, J, L))
    ns = np.ascontiguousarray(np.asarray(n_shift, omega.dtype))
    fn = hostlib.host_indices_f32 if omega.dtype == np.float32 else hostlib.host_indices_f64
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    fn(ctypes.c_int(d), ctypes.c_int64(M), p(omega), p(K), p(J), p(L), p(arr_ind), p(tab_idx), p(parg), p(ns))
    return arr_ind, tab_idx, parg


@pytest.mark.parametrize("prec", ["c64", "c128"])
@pytest.mark.parametrize("name", [n for n, c in CASES.items() if c["omega"] != "batched"])
def test_host_index_math_equals_reference(name, prec, hostlib, ref_cases):
    case = CASES[name]
    inp = case_inputs(case, np.complex64 if prec == "c64" else np.complex128)
    d = len(case["im_size"])
    J = case.get("numpoints", 6)
    J = [J] * d if isinstance(J, int) else list(J)
    L = case.get("table_oversamp", 1024)
    L = [L] * d if isinstance(L, int) else list(L)
    arr_ind, tab_idx, _ = run_host(hostlib, inp["omega"], inp["grid_size"], J, L, [0] * d)
    assert np.array_equal(arr_ind, ref_cases[f"{name}_{prec}_arr_ind"])
    o_arr, o_tab = orc.calc_coef_and_indices(inp["omega"], inp["grid_size"], J, L)
    assert np.array_equal(tab_idx, o_tab) and np.array_equal(arr_ind, o_arr)


def test_host_index_math_cfg1_full_size(hostlib, ref_cfg1):
    wl = workloads.WORKLOADS["cfg1"]
    omega = wl.trajectory(np.float32)
    arr_ind, tab_idx, _ = run_host(hostlib, omega, wl.grid_size, [6, 6], [1024, 1024], [128, 128])
    step = int(ref_cfg1["step"])
    assert np.array_equal(arr_ind[:, ::step], ref_cfg1["arr_ind_sub"])
    assert tab_idx.min() >= 0 and tab_idx.max() <= 6 * 1024


def test_host_index_math_adversarial_float32(hostlib):
    """Points on / next to cell boundaries and table-step ties, odd grid sizes (where
    (1/K)*2pi != 2pi/K in float32), negative and out-of-range coordinates."""
    rng = np.random.default_rng(7)
    for K in (19, 25, 57, 640, 511):
        gam = np.float32(np.float32(1.0) / np.float32(K)) * np.float32(2 * np.pi)
        cells = rng.integers(-2 * K, 2 * K, size=4000).astype(np.float64)
        frac = rng.integers(0, 2048, size=4000) / 2048.0  # half table steps -> rint ties
        om = ((cells + frac) * float(gam)).astype(np.float32)
        om = np.concatenate([om, np.nextafter(om, np.float32(np.inf)), np.nextafter(om, np.float32(-np.inf))])
        omega = np.ascontiguousarray(om[None, :])
        arr_ind, tab_idx, _ = run_host(hostlib, omega, [K], [6], [1024], [K // 4])
        o_arr, o_tab = orc.calc_coef_and_indices(omega, [K], [6], [1024])
        assert np.array_equal(arr_ind, o_arr) and np.array_equal(tab_idx, o_tab)
