"""The engine's mixed-radix Stockham FFT building blocks (csrc/b2n_fft_core.cuh), compiled for
the host, against numpy: every radix (2,3,4,5,7,8,11,13,16), forward and unnormalised inverse,
the sizes of the BASELINE configs, and rejection of sizes with a prime factor > 13."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def hostfft(tmp_path_factory):
    out = tmp_path_factory.mktemp("hostfft") / "libfftcore.so"
    src = os.path.join(ROOT, "tests", "fft_core_host.cpp")
    cmd = ["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I/usr/local/cuda/include", src, "-o", str(out)]
    subprocess.run(cmd, check=True)
    return ctypes.CDLL(str(out))


def run(lib, x, inverse):
    n = x.shape[0]
    out = np.empty(n, np.complex64)
    radix = (ctypes.c_int * 16)()
    stages = lib.host_fft(n, int(inverse), x.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p), radix)
    return stages, out, list(radix[: max(stages, 0)])


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 6, 7, 8, 11, 13, 16, 24, 25, 26, 28, 30, 35, 40, 64, 112,
                               128, 256, 384, 512, 640, 768, 1000, 1024, 1280, 2048, 4096])
def test_stockham_matches_numpy(n, hostfft):
    rng = np.random.default_rng(n)
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    for inverse in (False, True):
        stages, got, radix = run(hostfft, x, inverse)
        assert stages >= 0 and int(np.prod(radix)) == n if n > 1 else True
        want = np.fft.ifft(x.astype(np.complex128)) * n if inverse else np.fft.fft(x.astype(np.complex128))
        err = np.linalg.norm(got - want) / np.linalg.norm(want)
        assert err < 5e-7, (n, inverse, radix, err)


def test_balanced_power_of_two_split(hostfft):
    x = np.ones(512, np.complex64)
    assert run(hostfft, x, False)[2] == [8, 8, 8]
    assert run(hostfft, np.ones(640, np.complex64), False)[2] == [16, 8, 5]
    assert run(hostfft, np.ones(768, np.complex64), False)[2] == [16, 16, 3]


@pytest.mark.parametrize("n", [17, 19, 34, 57, 23 * 8, 1021])
def test_unsupported_sizes_are_rejected(n, hostfft):
    assert run(hostfft, np.ones(n, np.complex64), False)[0] == -1


@pytest.mark.parametrize("n", [64, 72, 96, 120, 128, 144, 160, 192, 200, 224, 240, 256, 288, 320, 360, 384, 400, 448, 480, 512, 576, 600, 640, 720, 768, 800, 896, 960, 1024, 1152, 1200, 1280, 1440, 1536, 1600, 1920, 2048])
@pytest.mark.parametrize("half_in", [False, True])
def test_compile_time_plans_match_numpy(n, half_in, hostfft):
    """b2n_fft_fast.cuh: index maps, staged twiddles and the pair butterflies (incl. radix 10, 12, 16),
    with and without the pruned first stage for zero-padded inputs."""
    rng = np.random.default_rng(1000 + n)
    a = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    b = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    if half_in:
        a[n // 2:] = 0
        b[n // 2:] = 0
        a[n // 2:] = np.nan  # must never be read
    for inverse in (False, True):
        oa, ob = np.empty(n, np.complex64), np.empty(n, np.complex64)
        p = lambda x: x.ctypes.data_as(ctypes.c_void_p)
        assert hostfft.host_fft_fast(n, int(inverse), int(half_in), p(a), p(b), p(oa), p(ob)) == 1
        for x, got in ((np.nan_to_num(a), oa), (b, ob)):
            want = np.fft.ifft(x.astype(np.complex128)) * n if inverse else np.fft.fft(x.astype(np.complex128))
            err = np.linalg.norm(got - want) / np.linalg.norm(want)
            assert err < 5e-7, (n, inverse, err)


@pytest.mark.parametrize("n", [64, 256, 384, 640, 768, 1280])
def test_toeplitz_column_composition_matches_numpy(n, hostfft):
    """What k_fft_cols_toep does to a column pair, on the host with the same butterflies: forward transform of the
    n/2 non-zero rows (pruned first stage), multiply by the kernel spectrum, inverse transform, keep n/2 rows."""
    rng = np.random.default_rng(2000 + n)
    cplx = lambda m: (rng.standard_normal(m) + 1j * rng.standard_normal(m)).astype(np.complex64)
    a, b = cplx(n), cplx(n)
    a[n // 2:] = np.nan  # zero padding: must never be read
    b[n // 2:] = 0
    ka, kb = cplx(n), cplx(n)
    p = lambda x: x.ctypes.data_as(ctypes.c_void_p)
    fa, fb = np.empty(n, np.complex64), np.empty(n, np.complex64)
    assert hostfft.host_fft_fast(n, 0, 1, p(a), p(b), p(fa), p(fb)) == 1
    fa, fb = (fa * ka).astype(np.complex64), (fb * kb).astype(np.complex64)
    oa, ob = np.empty(n, np.complex64), np.empty(n, np.complex64)
    assert hostfft.host_fft_fast(n, 1, 0, p(fa), p(fb), p(oa), p(ob)) == 1
    for x, k, got in ((np.nan_to_num(a), ka, oa), (b, kb, ob)):
        want = (np.fft.ifft(np.fft.fft(x.astype(np.complex128)) * k.astype(np.complex128)) * n)[: n // 2]
        err = np.linalg.norm(got[: n // 2] - want) / np.linalg.norm(want)
        assert err < 1e-6, (n, err)


def test_lengths_without_a_plan_use_the_runtime_passes(hostfft):
    z = np.zeros(100, np.complex64)
    p = z.ctypes.data_as(ctypes.c_void_p)
    assert hostfft.host_fft_fast(100, 0, 0, p, p, p, p) == 0
