"""The drop-in boundary without a GPU: the C-ABI library loads and exports every
symbol include/b200nufft.h declares, argument errors come back as negative status
codes (no compute without a device), the product never touches the oracle, and CPU
tensors / a missing library fail loudly instead of falling back."""
import ctypes
import os
import re

import pytest
import torch

import torchkbnufft_b200 as tkbn
from conftest import ROOT
from torchkbnufft_b200 import _lib


def header_symbols():
    text = open(os.path.join(ROOT, "include", "b200nufft.h")).read()
    return sorted(set(re.findall(r"B2N_API\s+[\w\s\*]+?\b(b2n_\w+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    syms = header_symbols()
    for name in ("b2n_points_build", "b2n_interp_forward", "b2n_interp_adjoint", "b2n_export_indices",
                 "b2n_apod_pad", "b2n_crop_apod_coilsum", "b2n_spectrum_mul"):
        assert name in syms


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    for name in header_symbols():
        assert hasattr(lib, name), name
    assert sorted(_lib.SIGNATURES) == header_symbols()  # ctypes binding covers the whole header
    assert lib.b2n_abi_version() == _lib.ABI_VERSION


def test_struct_layouts_match_header_sizes():
    # b2n_geom: 2*int32 + 3*int64 + 3*int32 + 3*int32 + 3*int64 + 3*ptr + 3*double + 3*ptr + 3*double
    assert ctypes.sizeof(_lib.Geom) == 8 + 24 + 12 + 12 + 24 + 24 + 24 + 24 + 24
    # ... + the owner-tile visit lists: 4*int32 + int64 + 4*ptr + 3*ptr (real-weight records) + ptr + int64 (exceptions)
    assert ctypes.sizeof(_lib.Points) == 16 + 16 + 12 + 12 + 16 + 13 * 8 + 16 + 8 + 8 + 4 * 8 + 3 * 8 + 16 + 24
    # and the library's own sizeof (compiled from the header) agrees with the ctypes mirrors
    lib = _lib.load()
    gsz, psz = ctypes.c_size_t(0), ctypes.c_size_t(0)
    assert lib.b2n_struct_sizes(ctypes.byref(gsz), ctypes.byref(psz)) == 0
    assert (gsz.value, psz.value) == (ctypes.sizeof(_lib.Geom), ctypes.sizeof(_lib.Points))


def test_process_wide_options_defaults_and_round_trip():
    """b2n_set_option / b2n_get_option are host-only: defaults of the launch-overlap knobs, round trip, and the status
    code for an unknown option; the enum in the header and the constants of the binding agree."""
    lib = _lib.load()
    text = open(os.path.join(ROOT, "include", "b200nufft.h")).read()
    enum = dict((k, int(v)) for k, v in re.findall(r"(B2N_OPT_[A-Z_]+)\s*=\s*(\d+)", text))
    assert enum["B2N_OPT_TILED_KERNELS"] == _lib.OPT_TILED_KERNELS == 0
    assert enum["B2N_OPT_FAST_FFT"] == _lib.OPT_FAST_FFT
    assert enum["B2N_OPT_PDL"] == _lib.OPT_PDL
    assert enum["B2N_OPT_FFT_PREFETCH"] == _lib.OPT_FFT_PREFETCH
    assert lib.b2n_get_option(_lib.OPT_PDL) == 1 and lib.b2n_get_option(_lib.OPT_FFT_PREFETCH) == 19
    assert lib.b2n_get_option(_lib.OPT_FAST_FFT) == 1 and lib.b2n_get_option(_lib.OPT_TILED_KERNELS) == 1
    try:
        for opt, val in ((_lib.OPT_PDL, 0), (_lib.OPT_PDL, 2), (_lib.OPT_FFT_PREFETCH, 63), (_lib.OPT_FFT_PREFETCH, 0)):
            assert lib.b2n_set_option(opt, val) == 0 and lib.b2n_get_option(opt) == val
    finally:
        lib.b2n_set_option(_lib.OPT_PDL, 1)
        lib.b2n_set_option(_lib.OPT_FFT_PREFETCH, 19)
    assert lib.b2n_set_option(max(enum.values()) + 1, 1) == -1  # B2N_E_ARG
    assert b"unknown option" in lib.b2n_last_error()


def test_argument_errors_are_status_codes():
    lib = _lib.load()
    g = _lib.Geom()
    g.ndim = 7
    n = ctypes.c_size_t(0)
    assert lib.b2n_points_workspace_bytes(ctypes.byref(g), 10, 1, ctypes.byref(n)) == -2  # B2N_E_RANGE
    assert b"ndim" in lib.b2n_last_error()
    g.ndim, g.dtype = 2, 0
    for d in range(2):
        g.grid_size[d], g.numpoints[d], g.table_oversamp[d] = 16, 6, 1024
    status = lib.b2n_points_workspace_bytes(ctypes.byref(g), 100, 1, ctypes.byref(n))
    if lib.b2n_device_count() > 0:
        assert status == 0 and n.value > 100 * (4 + 8 + 12 * 8 + 8)
    else:  # the sort-scratch query needs a device: a positive cudaError_t, never a fake answer
        assert status > 0 and b"CUDA error" in lib.b2n_last_error()
    g.numpoints[1] = 99
    assert lib.b2n_points_workspace_bytes(ctypes.byref(g), 100, 1, ctypes.byref(n)) == -2
    assert lib.b2n_interp_forward(None, None, None, 1, 1, 0, None, None) == -1  # B2N_E_ARG
    sizes = _lib.i64_array((32, 32)), _lib.i64_array((64, 64))
    assert lib.b2n_fft_toeplitz_fused(2, sizes[0], sizes[1], 1, 1, None, 1, None, 1, None, 1, 1.0, None, None, None,
                                      None) == -1
    assert b"twiddle_dev" in lib.b2n_last_error()
    assert lib.b2n_spectrum_mul(5, None, None, 1, 1, 1, 1, 0, 1.0, None) == -1


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "torchkbnufft_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "kbnufft_oracle" not in text and "cpu_engine_shim" not in text, f


def test_cpu_tensors_fail_loudly():
    ob = tkbn.KbNufft(im_size=(8, 8))
    image = torch.randn(1, 1, 8, 8, dtype=torch.complex64)
    omega = torch.rand(2, 10)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ob(image, omega)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        tkbn.KbInterpAdjoint(im_size=(8, 8))(torch.randn(1, 1, 10, dtype=torch.complex64), omega)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        tkbn.ToepNufft()(image, torch.randn(16, 16, dtype=torch.complex64))


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libb200nufft.so")
    with pytest.raises(_lib.EngineError, match="no CPU/PyTorch fallback"):
        _lib.load()


def test_spmat_mode_is_a_torch_sparse_shim():
    """the sparse-matrix mode exists for API completeness only -- built on the host like the
    reference does, applied with torch.sparse, never through libb200nufft.so (tests/test_spmat.py)."""
    real, imag = tkbn.calc_tensor_spmatrix(torch.rand(2, 4), (8, 8))
    assert real.is_sparse and imag.is_sparse and tuple(real.shape) == (4, 256)
    import inspect

    from torchkbnufft_b200._nufft import spmat
    assert "_lib" not in inspect.getsource(spmat)


def test_peer_window_entries_are_host_checkable():
    """The peer-memory all-reduce entries: header constants vs the binding, the communicator mirror's layout, and the
    host-only size query with its argument errors (no device work)."""
    lib = _lib.load()
    text = open(os.path.join(ROOT, "include", "b200nufft.h")).read()
    consts = dict((k, int(v)) for k, v in re.findall(r"#define (B2N_PEER_[A-Z_]+) (\d+)", text))
    assert consts["B2N_PEER_MAX_RANKS"] == _lib.PEER_MAX_RANKS == 16
    assert consts["B2N_PEER_HANDLE_BYTES"] == _lib.PEER_HANDLE_BYTES == 64
    assert ctypes.sizeof(_lib.PeerComm) == 4 + 4 + 8 + 8 * _lib.PEER_MAX_RANKS
    n = ctypes.c_size_t(0)
    assert lib.b2n_peer_window_bytes(8, 2 * 320 * 320, ctypes.byref(n)) == 0
    # header + three slot generations x ranks x (the image rounded up to 4096-float chunks + one chunk of slack)
    assert n.value == 256 + 4 * 3 * 8 * (2 * 320 * 320 + 4096)
    assert lib.b2n_peer_window_bytes(1, 1, ctypes.byref(n)) == 0 and n.value == 256 + 4 * 3 * 2 * 4096
    for world, floats in ((0, 10), (17, 10), (2, 0), (2, 1 << 31)):
        assert lib.b2n_peer_window_bytes(world, floats, ctypes.byref(n)) == -1  # B2N_E_ARG
    assert b"peer window" in lib.b2n_last_error()
    assert lib.b2n_peer_window_bytes(2, 10, None) == -1
    comm = _lib.PeerComm()
    comm.rank, comm.world, comm.max_floats = 3, 2, 100
    assert lib.b2n_peer_allreduce_sum(ctypes.byref(comm), 16, 16, 10, None) == -1  # rank outside the world
    assert lib.b2n_peer_window_close(None) == 0 and lib.b2n_peer_window_destroy(None) == 0
