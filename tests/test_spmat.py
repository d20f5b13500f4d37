"""Sparse-matrix interpolation shim (torchkbnufft_b200/_nufft/spmat.py) against the reference's
``calc_tensor_spmatrix`` and sparse NUFFT outputs stored by oracle/make_golden_spmat.py."""
import os

import numpy as np
import pytest
import torch

import torchkbnufft_b200 as tkbn
from conftest import GOLDEN, rel_l2
from golden_cases import SPMAT_CASES, spmat_case_inputs


@pytest.fixture(scope="module")
def ref():
    return np.load(os.path.join(GOLDEN, "ref_spmat.npz"))


def _kw(cfg):
    return {k: cfg[k] for k in ("grid_size", "numpoints", "n_shift") if cfg.get(k) is not None}


@pytest.mark.parametrize("name", sorted(SPMAT_CASES))
def test_matrices_match_reference(name, ref):
    """Host-side builder: same sparsity pattern, values to rounding (float64) / float32 resolution."""
    cfg = SPMAT_CASES[name]
    omega = spmat_case_inputs(name)[0]
    real, imag = tkbn.calc_tensor_spmatrix(torch.from_numpy(omega), cfg["im_size"], **_kw(cfg))
    assert real.dtype == torch.from_numpy(omega).dtype and real.layout == torch.sparse_coo
    r, i = real.coalesce(), imag.coalesce()
    assert tuple(r.shape) == (cfg["M"], int(np.prod(cfg.get("grid_size") or [2 * n for n in cfg["im_size"]])))
    assert np.array_equal(r.indices().numpy(), ref[f"{name}/index"])
    assert np.array_equal(i.indices().numpy(), ref[f"{name}/index"])
    tol = 1e-12 if omega.dtype == np.float64 else 2e-5  # the reference evaluates float32 trajectories in float32
    want = ref[f"{name}/real"] + 1j * ref[f"{name}/imag"]
    got = r.values().numpy() + 1j * i.values().numpy()
    assert rel_l2(got, want) <= tol


def test_builder_rejects_batched_omega():
    with pytest.raises(ValueError, match="batched omega"):
        tkbn.calc_tensor_spmatrix(torch.zeros(2, 2, 5), (8, 8))


def test_apply_has_no_cpu_path():
    mats = tkbn.calc_tensor_spmatrix(torch.rand(2, 5), (4, 4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        tkbn.functional.kb_spmat_interp(torch.zeros(1, 1, 8, 8, dtype=torch.complex64), mats)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        tkbn.KbInterpAdjoint(im_size=(4, 4))(torch.zeros(1, 1, 5, dtype=torch.complex64), torch.rand(2, 5), mats)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(SPMAT_CASES))
def test_sparse_nufft_matches_reference_and_table_path(name, ref):
    cfg = SPMAT_CASES[name]
    dev = torch.device("cuda:0")
    omega, image, kdata, smaps = spmat_case_inputs(name)
    cdt = torch.complex64 if omega.dtype == np.float32 else torch.complex128
    tol = 1e-5 if cdt == torch.complex64 else 1e-12
    om, x, y, s = (torch.from_numpy(a).to(dev) for a in (omega, image, kdata, smaps))
    mats = tkbn.calc_tensor_spmatrix(om, cfg["im_size"], **_kw(cfg))
    assert mats[0].device == om.device
    nu = tkbn.KbNufft(im_size=cfg["im_size"], dtype=cdt, **_kw(cfg)).to(dev)
    na = tkbn.KbNufftAdjoint(im_size=cfg["im_size"], dtype=cdt, **_kw(cfg)).to(dev)
    got = nu(x, om, mats, smaps=s)
    assert rel_l2(got.cpu().numpy(), ref[f"{name}/nufft"]) <= tol
    assert rel_l2(nu(x, om, mats, smaps=s, norm="ortho").cpu().numpy(), ref[f"{name}/nufft_ortho"]) <= tol
    got_adj = na(y, om, mats, smaps=s)
    assert rel_l2(got_adj.cpu().numpy(), ref[f"{name}/nufft_adj"]) <= 10 * tol
    # the table path approximates the same operator (table quantisation: ~1e-3)
    assert rel_l2(nu(x, om, smaps=s).cpu().numpy(), got.cpu().numpy()) <= 5e-3
    assert rel_l2(na(y, om, smaps=s).cpu().numpy(), got_adj.cpu().numpy()) <= 5e-3
    # interpolation-only modules + real (..., 2) views + autograd through torch.sparse
    ki = tkbn.KbInterp(im_size=cfg["im_size"], dtype=cdt, **_kw(cfg)).to(dev)
    kia = tkbn.KbInterpAdjoint(im_size=cfg["im_size"], dtype=cdt, **_kw(cfg)).to(dev)
    K = ki.grid_size.tolist()
    grid = torch.from_numpy(np.ascontiguousarray(np.broadcast_to(
        np.arange(int(np.prod(K)), dtype=np.float64).reshape(K) % 7 - 3, (1, 2) + tuple(K)))).to(cdt).to(dev)
    assert rel_l2(ki(grid, om, mats).cpu().numpy(), ref[f"{name}/interp"]) <= tol
    assert rel_l2(kia(y[:, :2].contiguous(), om, mats).cpu().numpy(), ref[f"{name}/interp_adj"]) <= 10 * tol
    as_real = torch.view_as_real(grid).contiguous()
    assert rel_l2(torch.view_as_complex(ki(as_real, om, mats)).cpu().numpy(), ref[f"{name}/interp"]) <= tol
    with pytest.raises(TypeError, match="2-tuple"):
        ki(grid, om, [mats[0], mats[1]])
    xg = x.clone().requires_grad_(True)
    out = nu(xg, om, mats, smaps=s)
    (out.abs() ** 2 / 2).sum().backward()
    assert rel_l2(xg.grad.cpu().numpy(), na(out.detach(), om, mats, smaps=s).cpu().numpy()) <= 100 * tol
