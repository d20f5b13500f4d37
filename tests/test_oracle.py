"""Pins the CPU oracle (oracle/kbnufft_oracle.{c,py}) to the REFERENCE:
 (i) the reference's own golden vectors (tests/data/*.pkl -> tests/golden/ref_*_golden.npz),
 (ii) live reference outputs and integer indices (tests/golden/ref_cases.npz, ref_cfg1.npz).
Tolerances: integer indices bit-exact; complex64 values <= 1e-6 rel-L2 (observed
~5e-8: the reference's vectorised complex multiply and libm differ in the last
ulp); complex128 <= 1e-13."""
import os

import numpy as np
import pytest
import torch

import kbnufft_oracle as orc
from conftest import GOLDEN, module_kwargs, rel_l2
from golden_cases import CASES, case_inputs
from torchkbnufft_b200 import workloads
from torchkbnufft_b200._nufft import utils

TOL = {"c64": 1e-6, "c128": 1e-13}
PRECS = {"c64": (np.complex64, torch.complex64), "c128": (np.complex128, torch.complex128)}


def buffers(case, tdtype):
    kw = module_kwargs(case, tdtype)
    pre = utils.init_fn(**kw)
    scaling = utils.compute_scaling_coefs(pre.im_size.tolist(), pre.grid_size.tolist(), pre.numpoints.tolist(),
                                          pre.alpha.tolist(), pre.order.tolist()).to(pre.tables[0].dtype)
    return pre, scaling.numpy()


@pytest.mark.parametrize("prec", ["c64", "c128"])
@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_reference_cases(name, prec, ref_cases):
    case = CASES[name]
    cd, td = PRECS[prec]
    inp = case_inputs(case, cd)
    pre, scaling = buffers(case, td)
    tables = [t.numpy() for t in pre.tables]
    J, L, K, ns = pre.numpoints.numpy(), pre.table_oversamp.numpy(), pre.grid_size.numpy(), pre.n_shift.numpy()
    key = f"{name}_{prec}_"
    tol = TOL[prec]
    fwd = orc.table_interp(inp["grid"], inp["omega"], tables, ns, J, L)
    assert rel_l2(fwd, ref_cases[key + "interp"]) <= tol
    adj = orc.table_interp_adjoint(inp["kdata"], inp["omega"], tables, ns, J, L, K)
    assert rel_l2(adj, ref_cases[key + "interp_adj"]) <= tol
    if inp["omega"].ndim == 2:
        arr_ind, tab_idx = orc.calc_coef_and_indices(inp["omega"], K, J, L)
        assert np.array_equal(arr_ind, ref_cases[key + "arr_ind"])  # bit-exact integer indices
        assert tab_idx.min() >= 0 and all(tab_idx[:, d].max() <= J[d] * L[d] for d in range(len(J)))
    for norm in (None, "ortho"):
        tag = "ortho" if norm else "none"
        sf = orc.nufft_forward(inp["image"], inp["omega"], tables, ns, J, L, scaling, case["im_size"], K,
                               smaps=inp["smaps"], norm=norm)
        assert rel_l2(sf, ref_cases[key + f"sense_fwd_{tag}"]) <= 5 * tol
        sa = orc.nufft_adjoint(inp["kdata"], inp["omega"], tables, ns, J, L, scaling, case["im_size"], K,
                               smaps=inp["smaps"], norm=norm)
        assert rel_l2(sa, ref_cases[key + f"sense_adj_{tag}"]) <= 5 * tol
        if case.get("toep", True):
            kern = ref_cases[key + f"toep_kernel_{tag}"]
            got = orc.toep_nufft(inp["image"], kern, smaps=inp["smaps"], norm=norm)
            ref = ref_cases[key + f"toep_apply_{tag}"]
            # the reference truncates the batch to len(smaps) when a single kernel is used
            # (zip() in modules/kbnufft.py:472); compare what it returned
            assert rel_l2(got[: ref.shape[0]], ref) <= 5 * tol


@pytest.mark.parametrize("which", ["interp", "nufft"])
def test_oracle_matches_reference_pickles(which):
    """The reference's own known-answer files (float64): tests/test_interp.py:16-31 uses
    KbInterp(grid_size=im_size); tests/test_nufft.py:15-30 uses KbNufft defaults."""
    data = np.load(os.path.join(GOLDEN, f"ref_{which}_golden.npz"))
    for i in range(int(data["n_cases"])):
        image, ktraj, kdata = data[f"image_{i}"], data[f"ktraj_{i}"], data[f"kdata_{i}"]
        im_size = image.shape[2:]
        grid_size = im_size if which == "interp" else None
        pre = utils.init_fn(im_size=im_size, grid_size=grid_size, dtype=torch.complex128)
        tables = [t.numpy() for t in pre.tables]
        J, L, K, ns = pre.numpoints.numpy(), pre.table_oversamp.numpy(), pre.grid_size.numpy(), pre.n_shift.numpy()
        if which == "interp":
            got = orc.table_interp(image, ktraj, tables, ns, J, L)
        else:
            scaling = utils.compute_scaling_coefs(pre.im_size.tolist(), pre.grid_size.tolist(),
                                                  pre.numpoints.tolist(), pre.alpha.tolist(),
                                                  pre.order.tolist()).numpy().astype(np.complex128)
            got = orc.nufft_forward(image, ktraj, tables, ns, J, L, scaling, im_size, K)
        assert np.allclose(got, kdata, rtol=1e-5, atol=1e-8)  # the reference's own criterion
        assert rel_l2(got, kdata) <= 1e-12


def test_oracle_matches_reference_cfg1_full_size(ref_cfg1):
    """BASELINE config 1 at full size in complex64 (M = 205 824, 512x512 grid)."""
    wl = workloads.WORKLOADS["cfg1"]
    image, smaps, kdata, omega = workloads.make_inputs(wl, seed=0)
    grid = workloads.complex_normal(np.random.default_rng(1), (1, 1) + wl.grid_size)
    pre = utils.init_fn(im_size=wl.im_size, dtype=torch.complex64)
    tables = [t.numpy() for t in pre.tables]
    J, L, K, ns = pre.numpoints.numpy(), pre.table_oversamp.numpy(), pre.grid_size.numpy(), pre.n_shift.numpy()
    step = int(ref_cfg1["step"])
    arr_ind, _ = orc.calc_coef_and_indices(omega, K, J, L)
    assert np.array_equal(arr_ind[:, ::step], ref_cfg1["arr_ind_sub"])
    fwd = orc.table_interp(grid, omega, tables, ns, J, L, nthreads=4)
    assert rel_l2(fwd[..., ::step], ref_cfg1["interp_sub"]) <= 1e-6
    adj = orc.table_interp_adjoint(kdata, omega, tables, ns, J, L, K, nthreads=4)
    assert rel_l2(adj[..., :48, :48], ref_cfg1["interp_adj_centre"]) <= 1e-6
    assert rel_l2(adj[..., 100:104, :], ref_cfg1["interp_adj_rows"]) <= 1e-6
    assert abs(np.linalg.norm(adj.astype(np.complex128)) / float(ref_cfg1["interp_adj_norm"]) - 1) <= 1e-6
