"""Parity of the CUDA engine (through the ctypes/C-ABI path) on a B200.

Checkers: (a) the reference's stored outputs (tests/golden/*.npz), (b) the CPU
oracle on the same seeded inputs, (c) size-independent properties at BASELINE's
full sizes.  Tolerances (north_star): complex64 rel-L2 <= 1e-5 forward and <= 1e-4
adjoint; integer grid / table indices bit-exact; complex128 <= 1e-12.
"""
import ctypes
import os

import numpy as np
import pytest
import torch

import kbnufft_oracle as orc
import torchkbnufft_b200 as tkbn
from conftest import GOLDEN, module_kwargs, rel_l2
from golden_cases import CASES, case_inputs
from torchkbnufft_b200 import _lib, workloads
from torchkbnufft_b200._nufft import fft as eng_fft
from torchkbnufft_b200._nufft import interp as eng_interp

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
PRECS = {"c64": (np.complex64, torch.complex64), "c128": (np.complex128, torch.complex128)}
FWD_TOL = {"c64": 1e-5, "c128": 1e-12}
ADJ_TOL = {"c64": 1e-4, "c128": 1e-12}


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def host(t):
    return t.detach().cpu().numpy()


@pytest.fixture(autouse=True)
def _native_library_is_loaded():
    assert os.path.exists(_lib.LIB_PATH), "libb200nufft.so missing: GPU tests must run the native engine"
    _lib.load()
    yield


@pytest.mark.parametrize("prec", ["c64", "c128"])
@pytest.mark.parametrize("name", list(CASES))
def test_cases_match_reference(name, prec, ref_cases):
    case = CASES[name]
    cd, td = PRECS[prec]
    inp = case_inputs(case, cd)
    kw = module_kwargs(case, td)
    key = f"{name}_{prec}_"
    D = lambda k: dev(inp[k])
    interp, interp_adj = tkbn.KbInterp(**kw).to(DEV), tkbn.KbInterpAdjoint(**kw).to(DEV)
    nu, na = tkbn.KbNufft(**kw).to(DEV), tkbn.KbNufftAdjoint(**kw).to(DEV)
    assert rel_l2(host(interp(D("grid"), D("omega"))), ref_cases[key + "interp"]) <= FWD_TOL[prec]
    for mode in ("atomic", "sorted"):
        tkbn.set_adjoint_mode(mode)
        try:
            assert rel_l2(host(interp_adj(D("kdata"), D("omega"))), ref_cases[key + "interp_adj"]) <= ADJ_TOL[prec]
            for norm in (None, "ortho"):
                tag = "ortho" if norm else "none"
                out = na(D("kdata"), D("omega"), smaps=D("smaps"), norm=norm)
                assert rel_l2(host(out), ref_cases[key + f"sense_adj_{tag}"]) <= ADJ_TOL[prec]
        finally:
            tkbn.set_adjoint_mode("atomic")
    for norm in (None, "ortho"):
        tag = "ortho" if norm else "none"
        out = nu(D("image"), D("omega"), smaps=D("smaps"), norm=norm)
        assert rel_l2(host(out), ref_cases[key + f"sense_fwd_{tag}"]) <= FWD_TOL[prec]
    assert rel_l2(host(nu(D("image_multi"), D("omega"))), ref_cases[key + "nufft_fwd_nosmap"]) <= FWD_TOL[prec]
    opts = dict(grid_size=case.get("grid_size"), numpoints=case.get("numpoints", 6),
                table_oversamp=case.get("table_oversamp", 2 ** 10))
    if case.get("toep", True):
        for norm in (None, "ortho"):
            tag = "ortho" if norm else "none"
            kern = tkbn.calc_toeplitz_kernel(D("omega"), case["im_size"], norm=norm, **opts)
            assert rel_l2(host(kern), ref_cases[key + f"toep_kernel_{tag}"]) <= ADJ_TOL[prec]
            got = host(tkbn.ToepNufft()(D("image"), dev(ref_cases[key + f"toep_kernel_{tag}"]), smaps=D("smaps"),
                                        norm=norm))
            ref = ref_cases[key + f"toep_apply_{tag}"]
            assert rel_l2(got[: ref.shape[0]], ref) <= FWD_TOL[prec]
        kern = tkbn.calc_toeplitz_kernel(D("omega"), case["im_size"], weights=D("weights"), norm="ortho", **opts)
        assert rel_l2(host(kern), ref_cases[key + "toep_kernel_weighted"]) <= ADJ_TOL[prec]
    dcomp = tkbn.calc_density_compensation_function(D("omega"), case["im_size"], num_iterations=3,
                                                    n_shift=case.get("n_shift"), **opts)
    assert rel_l2(host(dcomp), ref_cases[key + "dcomp"]) <= ADJ_TOL[prec]


@pytest.mark.parametrize("prec", ["c64", "c128"])
@pytest.mark.parametrize("name", [n for n, c in CASES.items() if c["omega"] != "batched"])
def test_integer_indices_bit_exact(name, prec, ref_cases):
    """Grid and table indices of every neighbour offset equal the reference's
    calc_coef_and_indices output (and the oracle's) exactly."""
    case = CASES[name]
    cd, td = PRECS[prec]
    inp = case_inputs(case, cd)
    ob = tkbn.KbInterp(**module_kwargs(case, td)).to(DEV)
    arr_ind, tab_idx = eng_interp.export_indices(dev(inp["omega"]), ob.tables, ob.n_shift, ob.numpoints,
                                                 ob.table_oversamp, ob.grid_size)
    assert np.array_equal(host(arr_ind), ref_cases[f"{name}_{prec}_arr_ind"])
    o_arr, o_tab = orc.calc_coef_and_indices(inp["omega"], inp["grid_size"], ob.numpoints.tolist(),
                                             ob.table_oversamp.tolist())
    assert np.array_equal(host(tab_idx).astype(np.int64), o_tab)


@pytest.mark.parametrize("which", ["interp", "nufft"])
def test_reference_pickle_goldens(which):
    """The reference's own known-answer tests (tests/test_interp.py:16-31,
    tests/test_nufft.py:15-30), float64, torch.allclose defaults."""
    data = np.load(os.path.join(GOLDEN, f"ref_{which}_golden.npz"))
    for i in range(int(data["n_cases"])):
        image, ktraj, kdata = data[f"image_{i}"], data[f"ktraj_{i}"], data[f"kdata_{i}"]
        im_size = image.shape[2:]
        if which == "interp":
            ob = tkbn.KbInterp(im_size=im_size, grid_size=im_size, dtype=torch.complex128).to(DEV)
        else:
            ob = tkbn.KbNufft(im_size=im_size, dtype=torch.complex128).to(DEV)
        # real-view float64 inputs, as the reference's test feeds them
        out = ob(torch.view_as_real(dev(image)), dev(ktraj))
        assert torch.allclose(out.cpu(), torch.view_as_real(torch.from_numpy(kdata)))


@pytest.mark.parametrize("layout", [_lib.COIL_MAJOR, _lib.CHANNEL_LAST])
@pytest.mark.parametrize("name", ["d1", "d2_mixed", "d3", "d2_batched", "d2_radial"])
def test_layouts_and_modes_agree_with_oracle(name, layout):
    case = CASES[name]
    inp = case_inputs(case, np.complex64)
    ob = tkbn.KbInterp(**module_kwargs(case, torch.complex64)).to(DEV)
    args = (ob.tables, ob.n_shift, ob.numpoints, ob.table_oversamp)
    tables = [host(t) for t in ob.tables]
    J, L, ns = ob.numpoints.tolist(), ob.table_oversamp.tolist(), host(ob.n_shift)
    grid = dev(inp["grid"])
    d = len(case["im_size"])
    to_layout = (lambda g: g.movedim(1, -1).contiguous()) if layout == _lib.CHANNEL_LAST else (lambda g: g)
    from_layout = (lambda g: g.movedim(-1, 1)) if layout == _lib.CHANNEL_LAST else (lambda g: g)
    want = orc.table_interp(inp["grid"], inp["omega"], tables, ns, J, L)
    got = eng_interp.table_interp(to_layout(grid), dev(inp["omega"]), *args, layout=layout)
    assert rel_l2(host(got), want) <= 1e-5
    want = orc.table_interp_adjoint(inp["kdata"], inp["omega"], tables, ns, J, L, inp["grid_size"])
    outs = {}
    for mode in ("atomic", "sorted"):
        got = eng_interp.table_interp_adjoint(dev(inp["kdata"]), dev(inp["omega"]), *args, None, ob.grid_size,
                                              layout=layout, mode=mode)
        outs[mode] = host(from_layout(got))
        assert rel_l2(outs[mode], want) <= 1e-4
    again = eng_interp.table_interp_adjoint(dev(inp["kdata"]), dev(inp["omega"]), *args, None, ob.grid_size,
                                            layout=layout, mode="sorted")
    assert np.array_equal(host(from_layout(again)), outs["sorted"])  # sorted mode is bit-reproducible


def test_fused_fft_side_kernels_match_torch():
    torch.manual_seed(0)
    for dt, tol in ((torch.complex64, 1e-6), (torch.complex128, 1e-14)):
        for N, K, B, C in (((7,), (12,), 2, 3), ((6, 9), (8, 16), 2, 4), ((4, 5, 6), (7, 5, 9), 1, 3)):
            image = torch.randn((B, 1) + N, dtype=dt, device=DEV)
            multi = torch.randn((B, C) + N, dtype=dt, device=DEV)
            smaps = torch.randn((1, C) + N, dtype=dt, device=DEV)
            smaps_b = torch.randn((B, C) + N, dtype=dt, device=DEV)
            scal = torch.randn(N, dtype=dt, device=DEV)
            grid = torch.randn((B, C) + K, dtype=dt, device=DEV)
            pad = []
            for k, n in zip(reversed(K), reversed(N)):
                pad += [0, k - n]
            crop = (slice(None), slice(None)) + tuple(slice(0, n) for n in N)
            for sm in (smaps, smaps_b):
                want = torch.nn.functional.pad(image * sm * scal * 0.5, pad)
                assert rel_l2(host(eng_fft.apod_pad(image, K, sm, scal, 0.5)), host(want)) <= tol
                cl = eng_fft.apod_pad(image, K, sm.movedim(1, -1).contiguous(), scal, 0.5, layout=_lib.CHANNEL_LAST)
                assert rel_l2(host(cl.movedim(-1, 1)), host(want)) <= tol
                want = torch.sum(grid[crop] * scal.conj() * sm.conj(), 1, keepdim=True) * 0.25
                assert rel_l2(host(eng_fft.crop_apod_coilsum(grid, N, sm, scal, 0.25)), host(want)) <= tol
                cl = eng_fft.crop_apod_coilsum(grid.movedim(1, -1).contiguous(), N, sm.movedim(1, -1).contiguous(),
                                               scal, 0.25, layout=_lib.CHANNEL_LAST)
                assert rel_l2(host(cl), host(want)) <= tol
            want = torch.nn.functional.pad(multi * scal, pad)
            assert rel_l2(host(eng_fft.apod_pad(multi, K, None, scal, 1.0)), host(want)) <= tol
            assert rel_l2(host(eng_fft.apod_pad(multi, K, None, None, 1.0)), host(torch.nn.functional.pad(multi, pad))) == 0
            want = grid[crop] * scal.conj()
            assert rel_l2(host(eng_fft.crop_apod_coilsum(grid, N, None, scal, 1.0)), host(want)) <= tol
            for kern in (torch.randn(K, dtype=dt, device=DEV), torch.randn((B,) + K, dtype=dt, device=DEV)):
                kb = kern if kern.ndim == len(K) else kern.unsqueeze(1)
                want = grid * kb * 2.0
                assert rel_l2(host(eng_fft.spectrum_mul_(grid.clone(), kern, 2.0)), host(want)) <= tol
                cl = eng_fft.spectrum_mul_(grid.movedim(1, -1).contiguous(), kern, 2.0, layout=_lib.CHANNEL_LAST)
                assert rel_l2(host(cl.movedim(-1, 1)), host(want)) <= tol


def test_cfg1_full_size_against_reference(ref_cfg1):
    """BASELINE config 1 (256^2, M = 205 824, complex64) against the reference's stored outputs."""
    wl = workloads.WORKLOADS["cfg1"]
    image, smaps, kdata, omega = workloads.make_inputs(wl, seed=0)
    grid = workloads.complex_normal(np.random.default_rng(1), (1, 1) + wl.grid_size)
    kw = dict(im_size=wl.im_size, dtype=torch.complex64)
    interp, interp_adj = tkbn.KbInterp(**kw).to(DEV), tkbn.KbInterpAdjoint(**kw).to(DEV)
    nu, na = tkbn.KbNufft(**kw).to(DEV), tkbn.KbNufftAdjoint(**kw).to(DEV)
    step = int(ref_cfg1["step"])
    om = dev(omega)
    arr_ind, _ = eng_interp.export_indices(om, interp.tables, interp.n_shift, interp.numpoints, interp.table_oversamp,
                                           interp.grid_size)
    assert np.array_equal(host(arr_ind)[:, ::step], ref_cfg1["arr_ind_sub"])
    assert rel_l2(host(interp(dev(grid), om))[..., ::step], ref_cfg1["interp_sub"]) <= 1e-5
    assert rel_l2(host(nu(dev(image), om))[..., ::step], ref_cfg1["nufft_sub"]) <= 1e-5
    for mode in ("atomic", "sorted"):
        tkbn.set_adjoint_mode(mode)
        try:
            adj = host(interp_adj(dev(kdata), om))
            assert rel_l2(adj[..., :48, :48], ref_cfg1["interp_adj_centre"]) <= 1e-4
            assert rel_l2(adj[..., 100:104, :], ref_cfg1["interp_adj_rows"]) <= 1e-4
            assert abs(np.linalg.norm(adj.astype(np.complex128)) / float(ref_cfg1["interp_adj_norm"]) - 1) <= 1e-5
            assert rel_l2(host(na(dev(kdata), om))[..., ::4, ::4], ref_cfg1["nufft_adj"]) <= 1e-4
        finally:
            tkbn.set_adjoint_mode("atomic")


def test_cfg2_full_size_against_oracle_and_properties():
    """BASELINE config 2 (the benchmark workload): SENSE forward/adjoint vs the oracle,
    plus adjointness, linearity and atomic-vs-sorted agreement at full size."""
    wl = workloads.WORKLOADS["cfg2"]
    image, smaps, kdata, omega = workloads.make_inputs(wl, seed=0)
    nu = tkbn.KbNufft(im_size=wl.im_size, dtype=torch.complex64).to(DEV)
    na = tkbn.KbNufftAdjoint(im_size=wl.im_size, dtype=torch.complex64).to(DEV)
    tables = [host(t) for t in nu.tables]
    J, L, ns = nu.numpoints.tolist(), nu.table_oversamp.tolist(), host(nu.n_shift)
    scaling = host(nu.scaling_coef)
    threads = os.cpu_count() or 1
    want_f = orc.nufft_forward(image, omega, tables, ns, J, L, scaling, wl.im_size, wl.grid_size, smaps=smaps,
                               nthreads=threads)
    want_a = orc.nufft_adjoint(kdata, omega, tables, ns, J, L, scaling, wl.im_size, wl.grid_size, smaps=smaps,
                               nthreads=threads)
    x, s, y, om = dev(image), dev(smaps), dev(kdata), dev(omega)
    got_f = nu(x, om, smaps=s)
    assert rel_l2(host(got_f), want_f) <= 1e-5
    outs = {}
    for mode in ("atomic", "sorted"):
        tkbn.set_adjoint_mode(mode)
        try:
            outs[mode] = na(y, om, smaps=s)
            assert rel_l2(host(outs[mode]), want_a) <= 1e-4
        finally:
            tkbn.set_adjoint_mode("atomic")
    assert rel_l2(host(outs["atomic"]), host(outs["sorted"])) <= 1e-5
    # <A x, y> = <x, A^H y> (float64 accumulation of complex64 results)
    lhs = torch.sum(got_f.to(torch.complex128).conj() * y.to(torch.complex128))
    rhs = torch.sum(x.to(torch.complex128).conj() * outs["sorted"].to(torch.complex128))
    assert float(abs(lhs - rhs) / abs(lhs)) <= 1e-5
    # linearity
    x2 = dev(workloads.complex_normal(np.random.default_rng(5), image.shape))
    lin = nu(x + 2.0 * x2, om, smaps=s)
    assert rel_l2(host(lin), host(got_f + 2.0 * nu(x2, om, smaps=s))) <= 1e-5


def test_cfg4_shape_3d_properties_reduced():
    """3-D kooshball at reduced size (the oracle finishes in seconds): forward and both
    adjoint modes vs the oracle, adjointness."""
    wl = workloads.Workload("cfg4s", (32, 32, 32), 4, 1, 512, 64, "koosh", "reduced 3-D kooshball")
    image, smaps, kdata, omega = workloads.make_inputs(wl, seed=3)
    nu = tkbn.KbNufft(im_size=wl.im_size, dtype=torch.complex64).to(DEV)
    na = tkbn.KbNufftAdjoint(im_size=wl.im_size, dtype=torch.complex64).to(DEV)
    tables = [host(t) for t in nu.tables]
    J, L, ns = nu.numpoints.tolist(), nu.table_oversamp.tolist(), host(nu.n_shift)
    scaling = host(nu.scaling_coef)
    threads = os.cpu_count() or 1
    x, s, y, om = dev(image), dev(smaps), dev(kdata), dev(omega)
    want_f = orc.nufft_forward(image, omega, tables, ns, J, L, scaling, wl.im_size, wl.grid_size, smaps=smaps,
                               nthreads=threads)
    assert rel_l2(host(nu(x, om, smaps=s)), want_f) <= 1e-5
    want_a = orc.nufft_adjoint(kdata, omega, tables, ns, J, L, scaling, wl.im_size, wl.grid_size, smaps=smaps,
                               nthreads=threads)
    for mode in ("atomic", "sorted"):
        tkbn.set_adjoint_mode(mode)
        try:
            assert rel_l2(host(na(y, om, smaps=s)), want_a) <= 1e-4
        finally:
            tkbn.set_adjoint_mode("atomic")


def _inner(a, b):
    return torch.sum(a.to(torch.complex128).conj() * b.to(torch.complex128))


def test_cfg3_full_size_toeplitz_properties():
    """BASELINE config 3 (384^2, 32 coils, ToepNufft) at full size through size-independent
    properties: the Toeplitz normal operator equals A^H A, is Hermitian and linear; A and A^H are
    adjoint.  A sub-sampled slice of the forward is checked against the oracle."""
    wl = workloads.WORKLOADS["cfg3"]
    image, smaps, kdata, omega = workloads.make_inputs(wl, seed=0)
    x, s, y, om = dev(image), dev(smaps), dev(kdata), dev(omega)
    nu = tkbn.KbNufft(im_size=wl.im_size, dtype=torch.complex64).to(DEV)
    na = tkbn.KbNufftAdjoint(im_size=wl.im_size, dtype=torch.complex64).to(DEV)
    toep = tkbn.ToepNufft()
    kern = tkbn.calc_toeplitz_kernel(om, wl.im_size, norm="ortho")
    ax = nu(x, om, smaps=s, norm="ortho")
    normal = na(ax, om, smaps=s, norm="ortho")
    tx = toep(x, kern, smaps=s, norm="ortho")
    assert rel_l2(host(tx), host(normal)) <= 1e-4
    x2 = dev(workloads.complex_normal(np.random.default_rng(9), image.shape))
    tx2 = toep(x2, kern, smaps=s, norm="ortho")
    assert float(abs(_inner(tx, x2) - _inner(x, tx2)) / abs(_inner(tx, x2))) <= 1e-4  # T = T^H
    assert rel_l2(host(toep(x - 0.5 * x2, kern, smaps=s, norm="ortho")), host(tx - 0.5 * tx2)) <= 1e-5
    assert float(abs(_inner(ax, y) - _inner(x, na(y, om, smaps=s, norm="ortho"))) / abs(_inner(ax, y))) <= 1e-5
    # first 4 coils x every 7th sample against the oracle (the full 32-coil oracle run takes too long here)
    tables = [host(t) for t in nu.tables]
    J, L, ns = nu.numpoints.tolist(), nu.table_oversamp.tolist(), host(nu.n_shift)
    sub = np.ascontiguousarray(omega[:, ::7])
    want = orc.nufft_forward(image, sub, tables, ns, J, L, host(nu.scaling_coef), wl.im_size, wl.grid_size,
                             smaps=smaps[:, :4], nthreads=os.cpu_count() or 1, norm="ortho")
    assert rel_l2(host(ax)[:, :4, ::7], want) <= 1e-5


def test_cfg4_full_size_3d_properties():
    """BASELINE config 4 (128^3, 8 coils, M = 8 388 608) at full size: adjointness, linearity,
    atomic run-to-run agreement and the autograd backward of 0.5*|Ax|^2 = A^H A x."""
    wl = workloads.WORKLOADS["cfg4"]
    image, smaps, kdata, omega = workloads.make_inputs(wl, seed=0)
    x, s, y, om = dev(image), dev(smaps), dev(kdata), dev(omega)
    nu = tkbn.KbNufft(im_size=wl.im_size, dtype=torch.complex64).to(DEV)
    na = tkbn.KbNufftAdjoint(im_size=wl.im_size, dtype=torch.complex64).to(DEV)
    ax = nu(x, om, smaps=s)
    ahy = na(y, om, smaps=s)
    assert float(abs(_inner(ax, y) - _inner(x, ahy)) / abs(_inner(ax, y))) <= 1e-5
    x2 = dev(workloads.complex_normal(np.random.default_rng(4), image.shape))
    assert rel_l2(host(nu(x + 3.0 * x2, om, smaps=s)), host(ax + 3.0 * nu(x2, om, smaps=s))) <= 1e-5
    assert rel_l2(host(na(y, om, smaps=s)), host(ahy)) <= 1e-6  # atomic order differs, values agree
    xg = x.clone().requires_grad_(True)
    out = nu(xg, om, smaps=s)
    (out.abs() ** 2 / 2).sum().backward()
    assert rel_l2(host(xg.grad), host(na(ax, om, smaps=s))) <= 1e-5
    # a random subset of samples against a direct per-sample evaluation of the oracle on the same grid
    pick = np.random.default_rng(11).choice(wl.n_points, size=4096, replace=False)
    tables = [host(t) for t in nu.tables]
    J, L, ns = nu.numpoints.tolist(), nu.table_oversamp.tolist(), host(nu.n_shift)
    want = orc.nufft_forward(image, np.ascontiguousarray(omega[:, pick]), tables, ns, J, L, host(nu.scaling_coef),
                             wl.im_size, wl.grid_size, smaps=smaps, nthreads=os.cpu_count() or 1)
    assert rel_l2(host(ax)[..., pick], want) <= 1e-5


def test_cfg5_one_gpu_share_batch_consistency():
    """BASELINE config 5, the 8 slices x 16 coils one GPU owns under 8-way batch sharding: the
    batched calls equal slice-by-slice calls, forward and adjoint, and are mutually adjoint."""
    wl = workloads.WORKLOADS["cfg5"]
    image, smaps, kdata, omega = workloads.make_inputs(wl, seed=0, n_batch=8)
    x, s, y, om = dev(image), dev(smaps), dev(kdata), dev(omega)
    nu = tkbn.KbNufft(im_size=wl.im_size, dtype=torch.complex64).to(DEV)
    na = tkbn.KbNufftAdjoint(im_size=wl.im_size, dtype=torch.complex64).to(DEV)
    ax, ahy = nu(x, om, smaps=s), na(y, om, smaps=s)
    for b in (0, 3, 7):
        sb = s if s.shape[0] == 1 else s[b:b + 1]
        assert rel_l2(host(nu(x[b:b + 1], om, smaps=sb)), host(ax[b:b + 1])) <= 1e-6
        assert rel_l2(host(na(y[b:b + 1], om, smaps=sb)), host(ahy[b:b + 1])) <= 1e-5
    assert float(abs(_inner(ax, y) - _inner(x, ahy)) / abs(_inner(ax, y))) <= 1e-5


def _oracle_args(nu):
    tables = [host(t) for t in nu.tables]
    return tables, host(nu.n_shift), nu.numpoints.tolist(), nu.table_oversamp.tolist(), host(nu.scaling_coef)


def test_cfg4_full_size_adjoint_against_oracle_on_a_spoke_subset():
    """BASELINE config 4 at FULL size (256^3 grid, the 8.4 M-point plan) against the oracle.  The adjoint is linear in
    the samples: with every sample zero except those of each 64th spoke, the engine's full-plan adjoint must equal
    the oracle's adjoint of that subset alone (131 072 points x 216 taps x 8 coils -- seconds on the host).  Both
    accumulation modes; then the autograd gradient of 0.5*|A_sub x|^2 against the oracle's A_sub^H A_sub x."""
    wl = workloads.WORKLOADS["cfg4"]
    image, smaps, kdata, omega = workloads.make_inputs(wl, seed=0)
    spoke = np.arange(wl.n_points) // wl.n_read
    pick = np.nonzero(spoke % 64 == 0)[0]
    assert pick.size == wl.n_points // 64
    sparse = np.zeros_like(kdata)
    sparse[..., pick] = kdata[..., pick]
    x, s, om = dev(image), dev(smaps), dev(omega)
    nu = tkbn.KbNufft(im_size=wl.im_size, dtype=torch.complex64).to(DEV)
    na = tkbn.KbNufftAdjoint(im_size=wl.im_size, dtype=torch.complex64).to(DEV)
    tables, ns, J, L, scaling = _oracle_args(nu)
    threads = os.cpu_count() or 1
    om_sub = np.ascontiguousarray(omega[:, pick])
    want = orc.nufft_adjoint(np.ascontiguousarray(kdata[..., pick]), om_sub, tables, ns, J, L, scaling, wl.im_size,
                             wl.grid_size, smaps=smaps, nthreads=threads)
    y = dev(sparse)
    for mode in ("atomic", "sorted"):
        tkbn.set_adjoint_mode(mode)
        try:
            assert rel_l2(host(na(y, om, smaps=s)), want) <= 1e-4, mode
        finally:
            tkbn.set_adjoint_mode("atomic")
    del y
    # gradient of 0.5 * |A_sub x|^2 = A_sub^H A_sub x, oracle on both legs
    oms = dev(om_sub)
    xg = x.clone().requires_grad_(True)
    (nu(xg, oms, smaps=s).abs() ** 2 / 2).sum().backward()
    ax = orc.nufft_forward(image, om_sub, tables, ns, J, L, scaling, wl.im_size, wl.grid_size, smaps=smaps,
                           nthreads=threads)
    want_g = orc.nufft_adjoint(ax, om_sub, tables, ns, J, L, scaling, wl.im_size, wl.grid_size, smaps=smaps,
                               nthreads=threads)
    assert rel_l2(host(xg.grad), want_g) <= 1e-4


def test_cfg5_slices_of_the_batch_against_oracle():
    """BASELINE config 5, one GPU's share (8 slices x 16 coils, shared trajectory): two slices of the BATCHED forward
    and adjoint calls against the oracle run on those slices alone."""
    wl = workloads.WORKLOADS["cfg5"]
    image, smaps, kdata, omega = workloads.make_inputs(wl, seed=0, n_batch=8)
    x, s, y, om = dev(image), dev(smaps), dev(kdata), dev(omega)
    nu = tkbn.KbNufft(im_size=wl.im_size, dtype=torch.complex64).to(DEV)
    na = tkbn.KbNufftAdjoint(im_size=wl.im_size, dtype=torch.complex64).to(DEV)
    tables, ns, J, L, scaling = _oracle_args(nu)
    threads = os.cpu_count() or 1
    ax = host(nu(x, om, smaps=s))
    outs = {}
    for mode in ("atomic", "sorted"):
        tkbn.set_adjoint_mode(mode)
        try:
            outs[mode] = host(na(y, om, smaps=s))
        finally:
            tkbn.set_adjoint_mode("atomic")
    for b in (0, 5):
        want_f = orc.nufft_forward(image[b:b + 1], omega, tables, ns, J, L, scaling, wl.im_size, wl.grid_size,
                                   smaps=smaps, nthreads=threads)
        assert rel_l2(ax[b:b + 1], want_f) <= 1e-5
        want_a = orc.nufft_adjoint(kdata[b:b + 1], omega, tables, ns, J, L, scaling, wl.im_size, wl.grid_size,
                                   smaps=smaps, nthreads=threads)
        for mode in outs:
            assert rel_l2(outs[mode][b:b + 1], want_a) <= 1e-4, (b, mode)


def test_cfg3_full_size_toeplitz_against_oracle():
    """BASELINE config 3 at full size (384^2 image, 768^2 kernel, M = 184 320) on a 4-coil subset: the engine's
    ToepNufft apply with the engine-built kernel against (i) the oracle's fft_filter pipeline given the same kernel
    (the apply alone, 1e-5) and (ii) the oracle's A^H A x, which involves neither the kernel builder nor the filter
    (the whole Toeplitz path; 1e-4, the embedding itself is exact up to the interpolation error)."""
    wl = workloads.WORKLOADS["cfg3"]
    image, smaps, _kdata, omega = workloads.make_inputs(wl, seed=0)
    sub = np.ascontiguousarray(smaps[:, :4])
    x, s4, om = dev(image), dev(sub), dev(omega)
    nu = tkbn.KbNufft(im_size=wl.im_size, dtype=torch.complex64).to(DEV)
    tables, ns, J, L, scaling = _oracle_args(nu)
    threads = os.cpu_count() or 1
    for norm in ("ortho", None):
        kern = tkbn.calc_toeplitz_kernel(om, wl.im_size, norm=norm)
        got = host(tkbn.ToepNufft()(x, kern, smaps=s4, norm=norm))
        want_apply = orc.toep_nufft(image, host(kern), smaps=sub, norm=norm, workers=threads)
        assert rel_l2(got, want_apply) <= 1e-5, norm
        ax = orc.nufft_forward(image, omega, tables, ns, J, L, scaling, wl.im_size, wl.grid_size, smaps=sub,
                               nthreads=threads, norm=norm)
        want_normal = orc.nufft_adjoint(ax, omega, tables, ns, J, L, scaling, wl.im_size, wl.grid_size, smaps=sub,
                                        nthreads=threads, norm=norm)
        assert rel_l2(got, want_normal) <= 1e-4, norm


def test_edge_shapes():
    kw = dict(im_size=(8, 6), dtype=torch.complex64)
    interp, adj = tkbn.KbInterp(**kw).to(DEV), tkbn.KbInterpAdjoint(**kw).to(DEV)
    grid = torch.randn(2, 3, 16, 12, dtype=torch.complex64, device=DEV)
    # a single sample; the same sample repeated (collisions); (1, d, M) broadcast; non-contiguous input
    om1 = torch.tensor([[0.3], [-2.0]], device=DEV)
    assert interp(grid, om1).shape == (2, 3, 1)
    om_rep = om1.repeat(1, 257)
    out = interp(grid, om_rep)
    assert torch.allclose(out, out[..., :1].expand_as(out))
    y = torch.ones(2, 3, 257, dtype=torch.complex64, device=DEV)
    for mode in ("atomic", "sorted"):
        tkbn.set_adjoint_mode(mode)
        try:
            g = adj(y, om_rep)
            g1 = adj(y[..., :1], om1)
            assert rel_l2(host(g), host(g1) * 257) <= 1e-5
        finally:
            tkbn.set_adjoint_mode("atomic")
    assert torch.equal(interp(grid, om_rep[None]), out)
    nc = torch.randn(2, 3, 12, 16, dtype=torch.complex64, device=DEV).transpose(-1, -2)
    assert torch.equal(interp(nc, om_rep), interp(nc.contiguous(), om_rep))
    # empty trajectory
    om0 = torch.zeros(2, 0, device=DEV)
    assert interp(grid, om0).shape == (2, 3, 0)
    z = adj(torch.zeros(2, 3, 0, dtype=torch.complex64, device=DEV), om0)
    assert z.shape == (2, 3, 16, 12) and float(z.abs().max()) == 0.0
    # dtype transfer smoke (tests/test_interp.py:340-362)
    ob64 = tkbn.KbNufft(im_size=(8, 6)).to(DEV).to(torch.float64)
    img = torch.randn(1, 1, 8, 6, dtype=torch.complex128, device=DEV)
    assert ob64(img, om_rep.double()).dtype == torch.complex128


def test_plan_cache_tracks_trajectory_edits():
    kw = dict(im_size=(8, 8), dtype=torch.complex64)
    interp = tkbn.KbInterp(**kw).to(DEV)
    grid = torch.randn(1, 1, 16, 16, dtype=torch.complex64, device=DEV)
    om = (torch.rand(2, 50, device=DEV) - 0.5) * 6
    a = interp(grid, om)
    assert torch.equal(interp(grid, om), a)  # cached plan, identical result
    om.mul_(0.5)  # in-place edit bumps the version -> new plan
    b = interp(grid, om)
    assert not torch.equal(a, b)
    assert torch.equal(b, interp(grid, om.clone()))


@pytest.mark.parametrize("grid_size", [(32, 32), (40, 56), (16, 20), (13, 37), (64, 19), (18, 27), (22, 36)])
@pytest.mark.parametrize("B, C, batched", [(1, 1, False), (2, 2, False), (1, 3, False), (1, 5, False), (2, 8, True),
                                           (1, 12, False), (1, 16, False), (1, 20, False), (3, 32, True)])
def test_tiled_kernels_match_generic_and_oracle(grid_size, B, C, batched):
    """The tiled / owner-tile kernels (complex64, 2-D, J=6) against the generic kernels
    and the oracle: every coil-chunk width, partial tiles (also of the 4 x 8 owner tiles: 18, 22, 27, 19 cells),
    grids smaller than a tile, periodic wrap, dense sub-problems / work items, batched trajectories.  "force" runs the
    owner-tile spread where the plan carries visit lists; the round-1 shared-memory variants (the fallback for
    non-standard tables) are exercised with the owner path switched off."""
    rng = np.random.default_rng(hash((grid_size, B, C)) & 0xFFFF)
    im_size = tuple(max(2, k // 2) for k in grid_size)
    ob = tkbn.KbInterp(im_size=im_size, grid_size=grid_size, dtype=torch.complex64).to(DEV)
    M = 3000
    shape = (B, 2, M) if batched else (2, M)
    omega = rng.uniform(-np.pi, np.pi, size=shape)
    omega[..., : M // 3] *= 0.05  # dense clump around k = 0: several sub-problems in one tile
    omega = np.ascontiguousarray(omega.astype(np.float32))
    grid = workloads.complex_normal(rng, (B, C) + tuple(grid_size))
    kdata = workloads.complex_normal(rng, (B, C, M))
    args = (ob.tables, ob.n_shift, ob.numpoints, ob.table_oversamp)
    tables = [host(t) for t in ob.tables]
    J, L, ns = ob.numpoints.tolist(), ob.table_oversamp.tolist(), host(ob.n_shift)
    want_f = orc.table_interp(grid, omega, tables, ns, J, L)
    want_a = orc.table_interp_adjoint(kdata, omega, tables, ns, J, L, grid_size)
    res = {}
    try:
        for tiled in ("force", False):
            tkbn.set_tiled_kernels(tiled)
            res[tiled] = (host(eng_interp.table_interp(dev(grid), dev(omega), *args)),
                          host(eng_interp.table_interp_adjoint(dev(kdata), dev(omega), *args, None, ob.grid_size,
                                                               mode="atomic")))
    finally:
        tkbn.set_tiled_kernels(True)
    # the forced tiled adjoint variants (warp-owned rows / warp-owned coils / warp-private tiles) and the
    # 8-coil forward chunks
    lib = _lib.load()
    for variant in (0, 1, 2, 3, 4, 5, 6):
        try:
            eng_interp.owned_spread = False
            lib.b2n_set_option(_lib.OPT_ADJ_ROW_OWNERSHIP, variant)
            alt = host(eng_interp.table_interp_adjoint(dev(kdata), dev(omega), *args, None, ob.grid_size, mode="atomic"))
        finally:
            eng_interp.owned_spread = True
            lib.b2n_set_option(_lib.OPT_ADJ_ROW_OWNERSHIP, 0)
        assert rel_l2(alt, want_a) <= 1e-4, f"adjoint variant {variant}"
    for chunk in (8, 1):  # default 0 = one 16-coil CTA per sub-problem; 1 = persistent kernel; 8 = 8-coil CTAs
        try:
            lib.b2n_set_option(_lib.OPT_FWD_COIL_CHUNK, chunk)
            alt = host(eng_interp.table_interp(dev(grid), dev(omega), *args))
        finally:
            lib.b2n_set_option(_lib.OPT_FWD_COIL_CHUNK, 0)
        assert rel_l2(alt, want_f) <= 1e-5, f"forward variant {chunk}"
    for tiled in ("force", False):
        assert rel_l2(res[tiled][0], want_f) <= 1e-5, f"forward tiled={tiled}"
        assert rel_l2(res[tiled][1], want_a) <= 1e-4, f"adjoint tiled={tiled}"
    assert rel_l2(res["force"][0], res[False][0]) <= 2e-6
    assert rel_l2(res["force"][1], res[False][1]) <= 2e-6


@pytest.mark.parametrize("N, K", [((24,), (40,)), ((19,), (64,)), ((16, 12), (32, 24)), ((13, 18), (26, 35)),
                                  ((320, 320), (640, 640)), ((33, 30), (70, 64)), ((8, 7, 6), (16, 14, 12)),
                                  ((17, 19, 12), (36, 40, 24)),
                                  # lengths with a compile-time plan (b2n_fft_fast.cuh), mixed with run-time ones
                                  ((300,), (512,)), ((129,), (256,)), ((33, 64), (64, 128)), ((160, 200), (320, 768)),
                                  ((5, 500), (7, 1024)), ((64, 9), (128, 18)), ((31, 15), (64, 30)),
                                  ((20, 33, 48), (64, 64, 128)), ((6, 128, 100), (12, 256, 320)),
                                  ((48, 96), (96, 192)), ((150, 288), (384, 576)), ((3, 640), (6, 1280)),
                                  ((1000,), (2048,)), ((4, 700), (8, 2048)), ((40, 40, 40), (96, 96, 96)),
                                  ((112, 224), (224, 448)), ((144, 240), (288, 480)), ((5, 448), (10, 896)),
                                  ((480,), (960,)),
                                  # the reference notebooks' 400 x 400 image (Toeplitz Example.ipynb) and the other
                                  # lengths added in round 2
                                  ((400, 400), (800, 800)), ((80, 100), (160, 200)), ((120, 200), (240, 400)),
                                  ((6, 576), (12, 1152)), ((5, 768), (8, 1536)), ((4, 800), (8, 1600)),
                                  ((3, 900), (6, 1920)),
                                  # radix-9 / radix-15 stages
                                  ((36, 60), (72, 120)), ((70, 180), (144, 360)), ((300, 5), (600, 10)),
                                  ((360, 4), (720, 8)), ((4, 600), (8, 1200)), ((5, 720), (10, 1440))])
def test_fused_pruned_fft_matches_torch(N, K):
    """Own Stockham passes (pruned inputs / cropped outputs, fused apodisation, SENSE multiply,
    coil sum and Toeplitz kernel multiply) against torch.fft + plain torch ops, complex64."""
    torch.manual_seed(1)
    dt = torch.complex64
    eng_fft.use_fused_fft = True
    assert eng_fft.fused_fft_available(dt, K)
    d = len(N)
    dims = list(range(-d, 0))
    pad = []
    for k, n in zip(reversed(K), reversed(N)):
        pad += [0, k - n]
    crop = (slice(None), slice(None)) + tuple(slice(0, n) for n in N)
    for B, C in ((1, 1), (2, 3), (1, 16)):
        image = torch.randn((B, 1) + N, dtype=dt, device=DEV)
        multi = torch.randn((B, C) + N, dtype=dt, device=DEV)
        scal = torch.randn(N, dtype=dt, device=DEV)
        grid = torch.randn((B, C) + K, dtype=dt, device=DEV)
        for smaps in (torch.randn((1, C) + N, dtype=dt, device=DEV), torch.randn((B, C) + N, dtype=dt, device=DEV)):
            want = torch.fft.fftn(torch.nn.functional.pad(image * smaps * scal * 0.5, pad), dim=dims)
            got = eng_fft.fused_fft_forward(image, K, smaps, scal, 0.5)
            assert rel_l2(host(got), host(want)) <= 2e-6
            want = torch.sum(torch.fft.ifftn(grid, dim=dims, norm="forward")[crop] * scal.conj() * smaps.conj(), 1,
                             keepdim=True) * 0.25
            got = eng_fft.fused_fft_adjoint(grid, N, smaps, scal, 0.25)
            assert rel_l2(host(got), host(want)) <= 2e-6
            if d > 1:
                for kern in (torch.randn(K, dtype=dt, device=DEV), torch.randn((B,) + K, dtype=dt, device=DEV)):
                    kb = kern if kern.ndim == d else kern.unsqueeze(1)
                    want = torch.sum(torch.fft.ifftn(grid * kb, dim=dims, norm="forward")[crop] * smaps.conj(), 1,
                                     keepdim=True)
                    got = eng_fft.fused_fft_adjoint(grid, N, smaps, None, 1.0, kernel=kern)
                    assert rel_l2(host(got), host(want)) <= 2e-6
        want = torch.fft.fftn(torch.nn.functional.pad(multi * scal, pad), dim=dims)
        assert rel_l2(host(eng_fft.fused_fft_forward(multi, K, None, scal, 1.0)), host(want)) <= 2e-6
        want = torch.fft.ifftn(grid, dim=dims, norm="forward")[crop] * scal.conj()
        assert rel_l2(host(eng_fft.fused_fft_adjoint(grid, N, None, scal, 1.0)), host(want)) <= 2e-6
    assert not eng_fft.fused_fft_available(dt, (57,)) and not eng_fft.fused_fft_available(torch.complex128, K)
    eng_fft.use_fused_fft = "auto"
    assert eng_fft.fused_fft_available(dt, (640, 256)) and not eng_fft.fused_fft_available(dt, (640, 24))
    assert all(eng_fft.fused_fft_available(dt, (n,)) for n in (64, 72, 96, 120, 128, 144, 160, 192, 200, 224, 240, 256, 288, 320, 360, 384, 400, 448, 480, 512, 576, 600, 640, 720, 768, 800, 896, 960, 1024, 1152, 1200, 1280, 1440, 1536, 1600, 1920, 2048))


@pytest.mark.parametrize("N, K", [((32, 32), (64, 64)), ((48, 30), (96, 60)), ((64, 100), (128, 200)),
                                  ((90, 18), (192, 36)), ((112, 64), (224, 128)), ((128, 128), (256, 256)),
                                  ((144, 33), (288, 66)), ((160, 160), (320, 320)), ((192, 50), (384, 100)),
                                  ((224, 21), (448, 42)), ((240, 64), (480, 128)), ((256, 256), (512, 512)),
                                  ((288, 40), (576, 80)), ((320, 320), (640, 640)), ((384, 384), (768, 768)),
                                  ((448, 20), (896, 40)), ((480, 12), (960, 24)), ((512, 30), (1024, 60)),
                                  ((640, 16), (1280, 32)), ((100, 150), (256, 320)), ((31, 37), (64, 96))])
def test_toeplitz_three_pass_matches_torch_fft(N, K):
    """b2n_fft_toeplitz_fused (forward rows, one column pass that transforms / filters / transforms back, inverse
    rows + coil sum) against torch.fft and against the route through the full spectrum; every planned column
    length, batched and shared kernels and smaps, with and without smaps."""
    torch.manual_seed(3)
    dt = torch.complex64
    assert eng_fft.toeplitz_fused_available(dt, N, K)
    pad = [0, K[1] - N[1], 0, K[0] - N[0]]
    crop = (slice(None), slice(None), slice(0, N[0]), slice(0, N[1]))
    for B, C in ((1, 1), (2, 3), (1, 16)):
        image = torch.randn((B, 1) + N, dtype=dt, device=DEV)
        for smaps in (torch.randn((1, C) + N, dtype=dt, device=DEV), torch.randn((B, C) + N, dtype=dt, device=DEV)):
            for kern in (torch.randn(K, dtype=dt, device=DEV), torch.randn((B,) + K, dtype=dt, device=DEV)):
                kb = kern if kern.ndim == 2 else kern.unsqueeze(1)
                spec = torch.fft.fft2(torch.nn.functional.pad(image * smaps, pad)) * kb
                want = torch.sum(torch.fft.ifft2(spec, norm="forward")[crop] * smaps.conj(), 1, keepdim=True) * 0.5
                got = eng_fft.fused_toeplitz(image, kern, smaps, 0.5)
                assert rel_l2(host(got), host(want)) <= 2e-6
                two = eng_fft.fused_fft_adjoint(eng_fft.fused_fft_forward(image, K, smaps, None, 1.0), N, smaps, None,
                                                0.5, kernel=kern)
                assert rel_l2(host(got), host(two)) <= 1e-6
        multi = torch.randn((B, C) + N, dtype=dt, device=DEV)
        kern = torch.randn(K, dtype=dt, device=DEV)
        want = torch.fft.ifft2(torch.fft.fft2(torch.nn.functional.pad(multi, pad)) * kern, norm="forward")[crop]
        assert rel_l2(host(eng_fft.fused_toeplitz(multi, kern, None, 1.0)), host(want)) <= 2e-6


def test_toeplitz_three_pass_dispatch_and_fallbacks():
    """ToepNufft takes the three-pass route where it applies and the route through the full spectrum elsewhere
    (3-D, unplanned or too long column lengths, grids under twice the image); both agree; the C entry refuses what
    it cannot do with B2N_E_UNSUPPORTED."""
    from torchkbnufft_b200._autograd import nufft as auto_nufft
    torch.manual_seed(4)
    dt = torch.complex64
    assert not eng_fft.toeplitz_fused_available(dt, (8, 8, 8), (16, 16, 16))
    assert not eng_fft.toeplitz_fused_available(dt, (1024, 8), (2048, 16))   # spectrum + exchange buffer > 227 KB
    assert not eng_fft.toeplitz_fused_available(dt, (30, 30), (60, 60))      # run-time column length
    assert not eng_fft.toeplitz_fused_available(dt, (40, 32), (64, 64))      # grid < 2 x image
    assert not eng_fft.toeplitz_fused_available(torch.complex128, (32, 32), (64, 64))
    image = torch.randn((1, 1, 40, 32), dtype=dt, device=DEV)
    kern = torch.randn((64, 64), dtype=dt, device=DEV)
    with pytest.raises(RuntimeError, match="Toeplitz"):
        eng_fft.fused_toeplitz(image, kern, None, 1.0)
    toep = tkbn.ToepNufft()
    im_size = (96, 160)
    om = torch.rand((2, 3000), device=DEV) * 2 * np.pi - np.pi
    kern = tkbn.calc_toeplitz_kernel(om, im_size)
    x = torch.randn((2, 1) + im_size, dtype=dt, device=DEV, requires_grad=True)
    s = torch.randn((1, 16) + im_size, dtype=dt, device=DEV)
    res = {}
    try:
        for fuse in (True, False):
            auto_nufft.fuse_toeplitz_columns = fuse
            before = _lib.load().b2n_launch_count()
            y = toep(x, kern, smaps=s)
            res[fuse] = (host(y), _lib.load().b2n_launch_count() - before)
            (g,) = torch.autograd.grad(y.abs().pow(2).sum(), x)
            res[fuse] += (host(g),)
    finally:
        auto_nufft.fuse_toeplitz_columns = True
    assert res[True][1] == 3 and res[False][1] == 4  # kernels of the library per apply
    assert rel_l2(res[True][0], res[False][0]) <= 1e-6 and rel_l2(res[True][2], res[False][2]) <= 1e-6


def test_dependent_launches_respect_producers_and_buffer_reuse():
    """Programmatic dependent launch must not let a kernel of the engine run ahead of the work it depends on: inputs
    produced by the torch kernel right before the call (k-space for the spread, which starts under the grid zeroing;
    the image for the first FFT pass), the forward grid's memory reused for the adjoint grid, and inputs that arrive
    through an event from a copy stream.  Back-to-back unsynchronised calls must equal the synchronised,
    plain-stream-order results."""
    torch.manual_seed(6)
    lib = _lib.load()
    wl_im, C, M = (160, 160), 16, 40000
    nu = tkbn.KbNufft(im_size=wl_im, dtype=torch.complex64).to(DEV)
    na = tkbn.KbNufftAdjoint(im_size=wl_im, dtype=torch.complex64).to(DEV)
    om = (torch.rand((2, M), device=DEV) * 2 - 1) * np.pi
    s = torch.randn((1, C) + wl_im, dtype=torch.complex64, device=DEV)
    x0 = torch.randn((1, 1) + wl_im, dtype=torch.complex64, device=DEV)
    scales = [0.5 + 0.25 * i for i in range(12)]
    want = []
    try:
        lib.b2n_set_option(_lib.OPT_PDL, 0)
        for a in scales:
            torch.cuda.synchronize()
            k = nu(x0 * a, om, smaps=s)
            torch.cuda.synchronize()
            y = na(k * (1.0 + a), om, smaps=s)
            torch.cuda.synchronize()
            want.append((host(k), host(y)))
    finally:
        lib.b2n_set_option(_lib.OPT_PDL, 1)
    got = []
    for a in scales:  # no synchronisation anywhere: every input comes straight out of the preceding torch kernel
        k = nu(x0 * a, om, smaps=s)
        y = na(k * (1.0 + a), om, smaps=s)
        got.append((k, y))
    for (k, y), (wk, wy) in zip(got, want):
        assert np.array_equal(host(k), wk)
        assert rel_l2(host(y), wy) <= 1e-6
    # inputs uploaded on a copy stream, handed over by an event
    copy = torch.cuda.Stream()
    hx = [(x0 * a).cpu().pin_memory() for a in scales]
    dx = [torch.empty_like(x0) for _ in scales]
    evs = [torch.cuda.Event() for _ in scales]
    outs = []
    for i in range(len(scales)):
        with torch.cuda.stream(copy):
            dx[i].copy_(hx[i], non_blocking=True)
            evs[i].record(copy)
        torch.cuda.current_stream().wait_event(evs[i])
        outs.append(nu(dx[i], om, smaps=s))
    for k, (wk, _) in zip(outs, want):
        assert np.array_equal(host(k), wk)


@pytest.mark.parametrize("N, K, C", [((320, 320), (640, 640), 16), ((200, 511), (448, 1024), 16),
                                     ((200, 510), (448, 1024), 16), ((31, 29), (64, 64), 3)])
def test_fft_prefetch_and_pdl_options_do_not_change_results(N, K, C):
    """B2N_OPT_FFT_PREFETCH (next-wave L2 prefetch; odd row lengths skip the bulk form) and B2N_OPT_PDL only move
    memory traffic and launch timing: outputs are bit-identical with them on and off, over several waves of CTAs."""
    torch.manual_seed(5)
    dt = torch.complex64
    lib = _lib.load()
    image = torch.randn((1, 1) + N, dtype=dt, device=DEV)
    smaps = torch.randn((1, C) + N, dtype=dt, device=DEV)
    grid = torch.randn((1, C) + K, dtype=dt, device=DEV)
    kern = torch.randn(K, dtype=dt, device=DEV)
    assert lib.b2n_get_option(_lib.OPT_FFT_PREFETCH) == 19 and lib.b2n_get_option(_lib.OPT_PDL) == 1
    res = []
    try:
        for prefetch, pdl in ((0, 0), (63, 1), (19, 1)):  # off; every pass at any size; default
            lib.b2n_set_option(_lib.OPT_FFT_PREFETCH, prefetch)
            lib.b2n_set_option(_lib.OPT_PDL, pdl)
            res.append((host(eng_fft.fused_fft_forward(image, K, smaps, None, 1.0)),
                        host(eng_fft.fused_fft_adjoint(grid, N, smaps, None, 1.0)),
                        host(eng_fft.fused_fft_adjoint(grid, N, smaps, None, 1.0, kernel=kern)),
                        host(eng_fft.fused_toeplitz(image, kern, smaps, 1.0))))
    finally:
        lib.b2n_set_option(_lib.OPT_FFT_PREFETCH, 19)
        lib.b2n_set_option(_lib.OPT_PDL, 1)
    for other in res[1:]:
        for a, b in zip(res[0], other):
            assert np.array_equal(a, b)


@pytest.mark.parametrize("N, K, C", [((320, 320), (640, 640), 16), ((200, 511), (448, 1024), 16),
                                     ((31, 29), (64, 64), 3), ((20, 33, 48), (64, 64, 128), 4),
                                     ((96, 90), (192, 192), 40), ((64, 35), (128, 70), 5)])
def test_streamed_fft_passes_do_not_change_results(N, K, C):
    """B2N_OPT_FFT_STREAM (persistent column passes whose CTAs keep the operands of their next tile in flight with
    asynchronous copies) runs the same butterflies in the same order: outputs are bit-identical with the classic
    one-tile-per-CTA passes, over several tiles per CTA, partial column blocks, odd inner extents (classic route)
    and 3-D grids."""
    torch.manual_seed(7)
    dt = torch.complex64
    lib = _lib.load()
    image = torch.randn((2, 1) + N, dtype=dt, device=DEV)
    smaps = torch.randn((1, C) + N, dtype=dt, device=DEV)
    grid = torch.randn((2, C) + K, dtype=dt, device=DEV)
    assert lib.b2n_get_option(_lib.OPT_FFT_STREAM) == 1
    res = []
    try:
        for mask in (0, 1, 3, 51):
            lib.b2n_set_option(_lib.OPT_FFT_STREAM, mask)
            res.append((host(eng_fft.fused_fft_forward(image, K, smaps, None, 1.0)),
                        host(eng_fft.fused_fft_adjoint(grid, N, smaps, None, 1.0))))
    finally:
        lib.b2n_set_option(_lib.OPT_FFT_STREAM, 1)
    for other in res[1:]:
        for a, b in zip(res[0], other):
            assert np.array_equal(a, b)


def test_fast_fft_plans_agree_with_runtime_passes():
    """B2N_OPT_FAST_FFT on/off must give the same transform (different kernels, same maths)."""
    torch.manual_seed(2)
    dt = torch.complex64
    lib = _lib.load()
    N, K = (100, 320), (256, 640)
    image = torch.randn((1, 1) + N, dtype=dt, device=DEV)
    smaps = torch.randn((1, 5) + N, dtype=dt, device=DEV)
    grid = torch.randn((1, 5) + K, dtype=dt, device=DEV)
    res = {}
    try:
        for on in (1, 0):
            lib.b2n_set_option(_lib.OPT_FAST_FFT, on)
            assert lib.b2n_get_option(_lib.OPT_FAST_FFT) == on
            res[on] = (host(eng_fft.fused_fft_forward(image, K, smaps, None, 1.0)),
                       host(eng_fft.fused_fft_adjoint(grid, N, smaps, None, 1.0)))
    finally:
        lib.b2n_set_option(_lib.OPT_FAST_FFT, 1)
    assert rel_l2(res[1][0], res[0][0]) <= 1e-6 and rel_l2(res[1][1], res[0][1]) <= 1e-6


@pytest.mark.parametrize("grid_size", [(16, 16, 16), (24, 20, 28), (10, 12, 14), (13, 21, 9), (32, 8, 40)])
@pytest.mark.parametrize("B, C, batched", [(1, 1, False), (2, 3, False), (1, 4, False), (1, 5, False), (2, 8, True)])
def test_tiled_3d_kernels_match_generic_and_oracle(grid_size, B, C, batched):
    """3-D shared-memory tiled kernels (complex64, J=6) against the generic kernels and the
    oracle: partial tiles, grids smaller than a tile, wrap-around, dense sub-problems, coil tails."""
    rng = np.random.default_rng(hash((grid_size, B, C)) & 0xFFFF)
    im_size = tuple(max(2, k // 2) for k in grid_size)
    ob = tkbn.KbInterp(im_size=im_size, grid_size=grid_size, dtype=torch.complex64).to(DEV)
    M = 2500
    shape = (B, 3, M) if batched else (3, M)
    omega = rng.uniform(-np.pi, np.pi, size=shape)
    omega[..., : M // 3] *= 0.08  # dense clump around k = 0
    omega = np.ascontiguousarray(omega.astype(np.float32))
    grid = workloads.complex_normal(rng, (B, C) + tuple(grid_size))
    kdata = workloads.complex_normal(rng, (B, C, M))
    args = (ob.tables, ob.n_shift, ob.numpoints, ob.table_oversamp)
    tables = [host(t) for t in ob.tables]
    J, L, ns = ob.numpoints.tolist(), ob.table_oversamp.tolist(), host(ob.n_shift)
    want_f = orc.table_interp(grid, omega, tables, ns, J, L, nthreads=4)
    want_a = orc.table_interp_adjoint(kdata, omega, tables, ns, J, L, grid_size, nthreads=4)
    res = {}
    try:
        for tiled in ("force", False):
            tkbn.set_tiled_kernels(tiled)
            res[tiled] = (host(eng_interp.table_interp(dev(grid), dev(omega), *args)),
                          host(eng_interp.table_interp_adjoint(dev(kdata), dev(omega), *args, None, ob.grid_size,
                                                               mode="atomic")))
    finally:
        tkbn.set_tiled_kernels(True)
    for tiled in ("force", False):
        assert rel_l2(res[tiled][0], want_f) <= 1e-5, f"forward tiled={tiled}"
        assert rel_l2(res[tiled][1], want_a) <= 1e-4, f"adjoint tiled={tiled}"
    assert rel_l2(res["force"][0], res[False][0]) <= 2e-6
    assert rel_l2(res["force"][1], res[False][1]) <= 2e-6


def test_cuda_graph_capture_of_forward_adjoint_pair():
    """with the trajectory plan cached, a forward+adjoint SENSE pair is pure
    stream work (no synchronisation, no host reads) and can be captured in a CUDA graph and replayed
    on new input contents."""
    wl = workloads.Workload("g", (64, 64), 5, 1, 40, 128, "golden", "graph capture case")
    image, smaps, kdata, omega = workloads.make_inputs(wl, seed=2)
    x, s, om = dev(image), dev(smaps), dev(omega)
    nu = tkbn.KbNufft(im_size=wl.im_size, dtype=torch.complex64).to(DEV)
    na = tkbn.KbNufftAdjoint(im_size=wl.im_size, dtype=torch.complex64).to(DEV)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):  # warm-up on the capture stream: plan, twiddles, function attributes
        for _ in range(2):
            na(nu(x, om, smaps=s), om, smaps=s)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        k = nu(x, om, smaps=s)
        im = na(k, om, smaps=s)
    for seed in (7, 8):
        x.copy_(dev(workloads.complex_normal(np.random.default_rng(seed), image.shape)))
        graph.replay()
        torch.cuda.synchronize()
        want_k = nu(x, om, smaps=s)
        want_im = na(want_k, om, smaps=s)
        assert rel_l2(host(k), host(want_k)) <= 1e-6
        assert rel_l2(host(im), host(want_im)) <= 1e-5


@pytest.mark.parametrize("grid_size, B, C, batched", [((64, 64), 1, 16, False), ((70, 66), 2, 3, False),
                                                      ((21, 40), 1, 1, False), ((96, 50), 3, 5, True),
                                                      ((37, 128), 1, 32, False)])
def test_ordered_tiled_adjoint_is_deterministic_and_matches_oracle(grid_size, B, C, batched):
    """'sorted' mode on the tiled kernels: per-sub-problem scratch tiles + fixed-order merge.  Bit-identical run to
    run and across plan rebuilds; equal to the oracle and to the per-cell gather within tolerance.  Grids with a
    partial last tile narrower than the halo (70, 66, 37), wrap-around, dense clumps (many chunks per tile),
    batched trajectories and coil counts on every CC variant."""
    rng = np.random.default_rng(hash((grid_size, B, C)) & 0xFFFF)
    im_size = tuple(max(2, k // 2) for k in grid_size)
    ob = tkbn.KbInterpAdjoint(im_size=im_size, grid_size=grid_size, dtype=torch.complex64).to(DEV)
    M = 3000
    shape = (B, 2, M) if batched else (2, M)
    omega = rng.uniform(-np.pi, np.pi, size=shape)
    omega[..., : M // 3] *= 0.05  # dense clump around k = 0: several sub-problems per tile
    omega = np.ascontiguousarray(omega.astype(np.float32))
    kdata = workloads.complex_normal(rng, (B, C, M))
    args = (ob.tables, ob.n_shift, ob.numpoints, ob.table_oversamp)
    tables = [host(t) for t in ob.tables]
    J, L, ns = ob.numpoints.tolist(), ob.table_oversamp.tolist(), host(ob.n_shift)
    want = orc.table_interp_adjoint(kdata, omega, tables, ns, J, L, grid_size, nthreads=4)
    y, om = dev(kdata), dev(omega)
    lib = _lib.load()
    nbytes = ctypes.c_size_t(0)
    runs = []
    for rep in range(3):
        if rep == 2:
            om = om.clone()  # a new trajectory tensor: the plan is rebuilt
        runs.append(eng_interp.table_interp_adjoint(y, om, *args, None, ob.grid_size, mode="sorted"))
    assert torch.equal(runs[0], runs[1]) and torch.equal(runs[0], runs[2])
    assert rel_l2(host(runs[0]), want) <= 1e-5
    try:  # the per-cell gather (tiled kernels off) is the other deterministic implementation
        tkbn.set_tiled_kernels(False)
        gather = eng_interp.table_interp_adjoint(y, om, *args, None, ob.grid_size, mode="sorted")
    finally:
        tkbn.set_tiled_kernels(True)
    assert rel_l2(host(gather), host(runs[0])) <= 2e-6
    atomic = eng_interp.table_interp_adjoint(y, om, *args, None, ob.grid_size, mode="atomic")
    assert rel_l2(host(atomic), host(runs[0])) <= 2e-6


@pytest.mark.parametrize("grid_size, B, C, batched", [((16, 16, 16), 1, 4, False), ((24, 20, 28), 2, 3, False),
                                                      ((13, 21, 17), 1, 1, False), ((32, 14, 40), 2, 8, True)])
def test_ordered_tiled_adjoint_3d(grid_size, B, C, batched):
    """3-D twin of the deterministic tiled adjoint: scratch tiles + 3x3x3 fixed-order merge."""
    rng = np.random.default_rng(hash((grid_size, B, C)) & 0xFFFF)
    im_size = tuple(max(2, k // 2) for k in grid_size)
    ob = tkbn.KbInterpAdjoint(im_size=im_size, grid_size=grid_size, dtype=torch.complex64).to(DEV)
    M = 2500
    shape = (B, 3, M) if batched else (3, M)
    omega = rng.uniform(-np.pi, np.pi, size=shape)
    omega[..., : M // 3] *= 0.08
    omega = np.ascontiguousarray(omega.astype(np.float32))
    kdata = workloads.complex_normal(rng, (B, C, M))
    args = (ob.tables, ob.n_shift, ob.numpoints, ob.table_oversamp)
    tables = [host(t) for t in ob.tables]
    J, L, ns = ob.numpoints.tolist(), ob.table_oversamp.tolist(), host(ob.n_shift)
    want = orc.table_interp_adjoint(kdata, omega, tables, ns, J, L, grid_size, nthreads=4)
    y, om = dev(kdata), dev(omega)
    nbytes = ctypes.c_size_t(0)
    runs = [eng_interp.table_interp_adjoint(y, om, *args, None, ob.grid_size, mode="sorted") for _ in range(2)]
    runs.append(eng_interp.table_interp_adjoint(y, om.clone(), *args, None, ob.grid_size, mode="sorted"))
    assert torch.equal(runs[0], runs[1]) and torch.equal(runs[0], runs[2])
    assert rel_l2(host(runs[0]), want) <= 1e-5
    atomic = eng_interp.table_interp_adjoint(y, om, *args, None, ob.grid_size, mode="atomic")
    assert rel_l2(host(atomic), host(runs[0])) <= 2e-6


@pytest.mark.parametrize("grid_size, B, C", [((64, 64), 1, 16), ((70, 52), 2, 32)])
def test_channel_last_tiled_spread_matches_coil_major(grid_size, B, C):
    """k_adj_tiled_cl_2d (channel-last grid, 16-coil lines, 128-bit read-modify-write) against the coil-major
    tiled kernel and the oracle."""
    rng = np.random.default_rng(C + grid_size[0])
    im_size = tuple(k // 2 for k in grid_size)
    ob = tkbn.KbInterpAdjoint(im_size=im_size, grid_size=grid_size, dtype=torch.complex64).to(DEV)
    M = 3000
    omega = rng.uniform(-np.pi, np.pi, size=(2, M))
    omega[:, : M // 3] *= 0.05
    omega = np.ascontiguousarray(omega.astype(np.float32))
    kdata = workloads.complex_normal(rng, (B, C, M))
    args = (ob.tables, ob.n_shift, ob.numpoints, ob.table_oversamp)
    y, om = dev(kdata), dev(omega)
    cm = eng_interp.table_interp_adjoint(y, om, *args, None, ob.grid_size, mode="atomic")
    cl = eng_interp.table_interp_adjoint(y, om, *args, None, ob.grid_size, mode="atomic", layout=_lib.CHANNEL_LAST)
    assert cl.shape == (B,) + tuple(grid_size) + (C,)
    assert rel_l2(host(cl.movedim(-1, 1)), host(cm)) <= 2e-6
    tables = [host(t) for t in ob.tables]
    want = orc.table_interp_adjoint(kdata, omega, tables, host(ob.n_shift), ob.numpoints.tolist(),
                                    ob.table_oversamp.tolist(), grid_size, nthreads=4)
    assert rel_l2(host(cl.movedim(-1, 1)), want) <= 1e-5


@pytest.mark.parametrize("grid_size", [(48, 40), (20, 24, 32)])
def test_modified_tables_take_the_complex_weight_kernels(grid_size):
    """The owner-tile spread needs tables of the reference's form (real kernel x linear phase).  A module whose table
    buffer was edited must not get visit lists -- its adjoint runs on the complex-weight kernels -- and still agrees
    with the oracle evaluated with the SAME edited tables; the untouched module keeps the fast path."""
    rng = np.random.default_rng(len(grid_size))
    d = len(grid_size)
    im_size = tuple(k // 2 for k in grid_size)
    M, C = 2500, 4
    omega = np.ascontiguousarray(rng.uniform(-np.pi, np.pi, size=(d, M)).astype(np.float32))
    kdata = workloads.complex_normal(rng, (1, C, M))
    from torchkbnufft_b200._nufft import plan as P

    for edit in (False, True):
        ob = tkbn.KbInterpAdjoint(im_size=im_size, grid_size=grid_size, dtype=torch.complex64).to(DEV)
        if edit:
            with torch.no_grad():
                ob.table_1[1000:1100] *= torch.polar(torch.tensor(1.0), torch.tensor(0.3)).to(DEV)  # foreign phase
        got = host(ob(dev(kdata), dev(omega)))
        geo = P.get_geometry(ob.tables, ob.n_shift, ob.numpoints, ob.table_oversamp, ob.grid_size)
        plan = P.get_plan(geo, dev(omega))
        assert bool(plan.struct.own_tile) == (not edit)
        tables = [host(t) for t in ob.tables]
        want = orc.table_interp_adjoint(kdata, omega, tables, host(ob.n_shift), ob.numpoints.tolist(),
                                        ob.table_oversamp.tolist(), grid_size)
        assert rel_l2(got, want) <= 1e-4, edit
