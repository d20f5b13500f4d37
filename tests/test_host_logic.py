"""Host-side logic of the product package on the CPU: the Python layers (modules,
functional API, autograd wiring, Toeplitz kernel assembly, density compensation)
run with the engine entry points monkey-patched by the oracle
(tests/cpu_engine_shim.py) and are compared with the reference's outputs stored in
tests/golden/ref_cases.npz.  What runs on the GPU is tested in test_gpu_*.py."""
import numpy as np
import pytest
import torch

import torchkbnufft_b200 as tkbn
from conftest import module_kwargs, rel_l2
from cpu_engine_shim import oracle_engine
from golden_cases import CASES, case_inputs

PRECS = {"c64": (np.complex64, torch.complex64, 2e-6), "c128": (np.complex128, torch.complex128, 1e-13)}


@pytest.fixture(autouse=True)
def _shim():
    with oracle_engine():
        yield


@pytest.mark.parametrize("prec", ["c64", "c128"])
@pytest.mark.parametrize("name", list(CASES))
def test_modules_match_reference(name, prec, ref_cases):
    case = CASES[name]
    cd, td, tol = PRECS[prec]
    inp = case_inputs(case, cd)
    kw = module_kwargs(case, td)
    T = lambda k: torch.from_numpy(inp[k])
    key = f"{name}_{prec}_"
    assert rel_l2(tkbn.KbInterp(**kw)(T("grid"), T("omega")).numpy(), ref_cases[key + "interp"]) <= tol
    assert rel_l2(tkbn.KbInterpAdjoint(**kw)(T("kdata"), T("omega")).numpy(), ref_cases[key + "interp_adj"]) <= tol
    nu, na = tkbn.KbNufft(**kw), tkbn.KbNufftAdjoint(**kw)
    for norm in (None, "ortho"):
        tag = "ortho" if norm else "none"
        out = nu(T("image"), T("omega"), smaps=T("smaps"), norm=norm)
        assert rel_l2(out.numpy(), ref_cases[key + f"sense_fwd_{tag}"]) <= tol
        out = na(T("kdata"), T("omega"), smaps=T("smaps"), norm=norm)
        assert out.shape[1] == 1
        assert rel_l2(out.numpy(), ref_cases[key + f"sense_adj_{tag}"]) <= tol
    assert rel_l2(nu(T("image_multi"), T("omega")).numpy(), ref_cases[key + "nufft_fwd_nosmap"]) <= tol
    opts = dict(grid_size=case.get("grid_size"), numpoints=case.get("numpoints", 6),
                table_oversamp=case.get("table_oversamp", 2 ** 10))
    if case.get("toep", True):
        for norm in (None, "ortho"):
            tag = "ortho" if norm else "none"
            kern = tkbn.calc_toeplitz_kernel(T("omega"), case["im_size"], norm=norm, **opts)
            assert rel_l2(kern.numpy(), ref_cases[key + f"toep_kernel_{tag}"]) <= tol
            got = tkbn.ToepNufft()(T("image"), kern, smaps=T("smaps"), norm=norm).numpy()
            ref = ref_cases[key + f"toep_apply_{tag}"]
            # deliberate deviation: a single smaps/kernel broadcasts over the batch, where the
            # reference's zip() silently truncates it (modules/kbnufft.py:472)
            assert got.shape[0] == inp["image"].shape[0]
            assert rel_l2(got[: ref.shape[0]], ref) <= tol
        kern = tkbn.calc_toeplitz_kernel(T("omega"), case["im_size"], weights=T("weights"), norm="ortho", **opts)
        assert rel_l2(kern.numpy(), ref_cases[key + "toep_kernel_weighted"]) <= tol
    dcomp = tkbn.calc_density_compensation_function(T("omega"), case["im_size"], num_iterations=3,
                                                    n_shift=case.get("n_shift"), **opts)
    assert dcomp.shape == (inp["omega"].shape[0] if inp["omega"].ndim == 3 else 1, 1, case["M"])
    assert rel_l2(dcomp.numpy(), ref_cases[key + "dcomp"]) <= tol


def test_real_view_inputs_match_complex():
    case = CASES["d2"]
    inp = case_inputs(case, np.complex128)
    kw = module_kwargs(case, torch.complex128)
    T = lambda k: torch.from_numpy(inp[k])
    nu, na = tkbn.KbNufft(**kw), tkbn.KbNufftAdjoint(**kw)
    vr = torch.view_as_real
    a = nu(T("image"), T("omega"), smaps=T("smaps"))
    b = nu(vr(T("image")), T("omega"), smaps=vr(T("smaps")))
    assert not b.is_complex() and b.shape[-1] == 2 and torch.equal(vr(a), b)
    a = na(T("kdata"), T("omega"), smaps=T("smaps"), norm="ortho")
    b = na(vr(T("kdata")), T("omega"), smaps=vr(T("smaps")), norm="ortho")
    assert torch.equal(vr(a), b)
    a = tkbn.KbInterp(**kw)(T("grid"), T("omega"))
    assert torch.equal(vr(a), tkbn.KbInterp(**kw)(vr(T("grid")), T("omega")))
    kern = tkbn.calc_toeplitz_kernel(T("omega"), case["im_size"])
    a = tkbn.ToepNufft()(T("image"), kern, smaps=T("smaps"))
    assert torch.equal(vr(a), tkbn.ToepNufft()(vr(T("image")), vr(kern), smaps=vr(T("smaps"))))


def test_error_types_match_reference():
    kw = dict(im_size=(8, 8), dtype=torch.complex128)
    nu, na, interp = tkbn.KbNufft(**kw), tkbn.KbNufftAdjoint(**kw), tkbn.KbInterp(**kw)
    image = torch.randn(2, 1, 8, 8, dtype=torch.complex128)
    omega = torch.rand(2, 12, dtype=torch.float64)
    with pytest.raises(ValueError, match="omega must have 2 or 3 dimensions"):
        nu(image, omega[0])
    with pytest.raises(ValueError, match="batch dimension must match"):
        nu(image, omega[None].repeat(3, 1, 1))
    with pytest.raises(ValueError, match="last dimension must be size 2"):
        nu(torch.randn(2, 1, 8, 8, dtype=torch.float64), omega)
    with pytest.raises(TypeError, match="does not match smaps dtype"):
        nu(image, omega, smaps=torch.randn(1, 3, 8, 8, dtype=torch.complex64))
    with pytest.raises(ValueError, match="Only option for norm"):
        nu(image, omega, norm="backward")
    with pytest.raises(ValueError, match="Only option for norm"):
        na(torch.randn(2, 1, 12, dtype=torch.complex128), omega, norm="x")
    with pytest.raises(TypeError, match="same dtype"):
        tkbn.ToepNufft()(image, torch.randn(16, 16, dtype=torch.complex64))
    with pytest.raises(ValueError, match="same batch size"):
        tkbn.ToepNufft()(image, torch.randn(3, 16, 16, dtype=torch.complex128))
    with pytest.raises(ValueError, match="Unrecognized k-space shape"):
        tkbn.calc_toeplitz_kernel(omega[0], (8, 8))
    with pytest.raises(ValueError, match="ktraj must have 2 or 3 dimensions"):
        tkbn.calc_density_compensation_function(omega[0], (8, 8))
    # (1, d, M) broadcasts a single trajectory over the batch
    out = interp(torch.randn(2, 1, 16, 16, dtype=torch.complex128), omega[None])
    assert out.shape == (2, 1, 12)


@pytest.mark.parametrize("name", ["d1", "d2", "d3", "d2_batched"])
def test_adjointness_and_autograd_wiring(name):
    """<A x, y> == <x, A^H y> and d/dx 0.5*||A x||^2 == A^H A x through the autograd
    Functions (the reference's own property tests, tests/conftest.py:40-87)."""
    case = CASES[name]
    inp = case_inputs(case, np.complex128)
    kw = module_kwargs(case, torch.complex128)
    T = lambda k: torch.from_numpy(inp[k])
    for fw, ad, x, y, extra in (
        (tkbn.KbInterp(**kw), tkbn.KbInterpAdjoint(**kw), T("grid"), T("kdata"), {}),
        (tkbn.KbNufft(**kw), tkbn.KbNufftAdjoint(**kw), T("image"), T("kdata"), dict(smaps=T("smaps"), norm="ortho")),
        (tkbn.KbNufft(**kw), tkbn.KbNufftAdjoint(**kw), T("image_multi"), T("kdata"), {}),
    ):
        om = T("omega")
        lhs = tkbn.inner_product(fw(x, om, **extra), y)
        rhs = tkbn.inner_product(x, ad(y, om, **extra))
        assert torch.allclose(lhs, rhs)
        xg = x.clone().requires_grad_(True)
        yg = y.clone().requires_grad_(True)
        fx = fw(xg, om, **extra)
        (torch.abs(fx) ** 2 / 2).sum().backward()
        assert torch.allclose(xg.grad, ad(fx.detach(), om, **extra))
        ay = ad(yg, om, **extra)
        (torch.abs(ay) ** 2 / 2).sum().backward()
        assert torch.allclose(yg.grad, fw(ay.detach(), om, **extra))


def test_smaps_gradient_path_matches_plain_torch():
    case = CASES["d2"]
    inp = case_inputs(case, np.complex128)
    kw = module_kwargs(case, torch.complex128)
    T = lambda k: torch.from_numpy(inp[k])
    nu, na = tkbn.KbNufft(**kw), tkbn.KbNufftAdjoint(**kw)
    smaps = T("smaps").clone().requires_grad_(True)
    out = nu(T("image"), T("omega"), smaps=smaps)
    (torch.abs(out) ** 2).sum().backward()
    g1 = smaps.grad.clone()
    smaps2 = T("smaps").clone().requires_grad_(True)
    out2 = nu(T("image") * smaps2, T("omega"))
    (torch.abs(out2) ** 2).sum().backward()
    assert torch.allclose(g1, smaps2.grad)
    smaps3 = T("smaps").clone().requires_grad_(True)
    img = na(T("kdata"), T("omega"), smaps=smaps3)
    (torch.abs(img) ** 2).sum().backward()
    assert smaps3.grad is not None and torch.isfinite(smaps3.grad.abs()).all()


def test_toeplitz_matches_normal_operator_and_is_differentiable():
    case = CASES["d2_radial"]
    inp = case_inputs(case, np.complex128)
    kw = module_kwargs(case, torch.complex128)
    T = lambda k: torch.from_numpy(inp[k])
    nu, na, toep = tkbn.KbNufft(**kw), tkbn.KbNufftAdjoint(**kw), tkbn.ToepNufft()
    kern = tkbn.calc_toeplitz_kernel(T("omega"), case["im_size"], norm="ortho")
    fbn = na(nu(T("image"), T("omega"), smaps=T("smaps"), norm="ortho"), T("omega"), smaps=T("smaps"), norm="ortho")
    fbt = toep(T("image"), kern, smaps=T("smaps"), norm="ortho")
    assert torch.norm(fbn - fbt) / torch.norm(fbn) < 1e-4  # the reference's tolerance, tests/test_toep.py:64
    x = T("image").clone().requires_grad_(True)
    y = toep(x, kern, smaps=T("smaps"), norm="ortho")
    w = T("image_multi")[:, :1]
    torch.real(torch.sum(y * w.conj())).backward()
    # gradient of Re<T x, w> is T^H w; T is Hermitian for a Hermitian-symmetric kernel
    expect = toep(w, kern, smaps=T("smaps"), norm="ortho")
    assert torch.allclose(x.grad, expect, rtol=1e-8, atol=1e-10)


def test_batched_dcomp_equals_loop():
    torch.manual_seed(0)
    ktraj = torch.rand(3, 2, 40, dtype=torch.float64) * 2 * np.pi - np.pi
    batched = tkbn.calc_density_compensation_function(ktraj, (10, 8), num_iterations=4)
    looped = torch.cat([tkbn.calc_density_compensation_function(k, (10, 8), num_iterations=4) for k in ktraj])
    assert torch.allclose(batched, looped)


def test_real_table_form_accepts_reference_tables_and_rejects_modified_ones():
    """The owner-tile spread works with the REAL Kaiser-Bessel kernel when the tables are r(x) exp(-1j p x) with
    p = pi (N - 1) / K (what the reference builds, torchkbnufft/_nufft/utils.py:160-204); the host check that decides
    it must recover p exactly, reproduce the table to float32 rounding, and refuse anything else."""
    import numpy as np

    from torchkbnufft_b200._nufft import plan as P
    from torchkbnufft_b200._nufft import utils

    for N, K in ((320, 640), (75, 150), (128, 200), (9, 16)):
        tabs = utils.build_table((N,), (K,), (6,), (1024,), (0,), (2.34 * 6,))
        for dt, tol in ((torch.complex64, 1e-7), (torch.complex128, 1e-14)):
            got = P.real_table_form([tabs[0].to(dt)], [6], [1024], [K])
            assert got is not None, (N, K, dt)
            real, slope = got[0]
            assert abs(slope * K / np.pi - (N - 1)) < 1e-9
            x = np.arange(6 * 1024 + 1) / 1024 - 3
            rebuilt = real.numpy().astype(np.float64) * np.exp(-1j * slope * x)
            want = tabs[0].to(dt).numpy().astype(np.complex128)
            assert np.abs(rebuilt - want).max() <= tol * np.abs(want).max()
            assert float(real.min()) >= 0.0  # a Kaiser-Bessel kernel is non-negative
    tab = utils.build_table((64,), (128,), (6,), (1024,), (0,), (2.34 * 6,))[0].to(torch.complex64)
    bad = tab.clone()
    bad[3000] = bad[3000] * torch.tensor(np.exp(0.01j), dtype=torch.complex64)  # one entry with a foreign phase
    assert P.real_table_form([bad], [6], [1024], [128]) is None
    assert P.real_table_form([torch.randn(6145, dtype=torch.complex64)], [6], [1024], [128]) is None
    assert P.real_table_form([tab[:-5]], [6], [1024], [128]) is None  # wrong length
