"""Host-boundary robustness of the CUDA engine on a B200: everything that reaches a raw ``data_ptr()``.

Lazy conjugate views, trajectory dtypes that differ from the table dtype, operands of the wrong
shape / dtype, double backward, plan-cache invalidation and use from a second stream.  The
reference's ATen ops handle all of these (``torchkbnufft/_nufft/interp.py:171-203``,
``_nufft/fft.py:72``, ``modules/kbnufft.py:183``, ``:405``); the engine must agree or raise.
Tolerances: complex64 rel-L2 <= 1e-5 forward, <= 1e-4 adjoint (north_star).
"""
import os

import numpy as np
import pytest
import torch

import torchkbnufft_b200 as tkbn
from conftest import rel_l2
from torchkbnufft_b200 import _lib
from torchkbnufft_b200._nufft import fft as eng_fft
from torchkbnufft_b200._nufft import interp as eng_interp

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
IM, C, M = (32, 48), 4, 700  # grid 64 x 96: own FFT plans
GRID = (64, 96)


@pytest.fixture(autouse=True)
def _native_library_is_loaded():
    assert os.path.exists(_lib.LIB_PATH), "libb200nufft.so missing: GPU tests must run the native engine"
    _lib.load()
    yield


def host(t):
    return t.detach().cpu().numpy()


def _setup(dtype=torch.complex64, seed=3):
    g = torch.Generator(device="cpu").manual_seed(seed)
    rdt = torch.float32 if dtype == torch.complex64 else torch.float64
    image = torch.randn((2, 1) + IM, dtype=dtype, generator=g).to(DEV)
    smaps = torch.randn((1, C) + IM, dtype=dtype, generator=g).to(DEV)
    kdata = torch.randn((2, C, M), dtype=dtype, generator=g).to(DEV)
    omega = ((torch.rand((2, M), dtype=torch.float64, generator=g) - 0.5) * 2 * np.pi).to(rdt).to(DEV)
    nu = tkbn.KbNufft(im_size=IM, dtype=dtype).to(DEV)
    na = tkbn.KbNufftAdjoint(im_size=IM, dtype=dtype).to(DEV)
    return image, smaps, kdata, omega, nu, na


def test_lazy_conjugate_inputs_are_materialised():
    image, smaps, kdata, omega, nu, na = _setup()
    ic, sc, kc = image.conj(), smaps.conj(), kdata.conj()
    assert ic.is_conj() and ic.contiguous().is_conj()  # what a raw pointer would silently ignore
    want = nu(ic.resolve_conj(), omega, smaps=sc.resolve_conj())
    assert rel_l2(host(nu(ic, omega, smaps=sc)), host(want)) <= 1e-6
    want = na(kc.resolve_conj(), omega, smaps=sc.resolve_conj())
    assert rel_l2(host(na(kc, omega, smaps=sc)), host(want)) <= 1e-6
    # the bare interpolators and the Toeplitz operator
    ob, oa = tkbn.KbInterp(im_size=IM).to(DEV), tkbn.KbInterpAdjoint(im_size=IM).to(DEV)
    grid = torch.randn((2, C) + GRID, dtype=torch.complex64, device=DEV)
    assert rel_l2(host(ob(grid.conj(), omega)), host(ob(grid.conj().resolve_conj(), omega))) <= 1e-6
    assert rel_l2(host(oa(kc, omega)), host(oa(kc.resolve_conj(), omega))) <= 1e-6
    kern = tkbn.calc_toeplitz_kernel(omega, IM)
    toep = tkbn.ToepNufft()
    want = toep(ic.resolve_conj(), kern.conj().resolve_conj(), smaps=sc.resolve_conj())
    assert rel_l2(host(toep(ic, kern.conj(), smaps=sc)), host(want)) <= 1e-6


def test_gradients_through_a_conjugated_output():
    """Autograd hands conj-bit gradients to a custom backward when the output flows through ``.conj()``."""
    image, smaps, kdata, omega, nu, na = _setup()
    w = torch.randn((2, C, M), dtype=torch.complex64, device=DEV)
    x = image.clone().requires_grad_(True)
    (nu(x, omega, smaps=smaps).conj() * w).sum().abs().backward()
    got = x.grad.clone()
    # the same loss with the conjugation resolved by hand
    x2 = image.clone().requires_grad_(True)
    y2 = nu(x2, omega, smaps=smaps)
    (torch.complex(y2.real, -y2.imag) * w).sum().abs().backward()
    assert rel_l2(host(got), host(x2.grad)) <= 1e-5
    y = kdata.clone().requires_grad_(True)
    v = torch.randn((2, 1) + IM, dtype=torch.complex64, device=DEV)
    (na(y, omega, smaps=smaps).conj() * v).sum().abs().backward()
    y3 = kdata.clone().requires_grad_(True)
    o3 = na(y3, omega, smaps=smaps)
    (torch.complex(o3.real, -o3.imag) * v).sum().abs().backward()
    assert rel_l2(host(y.grad), host(y3.grad)) <= 1e-5


def test_trajectory_dtype_is_converted_not_reinterpreted():
    image, smaps, kdata, omega, nu, na = _setup(torch.complex64)
    om64 = omega.double()  # what torch.tensor(numpy float64 array) gives
    assert rel_l2(host(nu(image, om64, smaps=smaps)), host(nu(image, omega, smaps=smaps))) <= 1e-6
    assert rel_l2(host(na(kdata, om64, smaps=smaps)), host(na(kdata, omega, smaps=smaps))) <= 1e-6
    # the conversion is cached: the second call reuses the plan built for the first
    from torchkbnufft_b200._nufft import plan as P
    n_plans = len(P._PLAN_CACHE)
    nu(image, om64, smaps=smaps)
    assert len(P._PLAN_CACHE) == n_plans
    # complex128 module with a float32 trajectory: used to read past the end of the buffer
    image, smaps, kdata, omega, nu, na = _setup(torch.complex128)
    om32 = omega.float()
    want = nu(image, om32.double(), smaps=smaps)
    assert rel_l2(host(nu(image, om32, smaps=smaps)), host(want)) <= 1e-12
    with pytest.raises(TypeError):
        nu(image, om32.to(torch.int32), smaps=smaps)
    arr, tab = eng_interp.export_indices(om32, nu.tables, nu.n_shift, nu.numpoints, nu.table_oversamp, nu.grid_size)
    arr2, tab2 = eng_interp.export_indices(om32.double(), nu.tables, nu.n_shift, nu.numpoints, nu.table_oversamp,
                                           nu.grid_size)
    assert torch.equal(arr, arr2) and torch.equal(tab, tab2)


def test_operand_shapes_and_dtypes_are_checked_before_the_kernels():
    image, smaps, kdata, omega, nu, na = _setup()
    big = torch.randn((2, 1) + GRID, dtype=torch.complex64, device=DEV)  # image larger than im_size
    with pytest.raises(ValueError):
        nu(big, omega, smaps=None)
    with pytest.raises(ValueError):
        nu(image, omega, smaps=torch.randn((1, C, 16, 48), dtype=torch.complex64, device=DEV))
    with pytest.raises(TypeError):
        nu(image, omega, smaps=smaps.to(torch.complex128))
    with pytest.raises(TypeError):
        nu(image.to(torch.complex128), omega)
    with pytest.raises(ValueError):
        na(kdata, omega, smaps=torch.randn((1, C, 16, 48), dtype=torch.complex64, device=DEV))
    with pytest.raises(ValueError):
        na(kdata, omega, smaps=torch.randn((3, C) + IM, dtype=torch.complex64, device=DEV))
    with pytest.raises((ValueError, RuntimeError)):
        nu(image, omega.cpu(), smaps=smaps)
    grid = torch.randn((2, C) + GRID, dtype=torch.complex64, device=DEV)
    with pytest.raises(ValueError):
        eng_fft.crop_apod_coilsum(grid, IM, None, nu.scaling_coef[:, :16])
    with pytest.raises(ValueError):
        eng_fft.fused_fft_adjoint(grid, IM, None, None, 1.0, kernel=torch.ones((40, 96), dtype=torch.complex64,
                                                                             device=DEV))
    kern = tkbn.calc_toeplitz_kernel(omega, IM)
    with pytest.raises(TypeError):
        tkbn.ToepNufft()(image, kern.to(torch.complex128), smaps=smaps)


def test_double_backward_raises_instead_of_cutting_the_graph():
    image, smaps, kdata, omega, nu, na = _setup()
    x = image.clone().requires_grad_(True)
    loss = (nu(x, omega, smaps=smaps).abs() ** 2).sum()
    (g,) = torch.autograd.grad(loss, x, create_graph=True)
    with pytest.raises(RuntimeError):
        g.abs().sum().backward()


def test_plan_invalidation_and_content_mode():
    image, smaps, kdata, omega, nu, na = _setup()
    om = omega.clone()
    first = nu(image, om, smaps=smaps)
    om.data.mul_(0.5)  # bypasses the version counter: the cached plan is stale by design ...
    tkbn.invalidate_plans(om)  # ... until the caller says so
    want = nu(image, (omega * 0.5), smaps=smaps)
    assert rel_l2(host(nu(image, om, smaps=smaps)), host(want)) <= 1e-6
    assert rel_l2(host(first), host(want)) > 1e-3
    # content mode notices by itself
    tkbn.set_plan_cache_mode("content")
    try:
        om2 = omega.clone()
        nu(image, om2, smaps=smaps)
        om2.data.mul_(0.25)
        want = nu(image, omega * 0.25, smaps=smaps)
        assert rel_l2(host(nu(image, om2, smaps=smaps)), host(want)) <= 1e-6
    finally:
        tkbn.set_plan_cache_mode("version")
    tkbn.set_plan_cache_mode("off")
    try:
        assert rel_l2(host(nu(image, omega, smaps=smaps)), host(nu(image, omega.clone(), smaps=smaps))) <= 1e-6
    finally:
        tkbn.set_plan_cache_mode("version")
    with pytest.raises(ValueError):
        tkbn.set_plan_cache_mode("sometimes")


def test_plan_built_on_one_stream_is_usable_from_another():
    image, smaps, kdata, omega, nu, na = _setup(seed=5)
    tkbn.clear_caches()
    want_f = nu(image, omega.clone(), smaps=smaps)
    want_a = na(kdata, omega.clone(), smaps=smaps)
    torch.cuda.synchronize()
    om = omega.clone()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    with torch.cuda.stream(s1):
        a = nu(image, om, smaps=smaps)  # builds the plan on s1
    with torch.cuda.stream(s2):
        b = nu(image, om, smaps=smaps)  # must wait for the build, not read a half-written plan
        c = na(kdata, om, smaps=smaps)
    torch.cuda.synchronize()
    assert rel_l2(host(a), host(want_f)) <= 1e-6 and rel_l2(host(b), host(want_f)) <= 1e-6
    assert rel_l2(host(c), host(want_a)) <= 1e-5


def test_twiddles_are_not_built_inside_a_graph_capture():
    tkbn.clear_caches()
    eng_fft._TWIDDLES.clear()
    eng_fft._TWIDDLE_SETS.clear()
    eng_fft._FFT_CTX.clear()
    image, smaps, kdata, omega, nu, na = _setup()
    nu(image, omega, smaps=smaps)  # warm-up builds plan + twiddles eagerly
    torch.cuda.synchronize()
    eng_fft._TWIDDLES.clear()
    eng_fft._TWIDDLE_SETS.clear()
    eng_fft._FFT_CTX.clear()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with pytest.raises(RuntimeError, match="before CUDA-graph capture"):
            with torch.cuda.graph(g, stream=s):
                nu(image, omega, smaps=smaps)
    torch.cuda.synchronize()
    assert rel_l2(host(nu(image, omega, smaps=smaps)), host(nu(image, omega.clone(), smaps=smaps))) <= 1e-6


def test_library_owned_graph_replay_matches_eager():
    """``set_graph_mode(True)``: from the fourth call with the same argument buffers the forward / adjoint NUFFT is a
    replayed CUDA graph; new CONTENTS in the same buffers give new results, an edited trajectory gets a new plan and a
    new graph, and autograd calls stay eager."""
    import torchkbnufft_b200 as tkbn
    from torchkbnufft_b200 import _lib
    from torchkbnufft_b200._nufft import graphs

    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(5)
    im_size = (64, 48)
    x = torch.randn((2, 1) + im_size, dtype=torch.complex64, generator=g).to(dev)
    s = torch.randn((1, 5) + im_size, dtype=torch.complex64, generator=g).to(dev)
    om = ((torch.rand((2, 900), generator=g) - 0.5) * 6.2).to(dev)
    nu = tkbn.KbNufft(im_size=im_size, dtype=torch.complex64).to(dev)
    na = tkbn.KbNufftAdjoint(im_size=im_size, dtype=torch.complex64).to(dev)
    want_k = nu(x, om, smaps=s).clone()
    want_im = na(want_k, om, smaps=s).clone()
    lib = _lib.load()
    tkbn.set_graph_mode(True)
    try:
        for it in range(8):
            k = nu(x, om, smaps=s)
            im = na(k, om, smaps=s)
            torch.cuda.synchronize()
            assert torch.equal(k, want_k) and torch.equal(im, want_im), it
        assert sum(e.graph is not None for e in graphs._CACHE.values()) == 2
        before = lib.b2n_launch_count()
        k = nu(x, om, smaps=s)
        assert lib.b2n_launch_count() == before  # replayed: the host path launched nothing itself
        # new contents, same buffers
        x.mul_(2.0)
        k2 = nu(x, om, smaps=s)
        torch.cuda.synchronize()
        assert torch.allclose(k2, 2.0 * want_k, rtol=1e-5, atol=1e-5)
        # edited trajectory: new version -> new plan, eager again, right answer
        om.mul_(0.5)
        k3 = nu(x, om, smaps=s)
        tkbn.set_graph_mode(False)
        ref3 = nu(x, om, smaps=s)
        assert torch.equal(k3, ref3)
        tkbn.set_graph_mode(True)
        # autograd calls are never replayed
        xg = x.clone().requires_grad_(True)
        nu(xg, om, smaps=s).abs().sum().backward()
        assert xg.grad is not None and torch.isfinite(torch.view_as_real(xg.grad)).all()
    finally:
        tkbn.set_graph_mode(False)
    assert not graphs._CACHE


@pytest.mark.parametrize("form", [1, 2])
@pytest.mark.parametrize("world", [1, 2, 5])
def test_peer_allreduce_protocol_on_one_device(world, form):
    """b2n_peer_allreduce_sum with every "rank" on ONE device (windows of the same process, one stream per rank, no
    IPC): pushes, flags, rank-ordered sums, double-buffered slots over repeated calls, the scalar path for odd and
    unaligned lengths, in place and out of place.  The multi-process version over NVLink is tests/test_gpu_multi.py."""
    import ctypes

    lib = _lib.load()
    # form 1: one-shot exchange; form 2: two-shot (reduce-scatter + all-gather) wherever the operands are float4-aligned
    lib.b2n_set_option(_lib.OPT_PEER_FORM, form)
    max_floats = 50000
    nbytes = ctypes.c_size_t(0)
    assert lib.b2n_peer_window_bytes(world, max_floats, ctypes.byref(nbytes)) == 0
    assert lib.b2n_peer_window_bytes(0, max_floats, ctypes.byref(nbytes)) == -1
    assert lib.b2n_peer_window_bytes(_lib.PEER_MAX_RANKS + 1, max_floats, ctypes.byref(nbytes)) == -1
    windows, comms = [], []
    for r in range(world):
        w, h = ctypes.c_void_p(None), ctypes.create_string_buffer(_lib.PEER_HANDLE_BYTES)
        _lib.check(lib.b2n_peer_window_create(nbytes.value, ctypes.byref(w), h), "b2n_peer_window_create")
        windows.append(w)
    for r in range(world):
        c = _lib.PeerComm()
        c.rank, c.world, c.max_floats = r, world, max_floats
        for q in range(world):
            c.window[q] = windows[q].value
        comms.append(c)
    streams = [torch.cuda.Stream() for _ in range(world)]
    gen = torch.Generator(device="cpu").manual_seed(17)
    try:
        assert lib.b2n_peer_allreduce_sum(ctypes.byref(comms[0]), 16, 16, max_floats + 1, 0) == -2  # B2N_E_RANGE
        for n, in_place in ((4096, True), (1, False), (8192 + 4, True), (12345, False), (49999, True), (50000, False),
                            (4, True), (20480, True), (20480, False)):
            parts = [torch.randn(n + 1, generator=gen).to(DEV) for _ in range(world)]
            want = parts[0].clone()
            for p in parts[1:]:
                want += p
            for off in (0, 1):  # off = 1: 4-byte aligned only -> scalar kernel
                ins = [p[off:off + n] if off else p[:n] for p in parts]
                outs = [x if in_place else torch.empty_like(x) for x in (i.clone() for i in ins)]
                srcs = outs if in_place else [i.clone() for i in ins]
                torch.cuda.synchronize()
                for r in range(world):
                    with torch.cuda.stream(streams[r]):
                        _lib.check(lib.b2n_peer_allreduce_sum(ctypes.byref(comms[r]), srcs[r].data_ptr(),
                                                              outs[r].data_ptr(), n, streams[r].cuda_stream),
                                   "b2n_peer_allreduce_sum")
                torch.cuda.synchronize()
                for r in range(world):
                    assert torch.equal(outs[r], want[off:off + n]), (n, off, r)
    finally:
        lib.b2n_set_option(_lib.OPT_PEER_FORM, 0)
        torch.cuda.synchronize()
        for w in windows:
            lib.b2n_peer_window_destroy(w)


@pytest.mark.parametrize("N, K, C, world, fused", [((32, 320), (64, 640), 4, 2, True), ((24, 320), (64, 640), 16, 3, True),
                                                   ((32, 320), (64, 640), 2, 2, True), ((40, 48), (96, 96), 3, 2, True),
                                                   ((40, 30), (96, 60), 3, 2, False)])
def test_adjoint_fft_with_the_allreduce_inside_its_last_pass(N, K, C, world, fused):
    """b2n_fft_adjoint_fused_allreduce with every "rank" on ONE device (one stream per rank): the last inverse pass
    pushes its finished rows into the other ranks' windows and adds what arrives (k_fft_rows_sense, one and several
    coil groups per row, down to two coils in a mostly idle CTA), or the stand-alone kernel runs behind the unfused
    route (row lengths without a compile-time plan).  All ranks
    must hold the rank-ordered sum of the separately computed partial images, over repeated calls."""
    import ctypes

    torch.manual_seed(11)
    lib = _lib.load()
    dt = torch.complex64
    n_img = 2 * N[0] * N[1]
    nbytes = ctypes.c_size_t(0)
    assert lib.b2n_peer_window_bytes(world, 2 * n_img, ctypes.byref(nbytes)) == 0
    windows, comms = [], []
    for r in range(world):
        w, h = ctypes.c_void_p(None), ctypes.create_string_buffer(_lib.PEER_HANDLE_BYTES)
        _lib.check(lib.b2n_peer_window_create(nbytes.value, ctypes.byref(w), h), "b2n_peer_window_create")
        windows.append(w)
    for r in range(world):
        c = _lib.PeerComm()
        c.rank, c.world, c.max_floats = r, world, 2 * n_img
        for q in range(world):
            c.window[q] = windows[q].value
        comms.append(c)
    streams = [torch.cuda.Stream() for _ in range(world)]
    scal = torch.randn(N, dtype=dt, device=DEV)
    try:
        for rep in range(4):
            grids = [torch.randn((1, C) + K, dtype=dt, device=DEV) for _ in range(world)]
            smaps = [torch.randn((1, C) + N, dtype=dt, device=DEV) for _ in range(world)]
            parts = [eng_fft.fused_fft_adjoint(g, N, s, scal, 0.5) for g, s in zip(grids, smaps)]
            want = parts[0].clone()
            for p in parts[1:]:
                want += p
            torch.cuda.synchronize()
            before = lib.b2n_launch_count()
            outs = []
            for r in range(world):
                with torch.cuda.stream(streams[r]):
                    outs.append(eng_fft.fused_fft_adjoint(grids[r], N, smaps[r], scal, 0.5, peer_comm=comms[r]))
            torch.cuda.synchronize()
            launches = (lib.b2n_launch_count() - before) // world
            for r in range(world):
                assert torch.equal(outs[r], want), (rep, r, float((outs[r] - want).abs().max()))
        # compile-time planned row lengths carry the exchange in the row pass itself (from two coils on): two launches
        # per call; a run-time planned row length (60) takes the unfused passes and the stand-alone kernel
        assert (launches == 2) if fused else (launches >= 3)
    finally:
        torch.cuda.synchronize()
        for w in windows:
            lib.b2n_peer_window_destroy(w)
