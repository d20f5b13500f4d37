"""Autograd drop-in behaviour on the GPU (the reference's property tests,
tests/conftest.py:40-87, tests/test_interp.py:45-109, tests/test_sense_nufft.py:28-157):
adjointness and `grad 0.5*||A x||^2 == A^H A x`, float64, torch.allclose defaults."""
import numpy as np
import pytest
import torch

import torchkbnufft_b200 as tkbn

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def noisy(shape, is_complex):
    x = torch.arange(int(np.prod(shape)), dtype=torch.float64).reshape(shape)
    if is_complex:
        return (x + torch.randn(shape, dtype=torch.float64) + 1j * torch.randn(shape, dtype=torch.float64)).to(DEV)
    return (x + torch.randn(shape, dtype=torch.float64)).to(DEV)


def ktraj(ndims, klength, batch=None):
    shape = (ndims, klength) if batch is None else (batch, ndims, klength)
    return (torch.rand(shape, dtype=torch.float64) * 2 * np.pi - np.pi).to(DEV)


CASES = [
    ([1, 1, 19], [1, 1, 25], True),
    ([3, 1, 13, 2], [3, 1, 18, 2], False),
    ([1, 1, 32, 16], [1, 1, 83], True),
    ([5, 1, 15, 12, 2], [5, 1, 83, 2], False),
    ([3, 2, 13, 18, 12], [3, 2, 112], True),
    ([1, 2, 17, 19, 12, 2], [1, 2, 112, 2], False),
]


def adjoint_and_autograd(fw, ad, x, y, om, **extra):
    assert torch.allclose(tkbn.inner_product(fw(x, om, **extra), y), tkbn.inner_product(x, ad(y, om, **extra)))
    x = x.clone().requires_grad_(True)
    y = y.clone().requires_grad_(True)
    fx, ay = fw(x, om, **extra), ad(y, om, **extra)
    mag2 = lambda t: (torch.abs(t) ** 2 if t.is_complex() else t ** 2)
    (mag2(fx) / 2).sum().backward()
    (mag2(ay) / 2).sum().backward()
    assert torch.allclose(x.grad, ad(fx.detach(), om, **extra))
    assert torch.allclose(y.grad, fw(ay.detach(), om, **extra))


@pytest.mark.parametrize("mode", ["atomic", "sorted"])
@pytest.mark.parametrize("shape, kdata_shape, is_complex", CASES)
def test_interp_adjoint_and_autograd(shape, kdata_shape, is_complex, mode):
    torch.manual_seed(123)
    im_size = shape[2:] if is_complex else shape[2:-1]
    kw = dict(im_size=im_size, grid_size=im_size, dtype=torch.complex128)
    fw, ad = tkbn.KbInterp(**kw).to(DEV), tkbn.KbInterpAdjoint(**kw).to(DEV)
    tkbn.set_adjoint_mode(mode)
    try:
        adjoint_and_autograd(fw, ad, noisy(shape, is_complex), noisy(kdata_shape, is_complex),
                             ktraj(len(im_size), kdata_shape[2]))
    finally:
        tkbn.set_adjoint_mode("atomic")


@pytest.mark.parametrize("batched", [False, True])
@pytest.mark.parametrize("shape, kdata_shape, is_complex", CASES)
def test_sense_nufft_adjoint_and_autograd(shape, kdata_shape, is_complex, batched):
    torch.manual_seed(123)
    im_size = shape[2:] if is_complex else shape[2:-1]
    ncoil = 4
    im_shape = list(shape)
    im_shape[1] = 1
    smap_shape = list(shape)
    smap_shape[1] = ncoil
    kd = list(kdata_shape)
    kd[1] = ncoil
    kw = dict(im_size=im_size, dtype=torch.complex128)
    fw, ad = tkbn.KbNufft(**kw).to(DEV), tkbn.KbNufftAdjoint(**kw).to(DEV)
    om = ktraj(len(im_size), kdata_shape[2], batch=shape[0] if batched else None)
    for norm in (None, "ortho"):
        adjoint_and_autograd(fw, ad, noisy(im_shape, is_complex), noisy(kd, is_complex), om,
                             smaps=noisy(smap_shape, is_complex), norm=norm)


def test_batched_trajectory_equals_loop():
    torch.manual_seed(0)
    kw = dict(im_size=(15, 12), dtype=torch.complex128)
    fw, ad = tkbn.KbNufft(**kw).to(DEV), tkbn.KbNufftAdjoint(**kw).to(DEV)
    x, y = noisy([5, 2, 15, 12], True), noisy([5, 2, 83], True)
    om = ktraj(2, 83, batch=5)
    loop_f = torch.cat([fw(x[i:i + 1], om[i]) for i in range(5)])
    loop_a = torch.cat([ad(y[i:i + 1], om[i]) for i in range(5)])
    assert torch.allclose(fw(x, om), loop_f) and torch.allclose(ad(y, om), loop_a)


@pytest.mark.parametrize("shape, grid_size, kdata_shape, norm", [
    ([1, 3, 19], [57], [1, 3, 25], "ortho"),
    ([1, 4, 32, 16], [64, 24], [1, 4, 83], None),
    ([3, 10, 13, 18, 12], [20, 26, 37], [3, 10, 112], None),
])
def test_toeplitz_vs_normal_operator(shape, grid_size, kdata_shape, norm):
    """tests/test_toep.py:9-60 (arbitrary grids, tolerance 1e-2)."""
    torch.manual_seed(123)
    im_size = shape[2:]
    im_shape = list(shape)
    im_shape[1] = 1
    image, smaps = noisy(im_shape, True), noisy(shape, True)
    om = ktraj(len(im_size), kdata_shape[2])
    fw = tkbn.KbNufft(im_size=im_size, grid_size=grid_size, dtype=torch.complex128).to(DEV)
    ad = tkbn.KbNufftAdjoint(im_size=im_size, grid_size=grid_size, dtype=torch.complex128).to(DEV)
    kern = tkbn.calc_toeplitz_kernel(om, im_size, grid_size=grid_size, norm=norm)
    fbn = ad(fw(image, om, smaps=smaps, norm=norm), om, smaps=smaps, norm=norm)
    fbt = tkbn.ToepNufft()(image, kern, smaps=smaps, norm=norm)
    assert torch.norm(fbn - fbt) / torch.norm(fbn) < 1e-2


def test_batched_weighted_toeplitz():
    """tests/test_toep.py:118-175 (batched trajectories, weights, ortho, tolerance 1e-4)."""
    torch.manual_seed(123)
    shape, klen = [2, 4, 32, 16], 83
    image, smaps = noisy([2, 1, 32, 16], True), noisy(shape, True)
    om = ktraj(2, klen, batch=2)
    weights = torch.rand(2, 1, klen, dtype=torch.float64, device=DEV)
    fw = tkbn.KbNufft(im_size=shape[2:], dtype=torch.complex128).to(DEV)
    ad = tkbn.KbNufftAdjoint(im_size=shape[2:], dtype=torch.complex128).to(DEV)
    kern = tkbn.calc_toeplitz_kernel(om, shape[2:], weights=weights, norm="ortho")
    fbn = ad(weights * fw(image, om, smaps=smaps, norm="ortho"), om, smaps=smaps, norm="ortho")
    fbt = tkbn.ToepNufft()(image, kern, smaps=smaps, norm="ortho")
    assert torch.norm(fbn - fbt) / torch.norm(fbn) < 1e-4
