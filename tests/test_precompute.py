"""Host precompute (tables, apodisation, offsets, dtype pairing) against the
reference's module buffers dumped in tests/golden/ref_buffers.npz -- bit-exact."""
import numpy as np
import pytest
import torch

import torchkbnufft_b200 as tkbn
from golden_cases import BUFFER_GEOMS

BUFFER_NAMES = ["im_size", "grid_size", "n_shift", "numpoints", "offsets", "table_oversamp", "order", "alpha",
                "scaling_coef"]


@pytest.mark.parametrize("gi", range(len(BUFFER_GEOMS)))
def test_module_buffers_equal_reference(gi, ref_buffers):
    g = dict(BUFFER_GEOMS[gi])
    dtype = torch.complex64 if g.pop("c64") else torch.complex128
    ob = tkbn.KbNufft(dtype=dtype, **g)
    names = [n for n, _ in ob.named_buffers()]
    ndim = len(g["im_size"])
    assert names == [f"table_{i}" for i in range(ndim)] + BUFFER_NAMES  # state_dict layout of the reference
    for name, buf in ob.named_buffers():
        ref = ref_buffers[f"g{gi}_{name}"]
        got = buf.numpy()
        assert got.dtype == ref.dtype and got.shape == ref.shape, name
        assert np.array_equal(got, ref), name


def test_interp_module_has_no_scaling_coef():
    ob = tkbn.KbInterp(im_size=(8, 8))
    assert "scaling_coef" not in dict(ob.named_buffers())
    assert ob.table_0.dtype == torch.complex64 and ob.n_shift.dtype == torch.float32
    assert ob.offsets.shape == (36, 2) and ob.offsets.dtype == torch.long
    assert ob.offsets[1].tolist() == [0, 1]  # row-major neighbour order


def test_to_pairs_real_and_complex_dtypes():
    ob = tkbn.KbNufft(im_size=(8, 6))
    ob = ob.to(torch.float64)
    assert ob.table_0.dtype == torch.complex128 and ob.scaling_coef.dtype == torch.complex128
    assert ob.n_shift.dtype == torch.float64 and ob.numpoints.dtype == torch.long
    ob = ob.to(torch.complex64)
    assert ob.table_1.dtype == torch.complex64 and ob.alpha.dtype == torch.float32
    with pytest.raises(TypeError):
        ob.to(torch.int32)


def test_default_dtype_follows_torch_default():
    prev = torch.get_default_dtype()
    try:
        torch.set_default_dtype(torch.float64)
        assert tkbn.KbInterp(im_size=(5,)).table_0.dtype == torch.complex128
    finally:
        torch.set_default_dtype(prev)


def test_bad_dimension_lists_assert():
    with pytest.raises(AssertionError):
        tkbn.KbNufft(im_size=(8, 8), grid_size=(16,))
    with pytest.raises(AssertionError):
        tkbn.KbNufft(im_size=(8, 8), numpoints=(6, 6, 6))


def test_repr_lists_buffers():
    text = repr(tkbn.KbInterp(im_size=(4, 4)))
    assert "KbInterp" in text and "tensor: table_0, shape: (6145,)" in text
