"""Generate tests/golden/ref_spmat.npz from the REFERENCE (test infrastructure; build container only:
``python oracle/make_golden_spmat.py``).  Pins the sparse-matrix interpolation shim
(torchkbnufft_b200/_nufft/spmat.py) against the reference's ``calc_tensor_spmatrix``
(``_nufft/spmat.py:10-105``) and its sparse forward / adjoint NUFFT (``modules/kbnufft.py``).
Inputs are regenerated from seeds by the tests (``spmat_case_inputs``)."""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(ROOT, "tests"))
warnings.filterwarnings("ignore")
import torchkbnufft as tkbn  # noqa: E402  (the reference)
from golden_cases import SPMAT_CASES, spmat_case_inputs  # noqa: E402


def coalesced(mat):
    m = mat.coalesce()
    return m.indices().numpy(), m.values().numpy()


out = {}
for name, cfg in SPMAT_CASES.items():
    omega, image, kdata, smaps = spmat_case_inputs(name)
    kw = {k: cfg[k] for k in ("grid_size", "numpoints", "n_shift") if cfg.get(k) is not None}
    real, imag = tkbn.calc_tensor_spmatrix(torch.from_numpy(omega), cfg["im_size"], **kw)
    idx, rv = coalesced(real)
    idx2, iv = coalesced(imag)
    assert np.array_equal(idx, idx2)
    out[f"{name}/index"], out[f"{name}/real"], out[f"{name}/imag"] = idx, rv, iv
    cdt = torch.complex64 if omega.dtype == np.float32 else torch.complex128
    nu = tkbn.KbNufft(im_size=cfg["im_size"], dtype=cdt, **kw)
    na = tkbn.KbNufftAdjoint(im_size=cfg["im_size"], dtype=cdt, **kw)
    om = torch.from_numpy(omega)
    x, y, s = torch.from_numpy(image), torch.from_numpy(kdata), torch.from_numpy(smaps)
    out[f"{name}/nufft"] = nu(x, om, (real, imag), smaps=s).numpy()
    out[f"{name}/nufft_adj"] = na(y, om, (real, imag), smaps=s).numpy()
    out[f"{name}/nufft_ortho"] = nu(x, om, (real, imag), smaps=s, norm="ortho").numpy()
    ki = tkbn.KbInterp(im_size=cfg["im_size"], dtype=cdt, **kw)
    kia = tkbn.KbInterpAdjoint(im_size=cfg["im_size"], dtype=cdt, **kw)
    grid = torch.from_numpy(np.ascontiguousarray(np.broadcast_to(
        np.arange(int(np.prod(ki.grid_size.tolist())), dtype=np.float64).reshape(ki.grid_size.tolist()) % 7 - 3,
        (1, 2) + tuple(ki.grid_size.tolist())))).to(cdt)
    out[f"{name}/interp"] = ki(grid, om, (real, imag)).numpy()
    out[f"{name}/interp_adj"] = kia(y[:, :2].contiguous(), om, (real, imag)).numpy()
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_spmat.npz"), **out)
print("wrote", len(out), "arrays")
