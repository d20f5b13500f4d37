"""Sparse-matrix interpolation mode: out of the engine's scope.

The reference offers precomputed sparse interpolation matrices as an alternative to
table interpolation (``torchkbnufft/_nufft/spmat.py:10-105``) and itself labels that
mode slow / not recommended (``README.md:33-37``).  Only the table path is
accelerated here; the symbol is kept so imports do not break."""
from __future__ import annotations


def calc_tensor_spmatrix(*args, **kwargs):
    raise NotImplementedError(
        "calc_tensor_spmatrix (sparse-matrix interpolation) is not provided by the B200 engine; "
        "use the default table interpolation (interp_mats=None)."
    )
