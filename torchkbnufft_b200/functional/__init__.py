from .interp import kb_spmat_interp, kb_spmat_interp_adjoint, kb_table_interp, kb_table_interp_adjoint
from .nufft import fft_filter, kb_spmat_nufft, kb_spmat_nufft_adjoint, kb_table_nufft, kb_table_nufft_adjoint

__all__ = [
    "fft_filter",
    "kb_spmat_interp",
    "kb_spmat_interp_adjoint",
    "kb_spmat_nufft",
    "kb_spmat_nufft_adjoint",
    "kb_table_interp",
    "kb_table_interp_adjoint",
    "kb_table_nufft",
    "kb_table_nufft_adjoint",
]
