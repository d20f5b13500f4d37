"""Generate tests/golden/*.npz from the REFERENCE (test infrastructure).

Run in the build container only (``python oracle/make_golden.py``): it imports
the unmodified reference from /root/reference, which does not exist on the GPU
box, and writes small fixtures that travel with the repo:

  ref_interp_golden.npz / ref_nufft_golden.npz
      the reference's own golden vectors (tests/data/interp_data.pkl and
      nufft_data.pkl, consumed by tests/test_interp.py:16-31 and
      tests/test_nufft.py:15-30), re-saved as complex128 arrays.
  ref_buffers.npz
      module buffers (tables, scaling_coef, offsets, n_shift, alpha) of the
      reference for several geometries -> pins the host precompute.
  ref_cases.npz
      live reference outputs (CPU) for small seeded cases in complex64 and
      complex128: forward/adjoint interpolation, per-offset integer indices of
      calc_coef_and_indices, SENSE NUFFT forward/adjoint, Toeplitz kernel and
      apply, density compensation, batched trajectories.
  ref_cfg1.npz
      BASELINE config 1 at FULL size in complex64 (subsampled outputs).

Inputs are regenerated from seeds by tests (see `case_inputs`), only the
reference's outputs are stored.
"""
from __future__ import annotations

import importlib.util
import os
import pickle
import sys
import warnings

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden")

sys.path.insert(0, REF)
warnings.filterwarnings("ignore")
import torchkbnufft as tkbn  # noqa: E402  (the reference)
from torchkbnufft._nufft.interp import calc_coef_and_indices  # noqa: E402

spec = importlib.util.spec_from_file_location("workloads", os.path.join(ROOT, "torchkbnufft_b200", "workloads.py"))
workloads = importlib.util.module_from_spec(spec)
sys.modules["workloads"] = workloads
spec.loader.exec_module(workloads)

sys.path.insert(0, os.path.join(ROOT, "tests"))
from golden_cases import BUFFER_GEOMS, CASES, case_inputs  # noqa: E402


def c128(t):
    t = t.detach().cpu()
    if not t.is_complex():
        t = torch.view_as_complex(t.contiguous())
    return t.numpy()


def resave_pickles():
    for name in ("interp", "nufft"):
        with open(os.path.join(REF, "tests", "data", f"{name}_data.pkl"), "rb") as f:
            data = pickle.load(f)
        out = {}
        for i, (image, ktraj, kdata) in enumerate(data):
            out[f"image_{i}"] = c128(image)
            out[f"ktraj_{i}"] = ktraj.numpy()
            out[f"kdata_{i}"] = c128(kdata)
        out["n_cases"] = np.array(len(data))
        np.savez_compressed(os.path.join(OUT, f"ref_{name}_golden.npz"), **out)
        print(name, "golden:", len(data), "cases")


def dump_buffers():
    out = {}
    for gi, g in enumerate(BUFFER_GEOMS):
        dtype = torch.complex64 if g["c64"] else torch.complex128
        ob = tkbn.KbNufft(im_size=g["im_size"], grid_size=g.get("grid_size"), numpoints=g.get("numpoints", 6),
                          n_shift=g.get("n_shift"), table_oversamp=g.get("table_oversamp", 2 ** 10),
                          kbwidth=g.get("kbwidth", 2.34), order=g.get("order", 0.0), dtype=dtype)
        for name, buf in ob.named_buffers():
            out[f"g{gi}_{name}"] = buf.numpy()
    np.savez_compressed(os.path.join(OUT, "ref_buffers.npz"), **out)
    print("buffers:", len(BUFFER_GEOMS), "geometries")


def ref_indices(omega, ob):
    """Per-offset arr_ind via the reference's own calc_coef_and_indices, with the
    caller prelude of table_interp_one_batch (_nufft/interp.py:171-177)."""
    grid_size = ob.grid_size
    numpoints, L = ob.numpoints, ob.table_oversamp
    tables = [getattr(ob, f"table_{i}") for i in range(len(grid_size))]
    tm = omega / (2 * np.pi / grid_size.to(omega).unsqueeze(-1))
    centers = torch.floor(numpoints * L / 2).to(dtype=torch.long)
    base_offset = 1 + torch.floor(tm - numpoints.unsqueeze(-1) / 2.0).to(dtype=torch.long)
    inds = []
    for offset in ob.offsets.to(torch.long):
        _, arr_ind = calc_coef_and_indices(tm, base_offset, offset, tables, centers, L, grid_size)
        inds.append(arr_ind)
    return torch.stack(inds).numpy()


def run_cases():
    out = {}
    for name, case in CASES.items():
        for prec in ("c64", "c128"):
            cd = np.complex64 if prec == "c64" else np.complex128
            td = torch.complex64 if prec == "c64" else torch.complex128
            rdt = torch.float32 if prec == "c64" else torch.float64
            torch.set_default_dtype(rdt)  # the reference derives J/2 etc. in the default dtype
            inp = case_inputs(case, cd)
            kw = dict(im_size=case["im_size"], grid_size=case.get("grid_size"), numpoints=case.get("numpoints", 6),
                      n_shift=case.get("n_shift"), table_oversamp=case.get("table_oversamp", 2 ** 10), dtype=td)
            omega = torch.from_numpy(inp["omega"])
            grid = torch.from_numpy(inp["grid"])
            kdata = torch.from_numpy(inp["kdata"])
            image = torch.from_numpy(inp["image"])
            smaps = torch.from_numpy(inp["smaps"])
            interp, interp_adj = tkbn.KbInterp(**kw), tkbn.KbInterpAdjoint(**kw)
            nufft, nufft_adj = tkbn.KbNufft(**kw), tkbn.KbNufftAdjoint(**kw)
            key = f"{name}_{prec}_"
            out[key + "interp"] = interp(grid, omega).numpy()
            out[key + "interp_adj"] = interp_adj(kdata, omega).numpy()
            if omega.ndim == 2:
                out[key + "arr_ind"] = ref_indices(omega, interp)
            for norm in (None, "ortho"):
                tag = "ortho" if norm else "none"
                out[key + f"sense_fwd_{tag}"] = nufft(image, omega, smaps=smaps, norm=norm).numpy()
                out[key + f"sense_adj_{tag}"] = nufft_adj(kdata, omega, smaps=smaps, norm=norm).numpy()
            out[key + "nufft_fwd_nosmap"] = nufft(grid_to_image(inp), omega).numpy()
            if case.get("toep", True):
                gs = case.get("grid_size")
                for norm in (None, "ortho"):
                    tag = "ortho" if norm else "none"
                    kern = tkbn.calc_toeplitz_kernel(omega, case["im_size"], norm=norm, grid_size=gs,
                                                     numpoints=case.get("numpoints", 6),
                                                     table_oversamp=case.get("table_oversamp", 2 ** 10))
                    out[key + f"toep_kernel_{tag}"] = kern.numpy()
                    out[key + f"toep_apply_{tag}"] = tkbn.ToepNufft()(image, kern, smaps=smaps, norm=norm).numpy()
                w = torch.from_numpy(inp["weights"])
                kern_w = tkbn.calc_toeplitz_kernel(omega, case["im_size"], weights=w, norm="ortho", grid_size=gs,
                                                   numpoints=case.get("numpoints", 6),
                                                   table_oversamp=case.get("table_oversamp", 2 ** 10))
                out[key + "toep_kernel_weighted"] = kern_w.numpy()
            if case.get("dcomp", True):
                out[key + "dcomp"] = tkbn.calc_density_compensation_function(
                    omega, case["im_size"], num_iterations=3, grid_size=case.get("grid_size"),
                    numpoints=case.get("numpoints", 6), n_shift=case.get("n_shift"),
                    table_oversamp=case.get("table_oversamp", 2 ** 10)).numpy()
        print("case", name, "done")
    torch.set_default_dtype(torch.float32)
    np.savez_compressed(os.path.join(OUT, "ref_cases.npz"), **out)


def grid_to_image(inp):
    return torch.from_numpy(inp["image_multi"])


def run_cfg1():
    """BASELINE config 1 at full size (complex64) through the reference."""
    torch.set_default_dtype(torch.float32)
    wl = workloads.WORKLOADS["cfg1"]
    image, smaps, kdata, omega = workloads.make_inputs(wl, seed=0)
    rng = np.random.default_rng(1)
    grid = workloads.complex_normal(rng, (1, 1) + wl.grid_size)
    interp = tkbn.KbInterp(im_size=wl.im_size, dtype=torch.complex64)
    interp_adj = tkbn.KbInterpAdjoint(im_size=wl.im_size, dtype=torch.complex64)
    nufft = tkbn.KbNufft(im_size=wl.im_size, dtype=torch.complex64)
    nufft_adj = tkbn.KbNufftAdjoint(im_size=wl.im_size, dtype=torch.complex64)
    om = torch.from_numpy(omega)
    out = {}
    step = 16
    out["step"] = np.array(step)
    out["interp_sub"] = interp(torch.from_numpy(grid), om).numpy()[..., ::step]
    adj = interp_adj(torch.from_numpy(kdata), om).numpy()
    out["interp_adj_norm"] = np.array(np.linalg.norm(adj.astype(np.complex128)))
    out["interp_adj_centre"] = adj[..., :48, :48]  # radial centre (k=0 wraps to the grid corner)
    out["interp_adj_rows"] = adj[..., 100:104, :]
    out["nufft_sub"] = nufft(torch.from_numpy(image), om).numpy()[..., ::step]
    out["nufft_adj"] = nufft_adj(torch.from_numpy(kdata), om).numpy()[..., ::4, ::4]
    out["arr_ind_sub"] = ref_indices(om, interp)[:, ::step]
    np.savez_compressed(os.path.join(OUT, "ref_cfg1.npz"), **out)
    print("cfg1 done")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    resave_pickles()
    dump_buffers()
    run_cases()
    run_cfg1()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")
