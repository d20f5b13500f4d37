"""Seeded synthetic workloads for the parity tests and ``bench.py``.

The reference ships no trajectory generators in its package (only notebooks and
``profile_torchkbnufft.py:205-216``); these restate the standard MRI sampling
patterns SURVEY.md section 8(d) fixes for the five BASELINE.json configs.  numpy only,
float64 generation then cast, so the same arrays can be rebuilt on any host.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

GOLDEN_RATIO = (1.0 + np.sqrt(5.0)) / 2.0


def radial_2d(n_spokes: int, n_read: int, golden: bool = False) -> np.ndarray:
    """2-D radial trajectory ``(2, n_spokes*n_read)`` in rad/voxel.

    Spoke ``i`` has angle ``i*pi/n_spokes`` (uniform) or ``i*pi/phi`` (golden
    angle); the readout is ``linspace(-pi, pi, n_read, endpoint=False)``;
    ``omega = (r sin(theta), r cos(theta))``.
    """
    i = np.arange(n_spokes, dtype=np.float64)
    theta = i * np.pi / GOLDEN_RATIO if golden else i * np.pi / n_spokes
    r = np.linspace(-np.pi, np.pi, n_read, endpoint=False)
    ky = np.outer(np.sin(theta), r).reshape(-1)
    kx = np.outer(np.cos(theta), r).reshape(-1)
    return np.stack([ky, kx])


def kooshball_3d(n_spokes: int, n_read: int) -> np.ndarray:
    """3-D radial ("kooshball") trajectory ``(3, n_spokes*n_read)``: spoke
    directions on a Fibonacci sphere, full-diameter readouts."""
    i = np.arange(n_spokes, dtype=np.float64) + 0.5
    z = 1.0 - 2.0 * i / n_spokes
    phi = np.pi * (1.0 + np.sqrt(5.0)) * i
    s = np.sqrt(np.maximum(0.0, 1.0 - z * z))
    dirs = np.stack([z, s * np.sin(phi), s * np.cos(phi)])  # (3, n_spokes)
    r = np.linspace(-np.pi, np.pi, n_read, endpoint=False)
    return (dirs[:, :, None] * r[None, None, :]).reshape(3, -1)


def complex_normal(rng: np.random.Generator, shape, cdtype=np.complex64) -> np.ndarray:
    """re, im ~ N(0, 1)."""
    rdtype = np.float32 if np.dtype(cdtype) == np.complex64 else np.float64
    out = rng.standard_normal(shape, dtype=np.float64) + 1j * rng.standard_normal(shape, dtype=np.float64)
    return out.astype(cdtype) if rdtype == np.float32 else out


def smooth_smaps(rng: np.random.Generator, n_coils: int, im_size, cdtype=np.complex64) -> np.ndarray:
    """Smooth Gaussian-profile coil maps ``(1, C, *im_size)`` with random centres
    and a linear phase, normalised to unit sum of squares."""
    axes = [np.linspace(-1.0, 1.0, n) for n in im_size]
    mesh = np.meshgrid(*axes, indexing="ij")
    maps = []
    for _ in range(n_coils):
        centre = rng.uniform(-1.0, 1.0, size=len(im_size))
        slope = rng.uniform(-1.5, 1.5, size=len(im_size))
        dist2 = sum((m - c) ** 2 for m, c in zip(mesh, centre))
        phase = sum(m * s for m, s in zip(mesh, slope))
        maps.append(np.exp(-dist2 / 1.2) * np.exp(1j * phase))
    maps = np.stack(maps)
    maps /= np.sqrt(np.sum(np.abs(maps) ** 2, axis=0, keepdims=True))
    return maps[None].astype(cdtype)


@dataclass(frozen=True)
class Workload:
    """One BASELINE.json config."""

    name: str
    im_size: Tuple[int, ...]
    n_coils: int
    n_batch: int
    n_spokes: int
    n_read: int
    kind: str  # "radial", "golden", "koosh"
    description: str

    @property
    def grid_size(self) -> Tuple[int, ...]:
        return tuple(2 * n for n in self.im_size)

    @property
    def n_points(self) -> int:
        return self.n_spokes * self.n_read

    def trajectory(self, real_dtype=np.float32) -> np.ndarray:
        if self.kind == "koosh":
            om = kooshball_3d(self.n_spokes, self.n_read)
        else:
            om = radial_2d(self.n_spokes, self.n_read, golden=(self.kind == "golden"))
        return np.ascontiguousarray(om.astype(real_dtype))

    def scaled(self, spoke_fraction: float) -> "Workload":
        """Same geometry with fewer spokes (bounded CPU-baseline samples)."""
        n = max(1, int(round(self.n_spokes * spoke_fraction)))
        return Workload(self.name, self.im_size, self.n_coils, self.n_batch, n, self.n_read, self.kind,
                        self.description)


WORKLOADS = {
    "cfg1": Workload("cfg1", (256, 256), 1, 1, 402, 512, "radial",
                     "2D single-coil 256x256, radial 402 spokes x 512 readout"),
    "cfg2": Workload("cfg2", (320, 320), 16, 1, 200, 640, "golden",
                     "2D SENSE 320x320, 16 coils, golden-angle radial 200x640"),
    "cfg3": Workload("cfg3", (384, 384), 32, 1, 240, 768, "golden",
                     "2D ToepNufft CG-SENSE 384x384, 32 coils, golden-angle radial 240x768"),
    "cfg4": Workload("cfg4", (128, 128, 128), 8, 1, 32768, 256, "koosh",
                     "3D kooshball 128^3, 8 coils, 32768 spokes x 256 readout"),
    "cfg5": Workload("cfg5", (256, 256), 16, 64, 402, 512, "radial",
                     "batched 2D SENSE 256x256, 64 slices x 16 coils, radial 402x512"),
}


def make_inputs(wl: Workload, seed: int = 0, n_batch: Optional[int] = None, cdtype=np.complex64):
    """Seeded ``(image, smaps, kdata, omega)`` for a workload (numpy arrays).

    image ``(B, 1, *N)``, smaps ``(1, C, *N)`` (complex normal, as the parity
    runs use), kdata ``(B, C, M)``, omega ``(d, M)``.
    """
    rng = np.random.default_rng(seed)
    B = wl.n_batch if n_batch is None else n_batch
    rdtype = np.float32 if np.dtype(cdtype) == np.complex64 else np.float64
    omega = wl.trajectory(rdtype)
    image = complex_normal(rng, (B, 1) + wl.im_size, cdtype)
    smaps = complex_normal(rng, (1, wl.n_coils) + wl.im_size, cdtype)
    kdata = complex_normal(rng, (B, wl.n_coils, wl.n_points), cdtype)
    return image, smaps, kdata, omega
