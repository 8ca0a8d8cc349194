"""Krylov seed generation of the reference (host side, not on the hot path).

``op_add_noise`` + ``mth_rand`` (core/utils.f90:297-359, 408-418): deterministic pseudo-noise from
the point coordinates, element id and local indices, then made C0 by ``dssum * vmult`` and masked.
The arithmetic is numpy on the host; the dssum / vmult / mask / normalisation run on the device
through the library, like every other vector operation.
"""
from __future__ import annotations

import numpy as np

from .api import Sem, nek_dvector, k_normalize


def hashed_field(glo: np.ndarray, component: int, seed: int = 0) -> np.ndarray:
    """Pseudo-random values in [-1, 1) that depend only on the GLOBAL node id (splitmix64 finaliser): every
    copy of a node gets the same value and the field does not depend on how the mesh is partitioned, so
    runs on 1, 2, 4 and 8 ranks start from the same Krylov seed (bench.py's parity block relies on it)."""
    with np.errstate(over='ignore'):
        z = glo.astype(np.uint64) * np.uint64(3) + np.uint64(component) + np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15)
        z = (z + np.uint64(0x9E3779B97F4A7C15))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (2.0 / (1 << 53)) - 1.0

# (fc1, fc2, fc3) per velocity component, core/utils.f90:321-329
NOISE_FC = ((3.0e4, -1.5e3, 0.5e5), (2.3e4, 2.3e3, -2.0e5), (2.0e4, 1.0e3, 1.0e5))


def mth_rand(ix, iy, iz, ieg, xl, fc, if3d: bool):
    """core/utils.f90:408-418; ix, iy, iz, ieg are 1-based, xl the point coordinates."""
    r = fc[0] * (ieg + xl[0] * np.sin(xl[1])) + fc[1] * ix * iy + fc[2] * ix
    if if3d:
        r = fc[0] * (ieg + xl[2] * np.sin(r)) + fc[1] * iz * ix + fc[2] * iz
    return np.cos(1.0e3 * np.sin(1.0e3 * np.sin(r)))


def noise_fields(coords, e0: int = 0):
    """Raw noise of op_add_noise before averaging; ``e0`` = global id offset of this rank's elements."""
    x = coords[0]
    if3d = len(coords) == 3
    nel, lx = x.shape[0], x.shape[-1]
    ieg = (e0 + np.arange(1, nel + 1)).reshape((nel,) + (1,) * (x.ndim - 1))
    i1 = np.arange(1, lx + 1)
    if if3d:
        ix, iy, iz = i1[None, None, None, :], i1[None, None, :, None], i1[None, :, None, None]
    else:
        ix, iy, iz = i1[None, None, :], i1[None, :, None], 1
    return [mth_rand(ix, iy, iz, ieg, coords, NOISE_FC[c], if3d) + np.zeros_like(x)
            for c in range(3 if if3d else 2)]


def seed_noise(sem: Sem, vec: nek_dvector, coords, e0: int = 0, extra_fields=0) -> float:
    """Fill ``vec`` with the reference's noise seed: noise -> dssum -> vmult -> mask -> unit norm
    (core/eigensolvers.f90:192-203 with ifseed_nois).  Returns the norm before normalisation."""
    fields = noise_fields(coords, e0)
    vec.upload(fields + [None] * extra_fields)
    for f in range(len(fields)):
        sem.dssum(vec, f)              # opdssum ; opcolv VMULT      (core/utils.f90:338-339)
        sem.col2(vec, f, 'vmult')
        sem.dssum(vec, f)              # dsavg = dssum ; col2 vmult  (:341-343)
        sem.col2(vec, f, 'vmult')
        sem.col2(vec, f, 'mask')       # bcdirVC                      (:346)
    return k_normalize(vec)
