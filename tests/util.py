"""Shared helpers for the parity tests."""
import ctypes as C

import numpy as np

import oracle_lib as ol
from oppositerenderer_b200 import capi

SEED = 1645301512  # RenderEngine/config.h:50 DEBUG_RANDOM_SEED


def synthetic_photons(n, seed, extent=(556.0, 548.8, 559.2), invalid_fraction=0.4, clustered=True):
    """Photon records in the reference's AoS layout (ppm/Photon.h) with the statistics of a real photon pass:
    a share of invalid (zero-power) slots, positions on the walls of a box, some dense clusters (caustics)."""
    rng = np.random.default_rng(seed)
    p = np.zeros(n, ol.PHOTON_DT)
    pos = rng.random((n, 3), dtype=np.float32) * np.array(extent, np.float32)
    wall = rng.integers(0, 6, n)
    for axis in range(3):
        pos[wall == 2 * axis, axis] = 0.0
        pos[wall == 2 * axis + 1, axis] = extent[axis]
    if clustered and n >= 64:
        k = n // 8
        c = np.array(extent, np.float32) * np.array([0.3, 0.0, 0.6], np.float32)
        pos[:k] = c + rng.normal(0, 2.0, (k, 3)).astype(np.float32) * np.array([1, 0, 1], np.float32)
    p["position"] = pos
    p["power"] = rng.random((n, 3), dtype=np.float32) * 3.0
    d = rng.normal(size=(n, 3)).astype(np.float32)
    p["rayDirection"] = d / np.linalg.norm(d, axis=1, keepdims=True)
    inval = rng.random(n) < invalid_fraction
    p["power"][inval] = 0.0
    p["position"][inval] = 0.0
    return p


def oracle_grid(photons):
    lib = ol.load("oracle")
    ph = photons.copy()
    n = ph.shape[0]
    keys = np.zeros(n, np.uint32)
    offsets = np.zeros(1000002, np.uint32)
    cell = C.c_float()
    gs = (C.c_uint32 * 3)()
    org = (C.c_float * 3)()
    nv = C.c_uint32()
    rc = lib.ora_build_photon_grid(ph.view(np.float32).ctypes.data_as(C.POINTER(C.c_float)), n, ol.uptr(keys), ol.uptr(offsets),
                                   C.byref(cell), gs, org, C.byref(nv))
    ncell = gs[0] * gs[1] * gs[2]
    return dict(rc=rc, photons=ph, keys=keys, offsets=offsets[:ncell + 1].copy() if rc == 0 else None, cellSize=np.float32(cell.value),
                gridSize=tuple(gs), origo=np.array(list(org), np.float32), numValid=nv.value)


def product_grid(photons):
    lib = capi.load_library()
    n = photons.shape[0]
    src = np.ascontiguousarray(photons).view(np.float32)
    sorted_ph = np.zeros(n, ol.PHOTON_DT)
    keys = np.zeros(n, np.uint32)
    offsets = np.zeros(1000002, np.uint32)
    cell = C.c_float()
    gs = (C.c_uint32 * 3)()
    org = (C.c_float * 3)()
    nv = C.c_uint32()
    PF = C.POINTER(C.c_float)
    rc = lib.opr_build_photon_grid(src.ctypes.data_as(PF), n, sorted_ph.view(np.float32).ctypes.data_as(PF), ol.uptr(keys), ol.uptr(offsets),
                                   C.byref(cell), gs, org, C.byref(nv))
    if rc == 1:
        raise RuntimeError(lib.opr_last_error().decode())
    ncell = gs[0] * gs[1] * gs[2]
    return dict(rc=rc, photons=sorted_ph, keys=keys, offsets=offsets[:ncell + 1].copy() if rc == 0 else None, cellSize=np.float32(cell.value),
                gridSize=tuple(gs), origo=np.array(list(org), np.float32), numValid=nv.value)


def camera_rays(cam, w, h, rng):
    """Primary rays of RayGeneratorPPM.cu:44-52 for random jitter, as n x 8 float32 (o, tmin, d, tmax)."""
    xs, ys = np.meshgrid(np.arange(w, dtype=np.float32), np.arange(h, dtype=np.float32))
    jx, jy = rng.random((h, w), dtype=np.float32), rng.random((h, w), dtype=np.float32)
    dx = (xs + jx) / np.float32(w) * 2 - 1
    dy = (ys + jy) / np.float32(h) * 2 - 1
    u, v, l = (np.array(list(getattr(cam, k)), np.float32) for k in ("camera_u", "camera_v", "lookdir"))
    d = dx[..., None] * u + dy[..., None] * v + l
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    rays = np.zeros((h * w, 8), np.float32)
    rays[:, 0:3] = np.array(list(cam.eye), np.float32)
    rays[:, 3] = 0.001
    rays[:, 4:7] = d.reshape(-1, 3)
    rays[:, 7] = 1e27
    return rays


def random_rays_in_box(n, lo, hi, rng, tmin=1e-4, tmax=1e27):
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0:3] = (rng.random((n, 3)) * (np.array(hi) - np.array(lo)) + np.array(lo)).astype(np.float32)
    d = rng.normal(size=(n, 3))
    rays[:, 4:7] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    rays[:, 3] = tmin
    rays[:, 7] = tmax
    return rays


def rel_close_fraction(a, b, rtol, atol):
    """Fraction of pixels whose every channel satisfies |a-b| <= atol + rtol*|b|."""
    ok = np.abs(a - b) <= atol + rtol * np.abs(b)
    return float(ok.reshape(ok.shape[0] if ok.ndim == 2 else -1, -1).all(axis=-1).mean()) if ok.ndim > 1 else float(ok.mean())
