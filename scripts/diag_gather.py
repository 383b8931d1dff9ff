#!/usr/bin/env python
"""Diagnostics for the gather stage check (tests/parity_rules.py gather_stage_exact) at full size: renders one PPM iteration of a bench
scene, runs the oracle's gather on the product's own inputs and prints, for the worst pixels, both values and the photons in range.
   python scripts/diag_gather.py sponza|conference"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oppositerenderer_b200 as product  # noqa: E402
import oracle_lib as ol  # noqa: E402
import parity_rules as pr  # noqa: E402
import util  # noqa: E402
from oppositerenderer_b200 import synthetic  # noqa: E402

kind = sys.argv[1]
W, H, pl = 1920, 1080, (1024, 1024)
scene = synthetic.make_scene(kind, seed=1, detail=1.0)
cam = scene.getDefaultCamera()
cam.setAspectRatio(W / H)
r = product.OptixRenderer(photon_launch=pl, seed=util.SEED)
r.initialize(product.ComputeDevice(0))
r.initScene(scene)
det = product.RenderServerRenderRequestDetails(cam, scene.getSceneName(), 2, W, H, 2.0 / 3.0)
radius = product.ppm_radius_schedule(scene.getSceneInitialPPMRadiusEstimate(), 1)[0]
r.renderNextIteration(0, 0, radius, True, det)
a = pr.ProductProbe(r, W, H, pl)
hp = a.hitpoints()
g = r.debug_buffer("grid", np.uint32, 8)
gf = g.view(np.float32)
nv = int(g[7])
ncell = int(g[1]) * int(g[2]) * int(g[3])
sp = r.debug_buffer("sorted_photons", ol.PHOTON_DT, nv)
off = r.debug_buffer("offsets", np.uint32, ncell + 1)
gs = (C.c_uint32 * 3)(int(g[1]), int(g[2]), int(g[3]))
org = (C.c_float * 3)(float(gf[4]), float(gf[5]), float(gf[6]))
out0 = np.zeros((W * H, 3), np.float32)
ol.load("oracle").ora_gather(W * H, ol.fptr(hp.view(np.float32)), nv, ol.fptr(sp.view(np.float32)), ol.uptr(off), float(gf[0]), gs, org, float(radius), float(pl[0] * pl[1]), ol.fptr(out0))
out = a.indirect()
err = (np.abs(out - out0) / (np.abs(out0) + 1e-6 * float(out0.max()))).max(axis=1)
bad = np.nonzero(~np.isclose(out, out0, rtol=1e-4, atol=1e-6 * float(out0.max())).all(axis=1))[0]
print("radius %.6g cell %.6g grid %s photons %d; pixels beyond rel 1e-4: %d of %d; max image value %.4g" % (radius, gf[0], list(gs), nv, bad.size, W * H, out0.max()))
r2 = np.float32(radius) * np.float32(radius)
for p in bad[np.argsort(-err[bad])][:8]:
    pos, nrm = hp["position"][p], hp["normal"][p]
    diff = pos[None, :] - sp["position"]
    d2 = (diff[:, 0] * diff[:, 0] + diff[:, 1] * diff[:, 1]) + diff[:, 2] * diff[:, 2]
    nd = (sp["rayDirection"][:, 0] * nrm[0] + sp["rayDirection"][:, 1] * nrm[1]) + sp["rayDirection"][:, 2] * nrm[2]
    acc = (d2 <= r2) & (nd <= 0)
    w = 1.818 * (1 - (1 - np.exp(-1.953 * d2[acc].astype(np.float64) / (2 * float(r2)))) / (1 - 0.141847))
    rec = (sp["power"][acc].astype(np.float64) * w[:, None]).sum(axis=0) * hp["attenuation"][p] / (np.pi * float(r2) * pl[0] * pl[1])
    k = np.argsort(np.abs(d2 - r2))[:3]
    cellf = (pos - np.array([gf[4], gf[5], gf[6]], np.float32)) / gf[0]
    print("pixel %d (x %d y %d) flags %08x gpu %s oracle-gather %s numpy %s | in range %d accepted %d | closest to sphere d2/r2-1 %s nd %s | smallest |nd| accepted %s | cell coords %s"
          % (p, p % W, p // W, hp["flags"][p], out[p], out0[p], rec, int((d2 <= r2).sum()), int(acc.sum()), (d2[k] / r2 - 1).tolist(), nd[k].tolist(),
             np.sort(np.abs(nd[d2 <= r2]))[:3].tolist(), cellf.tolist()))
