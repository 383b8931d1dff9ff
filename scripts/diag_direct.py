"""Diagnostics: the four direct-light samples (factor, light index) of chosen pixels of BASELINE config 4 next to the oracle's
direct term.   python scripts/diag_direct.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oppositerenderer_b200 as product, oracle_lib as ol, parity_rules as pr, util
from oppositerenderer_b200 import synthetic
W, H, pl = 1920, 1080, (1024, 1024)
scene = synthetic.make_scene("sponza", seed=1, detail=1.0)
cam = scene.getDefaultCamera(); cam.setAspectRatio(W / H)
r = product.OptixRenderer(photon_launch=pl, seed=util.SEED); r.initialize(product.ComputeDevice(0)); r.initScene(scene)
o = ol.OracleRenderer("oracle", pl, 10, util.SEED, threads=len(os.sched_getaffinity(0))); o.init_scene(scene.getSceneDescription())
det = product.RenderServerRenderRequestDetails(cam, scene.getSceneName(), 2, W, H, 2.0 / 3.0)
odet = ol.make_details(cam.pod, 2, W, H)
radius = product.ppm_radius_schedule(scene.getSceneInitialPPMRadiusEstimate(), 1)[0]
r.renderNextIteration(0, 0, radius, True, det); o.render_next_iteration(0, 0, radius, odet)
a, b = pr.ProductProbe(r, W, H, pl), pr.OracleProbe(o, W, H, pl)
sf = r.debug_buffer("sample_factors", np.float32, W * H * 4).reshape(-1, 4)
sl = r.debug_buffer("sample_lights", np.uint16, W * H * 4).reshape(-1, 4)
hpa, hpb = a.hitpoints(), b.hitpoints()
da, db = a.direct(), b.direct()
d = scene.getSceneDescription().contents
for i in range(d.numLights):
    L = d.lights[i]; print("light", i, "type", L.lightType, "power", list(L.power), "pos", list(L.position), "v1", list(L.v1), "v2", list(L.v2), "normal", list(L.normal))
sa, sb = a.random_states(), b.random_states()
for p in (1918889, 1938242):
    x, y = p % W, p // W
    print("pixel", p, "gpu direct", da[p], "oracle", db[p], "factors", sf[p], "lights", sl[p], "pos", hpa["position"][p], hpb["position"][p], "n", hpa["normal"][p], "att", hpa["attenuation"][p], hpb["attenuation"][p],
          "flags %08x %08x" % (hpa["flags"][p], hpb["flags"][p]), "state equal", (sa[y * 1920 + x] == sb[y * 1920 + x]).all())
    L = d.lights[1]; lp = np.array(list(L.position), np.float32)
    t = lp - hpa["position"][p]; dist = np.linalg.norm(t); nl = max(0.0, float(np.dot(hpa["normal"][p], t / dist)))
    print("   point light: dist %.4f n.l %.6f factor %.6g -> 2 * P * f = %s" % (dist, nl, 0.25 * nl / (np.pi * dist * dist), 2 * np.array(list(L.power)) * 0.25 * nl / (np.pi * dist * dist)))
