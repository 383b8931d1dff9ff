"""Render a built-in scene with the CPU oracle (test infrastructure) and write a PNG."""
import argparse, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import oracle_lib as ol
from oppositerenderer_b200 import capi
from scripts.imgio import write_png, display

ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="Cornell"); ap.add_argument("--config", type=int, default=0)
ap.add_argument("--method", default="ppm"); ap.add_argument("--size", type=int, default=256)
ap.add_argument("--photons", type=int, default=316); ap.add_argument("--iters", type=int, default=4)
ap.add_argument("--which", default="oracle"); ap.add_argument("--out", default="/tmp/ora.png")
ap.add_argument("--maxlen", type=int, default=10)
a = ap.parse_args()
sc = ol.BuiltinScene(a.scene, a.config, a.which)
r = ol.OracleRenderer(a.which, (a.photons, a.photons), a.maxlen)
r.init_scene(sc.desc)
cam = sc.camera(1.0)
method = capi.PROGRESSIVE_PHOTON_MAPPING if a.method == "ppm" else capi.VCM_BIDIRECTIONAL_PATH_TRACING
det = ol.make_details(cam, method, a.size, a.size)
radii = ol.ppm_radius_schedule(sc.initial_radius(), a.iters)
t0 = time.time()
for i in range(a.iters):
    r.render_next_iteration(i, i, radii[i], det)
    print("iter", i, "radius", radii[i], "t", time.time() - t0, flush=True)
out = r.output()
print("mean", out.mean(axis=(0, 1)), "rays", r.ray_counts())
write_png(a.out, display(out, a.iters))
