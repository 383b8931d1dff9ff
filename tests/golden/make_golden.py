"""Generates the golden fixtures in this directory by running the REFERENCE's own device programs (host-compiled
behind oracle/refshim into oracle/_ref/liboppref.so) on the built-in scenes.  Only possible where /root/reference
exists (the build container); the fixtures travel, the reference does not.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_lib as ol  # noqa: E402

SEED = 1645301512
CASES = [  # name, scene, CornellSmall config, method (2 = PPM, 1 = BDPT, 0 = PT), size, photon launch, iterations, max path length, fixed r0
    ("cornell_ppm_64x64_p64x64_it3", "Cornell", 0, 2, 64, 64, 3, 10, None),
    ("cornellsmall_ppm_48x48_p48x48_it2", "CornellSmall", 0, 2, 48, 48, 2, 10, 0.05),
    ("cornellsmall_mirrorfloor_glass_ppm_48x48_p48x48_it2", "CornellSmall", (1 << 1) | (1 << 7) | (1 << 11) | (1 << 9), 2, 48, 48, 2, 10, 0.05),
    ("cornellsmall_point_glossy_spheres_ppm_48x48_p48x48_it2", "CornellSmall", (1 << 3) | (1 << 8) | (1 << 12) | (1 << 13), 2, 48, 48, 2, 10, 0.05),
    ("cornellsmall_bdpt_48x48_it3_len8", "CornellSmall", 0, 1, 48, 48, 3, 8, None),
    ("cornellsmall_mirrorfloor_glass_bdpt_48x48_it2_len8", "CornellSmall", (1 << 1) | (1 << 7) | (1 << 11) | (1 << 9), 1, 48, 48, 2, 8, None),
    ("cornellsmall_point_glossy_spheres_bdpt_48x48_it2_len10", "CornellSmall", (1 << 3) | (1 << 8) | (1 << 12) | (1 << 13), 1, 48, 48, 2, 10, None),
    # method 0 = PATH_TRACING (pt/RayGeneratorPT.cu)
    ("cornell_pt_64x64_it3", "Cornell", 0, 0, 64, 64, 3, 10, None),
    ("cornellsmall_mirrorfloor_glass_pt_48x48_it3", "CornellSmall", (1 << 1) | (1 << 7) | (1 << 11) | (1 << 9), 0, 48, 48, 3, 10, None),
    ("cornellsmall_point_glossy_spheres_pt_48x48_it3", "CornellSmall", (1 << 3) | (1 << 8) | (1 << 12) | (1 << 13), 0, 48, 48, 3, 10, None),
]


# The bench workloads (BASELINE configs 3/4/5) at reduced tessellation: triangle meshes with shading normals / uvs / tangent
# frames, Texture materials with RGBA8 bilinear lookups and normal maps, tessellated glass + mirror spheres, area + point
# lights (TriangleMesh.cu:35-96, Texture.cu:83-110,116-175,210-286).
# name, kind, detail, texture size, method, width, height, photon launch, iterations, max path length
SYNTH_CASES = [
    ("synthetic_sponza_ppm_96x54_p96x96_it2", "sponza", 0.08, 64, 2, 96, 54, 96, 2, 10),
    ("synthetic_sponza_bdpt_96x54_it2_len8", "sponza", 0.08, 64, 1, 96, 54, 16, 2, 8),
    ("synthetic_sponza_pt_96x54_it2", "sponza", 0.08, 64, 0, 96, 54, 16, 2, 10),
    ("synthetic_conference_ppm_96x54_p96x96_it2", "conference", 0.12, 64, 2, 96, 54, 96, 2, 10),
    ("synthetic_conference_bdpt_96x54_it2_len8", "conference", 0.12, 64, 1, 96, 54, 16, 2, 8),
    ("synthetic_conference_pt_96x54_it2", "conference", 0.12, 64, 0, 96, 54, 16, 2, 10),
]

# BASELINE config 1 exactly: built-in Cornell, PPM, 256x256, 100 K photons per iteration (316 x 316 launch), 16 iterations
CONFIG1 = ("cornell_ppm_256x256_p316x316_it16", "Cornell", 0, 2, 256, 316, 16, 10, None)


def render_synthetic(which, kind, detail, texsize, method, W, H, pl, iters, maxlen):
    """One of the synthetic bench scenes through the oracle ('oracle') or the host-compiled reference programs ('ref')."""
    from oppositerenderer_b200 import synthetic
    scene = synthetic.make_scene(kind, seed=1, detail=detail, texture_size=texsize)
    cam = scene.getDefaultCamera()
    cam.setAspectRatio(W / H)
    r = ol.OracleRenderer(which, (pl, pl), maxlen, SEED, threads=1 if (which == "ref" and method == 1) else 0)
    r.init_scene(scene.getSceneDescription())
    det = ol.make_details(cam.pod, method, W, H)
    radii = ol.ppm_radius_schedule(scene.getSceneInitialPPMRadiusEstimate(), iters)
    for i in range(iters):
        r.render_next_iteration(i, i, radii[i], det)
    out = dict(output=r.output(), rays=np.array(r.ray_counts(), np.uint64))
    if method == 2:
        hp = r.hitpoints()
        out["photon_keys"] = r.photon_keys()
        out["offsets"] = r.offsets()
        out["hitpoint_flags"] = hp["flags"]
        out["hitpoint_normals"] = hp["normal"]
    return out


def render(which, scene, cfg, method, size, pl, iters, maxlen, r0):
    sc = ol.BuiltinScene(scene, cfg, which)
    # the reference's light-tracing splat is a plain `+=` (vcm.h:159): run its programs on one thread so that no update is lost
    r = ol.OracleRenderer(which, (pl, pl), maxlen, SEED, threads=1 if (which == "ref" and method == 1) else 0)
    r.init_scene(sc.desc)
    det = ol.make_details(sc.camera(1.0), method, size, size)
    radii = ol.ppm_radius_schedule(sc.initial_radius() if r0 is None else r0, iters)
    for i in range(iters):
        r.render_next_iteration(i, i, radii[i], det)
    out = dict(output=r.output(), rays=np.array(r.ray_counts(), np.uint64))
    if method == 2:
        out["photon_keys"] = r.photon_keys()
        out["offsets"] = r.offsets()
        out["hitpoint_flags"] = r.hitpoints()["flags"]
    return out


if __name__ == "__main__":
    assert ol.have_ref(), "build oracle/_ref first: make -C oracle/refshim"
    for case in SYNTH_CASES:
        data = render_synthetic("ref", *case[1:])
        np.savez_compressed(os.path.join(HERE, case[0] + ".npz"), **data)
        print(case[0], "rays", data["rays"], "mean", data["output"].mean())
    for case in CASES + [CONFIG1]:
        data = render("ref", *case[1:])
        np.savez_compressed(os.path.join(HERE, case[0] + ".npz"), **data)
        print(case[0], "rays", data["rays"], "mean", data["output"].mean())
