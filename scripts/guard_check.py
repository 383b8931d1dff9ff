"""Out-of-bounds write check of every kernel family of the path, with the library's guard bands (include/opr.h:
opr_debug_set_guards / opr_debug_check_guards) — compute-sanitizer is not available on the GPU pool.

    python scripts/guard_check.py

Small inputs, so that an index that runs past a buffer lands in the band rather than in a neighbouring allocation: BVH build
(Cornell parallelograms + a triangle mesh through PLOC), PPM, vertex-connection BDPT and path tracing iterations with the stream
pipeline on, odd image sizes, the stand-alone grid build and gather.  Product only: no oracle, no torch.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oppositerenderer_b200 as opr  # noqa: E402


def render(scene, method, size, iters, pl=(48, 48), maxlen=0):
    cam = scene.getDefaultCamera()
    cam.setAspectRatio(size[0] / size[1])
    r = opr.OptixRenderer(photon_launch=pl, vcm_max_path_length=maxlen, seed=1645301512)
    r.initialize(opr.ComputeDevice(0))
    r.initScene(scene)
    det = opr.RenderServerRenderRequestDetails(cam, scene.getSceneName(), method, size[0], size[1], 2.0 / 3.0)
    radii = opr.ppm_radius_schedule(scene.getSceneInitialPPMRadiusEstimate(), iters)
    for i in range(iters):
        r.renderNextIteration(i, i, radii[i], True, det)
    out = r.getOutputBuffer()
    assert np.isfinite(out).all()
    stats = r.stats()
    check_guards(str(method) + " " + scene.getSceneName())      # while the renderer's buffers are still live
    r.close()
    return float(out.mean()), stats


CHECKS = []


def check_guards(what):
    import ctypes as C
    lib = opr.capi.load_library()
    live = C.c_int(0)
    bad = lib.opr_debug_check_guards(C.byref(live))
    CHECKS.append((what, bad, live.value))
    return bad


def main(guards=True, verbose=True):
    """Returns [(what, damaged allocations, guarded allocations checked)]."""
    del CHECKS[:]
    opr.capi.load_library().opr_debug_set_guards(1 if guards else 0)
    try:
        _run(verbose)
    finally:
        opr.capi.load_library().opr_debug_set_guards(0)
    return list(CHECKS)


def _run(verbose):
    say = print if verbose else (lambda *a, **k: None)
    M = opr.RenderMethod
    glass = (1 << 1) | (1 << 7) | (1 << 11)
    jobs = [("Cornell PPM", opr.Cornell(), M.PROGRESSIVE_PHOTON_MAPPING, (73, 41), 2, 0),
            ("CornellSmall PPM", opr.CornellSmall(glass), M.PROGRESSIVE_PHOTON_MAPPING, (64, 64), 2, 0),
            ("CornellSmall BDPT", opr.CornellSmall(), M.VCM_BIDIRECTIONAL_PATH_TRACING, (64, 48), 2, 0),
            ("CornellSmall BDPT len 4", opr.CornellSmall(glass), M.VCM_BIDIRECTIONAL_PATH_TRACING, (40, 40), 2, 4),
            ("Cornell PT", opr.Cornell(), M.PATH_TRACING, (48, 48), 2, 0)]
    for name, scene, method, size, iters, maxlen in jobs:
        mean, stats = render(scene, method, size, iters, maxlen=maxlen)
        say("%-26s mean %.5f stats %s" % (name, mean, stats), flush=True)

    # a triangle-mesh scene through the PLOC builder (the bench generator at low detail)
    from oppositerenderer_b200 import synthetic
    scene = synthetic.make_scene("sponza", seed=1, detail=0.02)
    for method, nm in ((M.PROGRESSIVE_PHOTON_MAPPING, "mesh PPM"), (M.VCM_BIDIRECTIONAL_PATH_TRACING, "mesh BDPT")):
        mean, stats = render(scene, method, (64, 36), 2)
        say("%-26s mean %.5f triangles %d stats %s" % (nm, mean, scene.getNumTriangles(), stats), flush=True)

    # stand-alone grid build + gather + trace entry points
    import ctypes as C
    from oppositerenderer_b200 import capi
    lib = capi.load_library()
    PF, PU = C.POINTER(C.c_float), C.POINTER(C.c_uint32)
    rng = np.random.default_rng(5)
    n = 20000
    ph = np.zeros((n, 9), np.float32)                       # ppm/Photon.h: power, position, rayDirection
    ph[:, 0:3] = rng.random((n, 3), dtype=np.float32)
    ph[:, 3:6] = rng.random((n, 3), dtype=np.float32) * 500.0
    ph[:, 4] *= 0.2
    ph[:, 6:9] = (0.0, -1.0, 0.0)
    ph[rng.random(n) < 0.3] = 0.0
    sorted_ph, keys, offsets = np.zeros((n, 9), np.float32), np.zeros(n, np.uint32), np.zeros(1000002, np.uint32)
    cell, gs, org, nv = C.c_float(), (C.c_uint32 * 3)(), (C.c_float * 3)(), C.c_uint32()
    rc = lib.opr_build_photon_grid(ph.ctypes.data_as(PF), n, sorted_ph.ctypes.data_as(PF), keys.ctypes.data_as(PU), offsets.ctypes.data_as(PU),
                                   C.byref(cell), gs, org, C.byref(nv))
    assert rc == 0, lib.opr_last_error()
    npix = 3000
    hp = np.zeros((npix, 16), np.float32)                   # ppm/Hitpoint.h: position, normal, attenuation, radiance, flags, volumetric
    hp[:, 0:3] = rng.random((npix, 3), dtype=np.float32) * 500.0
    hp[:, 1] *= 0.2
    hp[:, 3:6] = (0.0, 1.0, 0.0)
    hp[:, 6:9] = 1.0
    hp.view(np.uint32)[:, 12] = 1 << 27                     # PRD_HIT_NON_SPECULAR
    indirect = np.zeros((npix, 3), np.float32)
    rc = lib.opr_gather(npix, hp.ctypes.data_as(PF), nv.value, sorted_ph.ctypes.data_as(PF), offsets.ctypes.data_as(PU), cell.value, gs, org,
                        C.c_float(25.0), C.c_float(float(n)), indirect.ctypes.data_as(PF))
    assert rc == 0, lib.opr_last_error()
    check_guards("stand-alone grid build + gather, BVH build temporaries")
    say("grid %s valid %d; gather mean %.5g" % (tuple(gs), nv.value, indirect.mean()), flush=True)


if __name__ == "__main__":
    res = main()
    for what, bad, live in res:
        print("guards: %-40s damaged %d of %d" % (what, bad, live))
    sys.exit(1 if any(bad != 0 for _, bad, _ in res) else 0)
