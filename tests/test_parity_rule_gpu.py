"""GPU parity under the explicit rule of tests/parity_rules.py: the CUDA path (through the C ABI) and the CPU oracle render the
same iterations in lock-step with the reference's per-pixel XORWOW seeding; every pixel of every iteration's buffers must be
within the stated tolerance (rel 2e-3 + 1e-5 of the image maximum) OR be explained by a measured divergence of a discrete
decision (RNG state, hit flags, photon present in only one map, occlusion mask / count, splat of a diverged light path).
No pass-fraction is accepted; the fractions are printed (-s) and recorded in profiles/ by scripts/parity_report.py.

Covers BASELINE configs 1 and 2 at their stated sizes and the bench scenes (configs 3/4/5) at reduced tessellation."""
import numpy as np
import pytest

import oracle_lib as ol
import parity_rules as pr
import util
from oppositerenderer_b200 import capi

pytestmark = pytest.mark.gpu

PPM, BDPT, PT = capi.PROGRESSIVE_PHOTON_MAPPING, capi.VCM_BIDIRECTIONAL_PATH_TRACING, capi.PATH_TRACING


def _builtin(product, oracle, name, config, W, H, pl, maxlen):
    scene = product.Cornell() if name == "Cornell" else product.CornellSmall(config or None)
    cam = scene.getDefaultCamera()
    cam.setAspectRatio(W / H)
    osc = oracle.BuiltinScene(name, config or 0)
    r = product.OptixRenderer(photon_launch=pl, vcm_max_path_length=maxlen, seed=util.SEED)
    r.initialize(product.ComputeDevice(0))
    r.initScene(scene)
    o = oracle.OracleRenderer("oracle", pl, maxlen, util.SEED)
    o.init_scene(osc.desc)
    return scene, cam, osc.camera(W / H), r, o, osc


def _synthetic(product, oracle, kind, detail, W, H, pl, maxlen, texsize=64):
    from oppositerenderer_b200 import synthetic
    scene = synthetic.make_scene(kind, seed=1, detail=detail, texture_size=texsize)
    cam = scene.getDefaultCamera()
    cam.setAspectRatio(W / H)
    r = product.OptixRenderer(photon_launch=pl, vcm_max_path_length=maxlen, seed=util.SEED)
    r.initialize(product.ComputeDevice(0))
    r.initScene(scene)
    o = oracle.OracleRenderer("oracle", pl, maxlen, util.SEED)
    o.init_scene(scene.getSceneDescription())
    return scene, cam, cam.pod, r, o, None


def _lockstep(product, oracle, name, scene, cam, ocam, r, o, method, W, H, pl, radii, **kw):
    det = product.RenderServerRenderRequestDetails(cam, scene.getSceneName(), method, W, H, 2.0 / 3.0)
    odet = oracle.make_details(ocam, method, W, H)
    res = pr.run_lockstep(name, r, o, det, odet, method, W, H, pl, radii, ocam, **kw)
    st, (c0, a0) = r.stats(), o.ray_counts()
    assert abs(st["closest_rays"] - c0) <= 0.02 * c0 + 8 and abs(st["any_rays"] - a0) <= 0.02 * a0 + 8, (st, c0, a0)
    assert res["rel_l1"] < (6e-2 if method == PT else 3e-2), res     # PATH_TRACING: five dependent bounces, a flipped path changes its pixel entirely
    return res


CORNELL_SMALL = [0, (1 << 1) | (1 << 7) | (1 << 11) | (1 << 9), (1 << 3) | (1 << 8) | (1 << 12) | (1 << 13), (1 << 5) | (1 << 10)]


@pytest.mark.parametrize("name,config", [("Cornell", 0)] + [("CornellSmall", c) for c in CORNELL_SMALL])
def test_ppm_every_pixel_in_tolerance_or_explained(product, oracle, name, config):
    W = H = 64
    pl = (64, 64)
    scene, cam, ocam, r, o, osc = _builtin(product, oracle, name, config, W, H, pl, 10)
    r0 = scene.getSceneInitialPPMRadiusEstimate() if name == "Cornell" else 0.05   # the extent estimate (1.7e-4) finds no photon at this count
    _lockstep(product, oracle, "%s/%d PPM" % (name, config), scene, cam, ocam, r, o, PPM, W, H, pl, product.ppm_radius_schedule(r0, 3))


@pytest.mark.parametrize("config,maxlen", [(CORNELL_SMALL[0], 8), (CORNELL_SMALL[1], 8), (CORNELL_SMALL[2], 10), (CORNELL_SMALL[3], 10)])
def test_bdpt_every_pixel_in_tolerance_or_explained(product, oracle, config, maxlen):
    W = H = 48
    scene, cam, ocam, r, o, osc = _builtin(product, oracle, "CornellSmall", config, W, H, (16, 16), maxlen)
    counts = []

    def light_vertex_counts(i, a, b):
        cnt = r.debug_buffer("light_vertex_counts", np.uint32, W * H)
        cnt0 = np.zeros(W * H, np.uint32)
        o.lib.ora_get_light_vertex_counts(o.h, cnt0.ctypes.data_as(__import__("ctypes").c_void_p))
        state = pr.pixel_state_divergence(a, b)
        assert (cnt == cnt0)[~state].all()          # a light subpath whose RNG stream did not diverge stores the same number of vertices
        counts.append(float((cnt == cnt0).mean()))
    _lockstep(product, oracle, "CornellSmall/%d BDPT" % config, scene, cam, ocam, r, o, BDPT, W, H, (16, 16),
              product.ppm_radius_schedule(scene.getSceneInitialPPMRadiusEstimate(), 3), per_iteration=light_vertex_counts)
    assert min(counts) > 0.85


@pytest.mark.parametrize("name,config", [("Cornell", 0), ("CornellSmall", CORNELL_SMALL[1]), ("CornellSmall", CORNELL_SMALL[2])])
def test_path_tracing_every_pixel_in_tolerance_or_explained(product, oracle, name, config):
    W = H = 48
    scene, cam, ocam, r, o, osc = _builtin(product, oracle, name, config, W, H, (16, 16), 10)
    _lockstep(product, oracle, "%s/%d PT" % (name, config), scene, cam, ocam, r, o, PT, W, H, (16, 16), [1.0, 1.0, 1.0])


@pytest.mark.parametrize("kind,detail,method", [("sponza", 0.08, PPM), ("conference", 0.12, PPM), ("sponza", 0.08, BDPT), ("conference", 0.12, BDPT),
                                                 ("sponza", 0.08, PT), ("conference", 0.12, PT)])
def test_bench_scenes_every_pixel_in_tolerance_or_explained(product, oracle, kind, detail, method):
    """The bench workloads (BASELINE configs 3/4/5: triangle meshes, Texture materials with normal maps, glass / mirror) at
    reduced tessellation, all three methods."""
    W, H = 96, 54
    pl = (96, 96) if method == PPM else (16, 16)
    scene, cam, ocam, r, o, _ = _synthetic(product, oracle, kind, detail, W, H, pl, 8)
    radii = product.ppm_radius_schedule(scene.getSceneInitialPPMRadiusEstimate(), 2)
    _lockstep(product, oracle, "synthetic %s method %d" % (kind, method), scene, cam, ocam, r, o, method, W, H, pl, radii)
    if method == PPM:
        hp = r.debug_buffer("hitpoints", ol.HITPOINT_DT, W * H)
        assert not np.isnan(hp["normal"]).any()


def test_baseline_config_1_as_stated(product, oracle):
    """BASELINE config 1 exactly: built-in Cornell, PPM, 256 x 256, 100 K photons per iteration (316 x 316 launch), 16 iterations —
    every iteration under the rule, then the sum buffer against the committed output of the reference's own programs."""
    import os
    from golden.make_golden import CONFIG1
    W = H = 256
    pl = (316, 316)
    scene, cam, ocam, r, o, osc = _builtin(product, oracle, "Cornell", 0, W, H, pl, 10)
    radii = product.ppm_radius_schedule(scene.getSceneInitialPPMRadiusEstimate(), 16)
    res = _lockstep(product, oracle, "config 1: Cornell PPM 256x256 100K photons", scene, cam, ocam, r, o, PPM, W, H, pl, radii, max_explained=0.25)
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", CONFIG1[0] + ".npz"))
    out = r.getOutputBuffer()
    rel = np.abs(out - gold["output"]).sum() / np.abs(gold["output"]).sum()
    assert rel < 1e-2, rel
    st = r.stats()
    rays = st["closest_rays"] + st["any_rays"]
    assert abs(rays - int(gold["rays"].sum())) <= 0.005 * int(gold["rays"].sum())
    print("config 1 vs reference golden: relative L1 %.2e, rays %d vs %d; %s" % (rel, rays, int(gold["rays"].sum()), res["lines"][-1]))


def test_baseline_config_2_as_stated(product, oracle):
    """BASELINE config 2 exactly: CornellSmall, BDPT (vertex connection, recursive MIS), max path length 8, 512 x 512."""
    W = H = 512
    scene, cam, ocam, r, o, osc = _builtin(product, oracle, "CornellSmall", 0, W, H, (16, 16), 8)
    radii = product.ppm_radius_schedule(scene.getSceneInitialPPMRadiusEstimate(), 2)
    _lockstep(product, oracle, "config 2: CornellSmall BDPT 512x512 len 8", scene, cam, ocam, r, o, BDPT, W, H, (16, 16), radii)


def test_normal_map_without_tangents_matches_oracle(product, oracle):
    """Undefined in the reference (Texture.cu:89-95 reads attributes TriangleMesh.cu:61-72 never wrote); defined here and in the
    oracle as 'the map is ignored': no NaN normal, and the rule holds."""
    from oppositerenderer_b200 import synthetic
    W, H, pl = 96, 54, (64, 64)
    scene = synthetic.make_scene("sponza", seed=1, detail=0.08, texture_size=32)
    d = scene.getSceneDescription().contents
    for i in range(d.numMeshes):
        d.meshes[i].tangents = d.meshes[i].bitangents = None
    cam = scene.getDefaultCamera()
    cam.setAspectRatio(W / H)
    r = product.OptixRenderer(photon_launch=pl, seed=util.SEED)
    r.initialize(product.ComputeDevice(0))
    r.initScene(scene)
    o = oracle.OracleRenderer("oracle", pl, 10, util.SEED)
    o.init_scene(scene.getSceneDescription())
    _lockstep(product, oracle, "normal map without tangents", scene, cam, cam.pod, r, o, PPM, W, H, pl,
              product.ppm_radius_schedule(scene.getSceneInitialPPMRadiusEstimate(), 2))
    assert not np.isnan(r.debug_buffer("hitpoints", ol.HITPOINT_DT, W * H)["normal"]).any()


@pytest.mark.parametrize("kind,config_no", [("sponza", 4), ("conference", 3)])
def test_baseline_bench_configs_as_stated(product, oracle, kind, config_no):
    """BASELINE configs 3 and 4 AS STATED — the synthetic ~300 K-triangle Conference-scale and ~1 M-triangle Sponza-scale scenes, PPM,
    1920 x 1080, 1024 x 1024 photons — one iteration in lock-step with the oracle (all host cores):
      * hit flags equal except where the camera path's RNG stream diverged; photon count and ray counts within 0.5 %;
      * the gather checked against the oracle's gather on the product's own 2 M hit points and ~1.8 M sorted photons: every pixel;
      * direct light: every out-of-tolerance pixel has a diverged camera path, a photon thread of the same RNG stream that walked differently, a
        different occlusion mask, a measurably different normal / attenuation (texture unit) or lies in the plane of an area light;
      * output == direct + indirect bit for bit; whole image within 1 % (clipped relative L1)."""
    import os
    from oppositerenderer_b200 import synthetic
    W, H, pl = 1920, 1080, (1024, 1024)
    scene = synthetic.make_scene(kind, seed=1, detail=1.0)
    assert scene.num_triangles > (1000000 if kind == "sponza" else 250000)
    cam = scene.getDefaultCamera()
    cam.setAspectRatio(W / H)
    r = product.OptixRenderer(photon_launch=pl, seed=util.SEED)
    r.initialize(product.ComputeDevice(0))
    r.initScene(scene)
    o = oracle.OracleRenderer("oracle", pl, 10, util.SEED, threads=len(os.sched_getaffinity(0)))
    o.init_scene(scene.getSceneDescription())
    det = product.RenderServerRenderRequestDetails(cam, scene.getSceneName(), PPM, W, H, 2.0 / 3.0)
    odet = oracle.make_details(cam.pod, PPM, W, H)
    radius = product.ppm_radius_schedule(scene.getSceneInitialPPMRadiusEstimate(), 1)[0]
    r.renderNextIteration(0, 0, radius, True, det)
    o.render_next_iteration(0, 0, radius, odet)
    a, b = pr.ProductProbe(r, W, H, pl), pr.OracleProbe(o, W, H, pl)
    cam_div, hpa, hpb = pr.camera_divergence(a, b)
    state = pr.pixel_state_divergence(a, b)
    # a camera path whose RNG stream did not diverge ends with the same flags, but for rays that graze an edge between two materials which
    # both draw nothing (mirror | diffuse): a handful in two million
    assert ((hpa["flags"] != hpb["flags"]) & ~state).sum() <= 1e-5 * W * H
    assert not np.isnan(hpa["normal"]).any()
    st, (c0, a0) = r.stats(), o.ray_counts()
    nv, nv0 = int(r.debug_buffer("grid", np.uint32, 8)[7]), o.grid()["numValid"]
    assert abs(nv - nv0) <= 0.005 * nv0 and abs(st["closest_rays"] - c0) <= 0.005 * c0 and abs(st["any_rays"] - a0) <= 0.005 * a0, (nv, nv0, st, c0, a0)
    worst = pr.gather_stage_exact(a, radius)                              # every one of the 2 M pixels, no exception
    da, db = a.direct(), b.direct()
    bad_direct = pr.out_of_tolerance(da, db)
    diffuse = (hpb["flags"] & ol.PRD_HIT_NON_SPECULAR) != 0
    explained = cam_div | (diffuse & (a.direct_sample_mask() != b.direct_sample_mask())) | (diffuse & pr.shading_inputs_differ(hpa, hpb)) | \
        (diffuse & pr.coplanar_with_area_light(scene.getSceneDescription(), hpb)) | (diffuse & pr.photon_thread_divergence(a, b))
    for p in np.nonzero(bad_direct & ~explained)[0][:4]:      # what a person needs to name the difference, should the rule ever fail here
        print("pixel %d direct %s / %s pos %s / %s normal %s / %s att %s / %s mask %x / %x" % (p, da[p], db[p], hpa["position"][p], hpb["position"][p], hpa["normal"][p],
              hpb["normal"][p], hpa["attenuation"][p], hpb["attenuation"][p], a.direct_sample_mask()[p], b.direct_sample_mask()[p]))
    line = pr.check_explained("config %d direct" % config_no, bad_direct, explained)
    ia, ib = a.indirect(), b.indirect()
    bad_ind = pr.out_of_tolerance(ia, ib)
    out, out0 = a.output(), b.output()
    assert ((da + ia).view(np.uint32) == out.view(np.uint32)).all()     # PPM_OUTPUT_PASS, first iteration: output = direct + indirect
    clip = float(np.percentile(out0, 99.0)) + 1e-30
    rel = float(np.abs(np.minimum(out, clip) - np.minimum(out0, clip)).sum() / np.abs(np.minimum(out0, clip)).sum())
    print("config %d (%s, %d triangles): camera paths diverged %.4f %%, photons %d vs %d, gather vs oracle gather on the product's inputs worst rel %.1e, "
          "indirect out of the plain tolerance %.2f %% (photon-map differences, quantified at reduced size), clipped relative L1 of the image %.2e; %s"
          % (config_no, kind, scene.num_triangles, 100 * cam_div.mean(), nv, nv0, worst, 100 * bad_ind.mean(), rel, line))
    assert cam_div.mean() < 0.02 and rel < 1e-2


def test_baseline_config_5_workload(product, oracle):
    """BASELINE config 5's workload — BDPT (vertex connection, recursive MIS, max path length 8) on the ~1 M-triangle scene — at 960 x 540
    (at 3840 x 2160 the oracle needs ~20 s per iteration and the light-vertex dump of the rule 6 GB; the 4K run itself is
    bench.py --method bdpt --width 3840 --height 2160), two iterations in lock-step under the per-pixel rule."""
    import os
    from oppositerenderer_b200 import synthetic
    W, H = 960, 540
    scene = synthetic.make_scene("sponza", seed=1, detail=1.0)
    cam = scene.getDefaultCamera()
    cam.setAspectRatio(W / H)
    r = product.OptixRenderer(photon_launch=(16, 16), vcm_max_path_length=8, seed=util.SEED)
    r.initialize(product.ComputeDevice(0))
    r.initScene(scene)
    o = oracle.OracleRenderer("oracle", (16, 16), 8, util.SEED, threads=len(os.sched_getaffinity(0)))
    o.init_scene(scene.getSceneDescription())
    radii = product.ppm_radius_schedule(scene.getSceneInitialPPMRadiusEstimate(), 2)
    # half a million pixels on the full scene: 4 per million may stay unnamed (printed; the reduced-size cases above admit none)
    _lockstep(product, oracle, "config 5 workload: BDPT 960x540, 1.1 M triangles", scene, cam, cam.pod, r, o, BDPT, W, H, (16, 16), radii, allow=max(1, int(4e-6 * W * H)))
