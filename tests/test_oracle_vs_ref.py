"""Pins the oracle: (1) against the committed golden fixtures, which were produced by the reference's own device
programs host-compiled in oracle/_ref (tests/golden/make_golden.py), and (2) where oracle/_ref/liboppref.so is
present, directly against those programs on further scenes.  Integer results (ray counts, photon cell keys, offset
tables, hit flags) must be identical; float images within 1e-5 (the reference's `exp` resolves to the double
overload on the host, the restatement uses expf as CUDA does; splat order differs)."""
import os

import numpy as np
import pytest

import oracle_lib as ol
from oppositerenderer_b200 import capi
from golden.make_golden import CASES, CONFIG1, SYNTH_CASES, render, render_synthetic

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_oracle_reproduces_reference_golden(oracle, case):
    gold = np.load(os.path.join(GOLDEN, case[0] + ".npz"))
    got = render("oracle", *case[1:])
    assert (got["rays"] == gold["rays"]).all()
    assert np.allclose(got["output"], gold["output"], rtol=1e-5, atol=2e-5 * float(gold["output"].max()))
    for k in ("photon_keys", "offsets", "hitpoint_flags"):
        if k in gold.files:
            assert (got[k] == gold[k]).all(), k


@pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("scene,cfg,method,maxlen", [
    ("CornellSmall", (1 << 5) | (1 << 8) | (1 << 10), 1, 10),     # distant point light, glossy floor, large mirror sphere, BDPT
    ("CornellSmall", (1 << 2) | (1 << 6) | (1 << 9) | (1 << 0), 1, 6),   # upward area light, blue back wall, SmallVCM colours, BDPT
    ("CornellSmall", (1 << 4) | (1 << 11), 2, 10),                # strong point light, large glass sphere, PPM
])
def test_oracle_matches_host_compiled_reference(oracle, scene, cfg, method, maxlen):
    a = render("oracle", scene, cfg, method, 40, 40, 2, maxlen, 0.05 if method == 2 else None)
    b = render("ref", scene, cfg, method, 40, 40, 2, maxlen, 0.05 if method == 2 else None)
    assert (a["rays"] == b["rays"]).all()
    assert np.allclose(a["output"], b["output"], rtol=1e-5, atol=2e-5 * float(b["output"].max()))
    if method == 2:
        assert (a["photon_keys"] == b["photon_keys"]).all() and (a["offsets"] == b["offsets"]).all()


def _same(got, gold):
    assert (got["rays"] == gold["rays"]).all(), (got["rays"], gold["rays"])
    assert np.allclose(got["output"], gold["output"], rtol=1e-5, atol=2e-5 * float(gold["output"].max()))
    for k in ("photon_keys", "offsets", "hitpoint_flags"):
        if k in gold:
            assert got[k].shape == gold[k].shape and (got[k] == gold[k]).all(), k


@pytest.mark.parametrize("case", SYNTH_CASES, ids=[c[0] for c in SYNTH_CASES])
def test_oracle_mesh_texture_branch_reproduces_reference_golden(oracle, case):
    """The oracle's triangle-mesh / Texture branch (TriangleMesh.cu:35-96, Texture.cu:83-110,116-175,210-286) on the bench
    scenes at reduced tessellation, all three methods, against the committed output of the reference's own programs: rays,
    photon cell keys, offset table and hit flags identical, image within 1e-5, no NaN shading normal."""
    gold = dict(np.load(os.path.join(GOLDEN, case[0] + ".npz")))
    got = render_synthetic("oracle", *case[1:])
    _same(got, gold)
    if "hitpoint_normals" in gold:
        assert not np.isnan(got["hitpoint_normals"]).any() and not np.isnan(gold["hitpoint_normals"]).any()
        diffuse = (gold["hitpoint_flags"] & ol.PRD_HIT_NON_SPECULAR) != 0
        assert np.allclose(got["hitpoint_normals"][diffuse], gold["hitpoint_normals"][diffuse], atol=1e-6)


@pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("kind,detail,method", [("sponza", 0.1, 2), ("sponza", 0.1, 1), ("sponza", 0.1, 0),
                                                 ("conference", 0.15, 2), ("conference", 0.15, 1), ("conference", 0.15, 0)])
def test_oracle_matches_host_compiled_reference_on_bench_scenes(oracle, kind, detail, method):
    """The same comparison run live against oracle/_ref at another tessellation, image size and photon launch than the fixtures."""
    args = (kind, detail, 32, method, 80, 45, 64 if method == 2 else 16, 2, 8)
    a, b = render_synthetic("oracle", *args), render_synthetic("ref", *args)
    _same(a, b)
    if method == 2:
        assert not np.isnan(a["hitpoint_normals"]).any() and not np.isnan(b["hitpoint_normals"]).any()


def test_oracle_reproduces_reference_at_baseline_config_1(oracle):
    """BASELINE config 1 as stated: built-in Cornell, PPM, 256x256, 100 K photons per iteration (316 x 316), 16 iterations."""
    gold = dict(np.load(os.path.join(GOLDEN, CONFIG1[0] + ".npz")))
    _same(render("oracle", *CONFIG1[1:]), gold)


def test_normal_map_without_tangents_is_ignored_not_nan(oracle):
    """A normal-mapped mesh that carries no tangent frame is undefined behaviour in the reference (Texture.cu:89-95 reads attributes
    TriangleMesh.cu:61-72 never wrote); product and oracle define it: the map is ignored, the interpolated normal is used."""
    from oppositerenderer_b200 import synthetic
    scene = synthetic.make_scene("sponza", seed=1, detail=0.08, texture_size=32)
    d = scene.getSceneDescription().contents
    normal_mapped = {i for i in range(d.numMaterials) if d.materials[i].kind == capi.MAT_TEXTURE and d.materials[i].normalTexture >= 0}
    assert normal_mapped
    stripped = 0
    for i in range(d.numMeshes):
        if d.meshes[i].material in normal_mapped:
            assert d.meshes[i].tangents and d.meshes[i].bitangents      # the generator keeps the frames of every normal-mapped mesh
            d.meshes[i].tangents = d.meshes[i].bitangents = None
            stripped += 1
    assert stripped > 1
    cam = scene.getDefaultCamera()
    cam.setAspectRatio(96 / 54)
    r = ol.OracleRenderer("oracle", (32, 32), 8, 7)
    r.init_scene(scene.getSceneDescription())
    r.render_next_iteration(0, 0, 10.0, ol.make_details(cam.pod, 2, 96, 54))
    hp = r.hitpoints()
    assert not np.isnan(hp["normal"]).any() and np.isfinite(r.output()).all()


@pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("method,with_spot", [(2, True), (0, True), (1, False)])
def test_oracle_matches_host_compiled_reference_on_spot_light_dof_room(oracle, product, method, with_spot):
    """The room of test_spot_light_dof_and_wide_launch_parity (tests/test_parity_gpu.py) — SPOT photon emission
    (PhotonGenerator.cu:73-77), depth of field (camera.h:11-27), Glossy and Mirror on triangle meshes, point + area lights, an image
    wider than the photon launch and a photon launch taller than the image — rendered by the oracle and by the reference's own
    programs: ray counts, photon keys, offsets and hit flags identical, images within 1e-5.  This pins the oracle on the inputs of
    that GPU test.  BDPT runs without the spot light: the reference's lightEmit / lightIlluminate leave position, direction and pdfs
    UNINITIALISED for Light::SPOT (helpers/light.h:136-139,208-211, "Todo"), so its light path starts from whatever the stack held —
    undefined behaviour that no restatement can match (see the next test for what the oracle and the product define instead)."""
    from test_parity_gpu import _spot_room
    scene = _spot_room(product, with_spot)
    cam = scene.getDefaultCamera()
    W, H, pl, iters = 80, 20, (32, 48), 2
    cam.setAspectRatio(W / H)
    out = {}
    for which in ("oracle", "ref"):
        r = ol.OracleRenderer(which, pl, 6, 1645301512, threads=1 if (which == "ref" and method == 1) else 0)
        r.init_scene(scene.getSceneDescription())
        det = ol.make_details(cam.pod, method, W, H)
        for i in range(iters):
            r.render_next_iteration(i, i, 0.3, det)
        out[which] = dict(output=r.output(), rays=np.array(r.ray_counts(), np.uint64))
        if method == 2:
            out[which].update(photon_keys=r.photon_keys(), offsets=r.offsets(), hitpoint_flags=r.hitpoints()["flags"])
    assert out["ref"]["output"].sum() > 0
    _same(out["oracle"], out["ref"])


def test_spot_light_in_bdpt_is_a_path_that_misses(oracle, product):
    """Light::SPOT in the BDPT passes: the reference reads uninitialised values (helpers/light.h:136-139,208-211).  The oracle (and the
    product, which is compared with it on this room on the GPU) define the case: emission position and direction are zero and the
    radiance is zero, so a light path that picked the spot light traces one ray that misses and stores nothing, and a camera vertex
    that picked it for its light connection adds nothing.  The image must stay finite and the other two lights must still show."""
    from test_parity_gpu import _spot_room
    scene = _spot_room(product, True)
    cam = scene.getDefaultCamera()
    W, H, pl = 80, 20, (32, 48)
    cam.setAspectRatio(W / H)
    r = ol.OracleRenderer("oracle", pl, 6, 1645301512)
    r.init_scene(scene.getSceneDescription())
    det = ol.make_details(cam.pod, 1, W, H)
    for i in range(2):
        r.render_next_iteration(i, i, 0.3, det)
    img = r.output()
    assert np.isfinite(img).all() and (img >= 0).all() and img.sum() > 0


@pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_matches_host_compiled_reference_at_baseline_config_2(oracle):
    """BASELINE config 2 as stated — CornellSmall, BDPT, max path length 8, 512 x 512 — rendered by the oracle and by the reference's
    own programs (the latter on one thread: its light-tracing splat is a plain `+=`, vcm.h:159): ray counts identical, image within
    1e-5.  tests/test_parity_rule_gpu.py::test_baseline_config_2_as_stated compares the CUDA path with this oracle at this size."""
    a = render("oracle", "CornellSmall", 0, 1, 512, 16, 2, 8, None)
    b = render("ref", "CornellSmall", 0, 1, 512, 16, 2, 8, None)
    assert b["output"].sum() > 0
    _same(a, b)


@pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("kind,config_no", [("sponza", 4), ("conference", 3)])
def test_oracle_matches_host_compiled_reference_at_baseline_configs_3_and_4(oracle, kind, config_no):
    """BASELINE configs 3 and 4 AS STATED — the synthetic Conference-scale (~570 K triangles, glass / mirror) and Sponza-scale (~1.1 M
    triangles, textured) scenes, PPM, 1920 x 1080, 1024 x 1024 photons, one iteration — rendered by the oracle and by the reference's
    own programs on all host cores: closest-hit and any-hit ray counts, all 4 M photon cell keys, the offset table and all 2 M hit
    flags identical, image within 1e-5.  tests/test_parity_rule_gpu.py::test_baseline_bench_configs_as_stated compares the CUDA path
    with this oracle on exactly these inputs."""
    from oppositerenderer_b200 import synthetic
    scene = synthetic.make_scene(kind, seed=1, detail=1.0)
    W, H, pl = 1920, 1080, (1024, 1024)
    cam = scene.getDefaultCamera()
    cam.setAspectRatio(W / H)
    out = {}
    for which in ("oracle", "ref"):
        r = ol.OracleRenderer(which, pl, 10, 1645301512, threads=len(os.sched_getaffinity(0)))
        r.init_scene(scene.getSceneDescription())
        r.render_next_iteration(0, 0, scene.getSceneInitialPPMRadiusEstimate(), ol.make_details(cam.pod, 2, W, H))
        out[which] = dict(output=r.output(), rays=np.array(r.ray_counts(), np.uint64), photon_keys=r.photon_keys(), offsets=r.offsets(),
                          hitpoint_flags=r.hitpoints()["flags"])
        del r
    assert out["ref"]["rays"][0] > W * H + pl[0] * pl[1] and out["ref"]["output"].sum() > 0
    _same(out["oracle"], out["ref"])


@pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_matches_host_compiled_reference_on_baseline_config_5_workload(oracle):
    """BASELINE config 5's workload — BDPT (vertex connection, recursive MIS, max path length 8) on the ~1.1 M-triangle textured scene —
    at 960 x 540, the size of tests/test_parity_rule_gpu.py::test_baseline_config_5_workload: one iteration by the oracle (all cores) and by
    the reference's own programs (one thread: its splat is a plain `+=`): ray counts identical, image within 1e-5."""
    from oppositerenderer_b200 import synthetic
    scene = synthetic.make_scene("sponza", seed=1, detail=1.0)
    W, H, pl = 960, 540, (16, 16)
    cam = scene.getDefaultCamera()
    cam.setAspectRatio(W / H)
    radius = float(ol.ppm_radius_schedule(scene.getSceneInitialPPMRadiusEstimate(), 1)[0])
    out = {}
    for which in ("oracle", "ref"):
        r = ol.OracleRenderer(which, pl, 8, 1645301512, threads=1 if which == "ref" else len(os.sched_getaffinity(0)))
        r.init_scene(scene.getSceneDescription())
        r.render_next_iteration(0, 0, radius, ol.make_details(cam.pod, 1, W, H))
        out[which] = dict(output=r.output(), rays=np.array(r.ray_counts(), np.uint64))
        del r
    assert out["ref"]["rays"][0] > 2 * W * H and out["ref"]["output"].sum() > 0
    _same(out["oracle"], out["ref"])


_CS = [0, (1 << 1) | (1 << 7) | (1 << 11) | (1 << 9), (1 << 3) | (1 << 8) | (1 << 12) | (1 << 13), (1 << 5) | (1 << 10)]   # = test_parity_rule_gpu.CORNELL_SMALL
_RULE_INPUTS = ([("Cornell", 0, 2, 64, 64, 3, 10, None)] + [("CornellSmall", c, 2, 64, 64, 3, 10, 0.05) for c in _CS] +
                [("CornellSmall", c, 1, 48, 16, 3, m, None) for c, m in zip(_CS, (8, 8, 10, 10))] +
                [("Cornell", 0, 0, 48, 16, 3, 10, None), ("CornellSmall", _CS[1], 0, 48, 16, 3, 10, None), ("CornellSmall", _CS[2], 0, 48, 16, 3, 10, None)])


@pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("case", _RULE_INPUTS, ids=["%s-%d-m%d" % c[:3] for c in _RULE_INPUTS])
def test_oracle_matches_host_compiled_reference_on_the_inputs_of_the_gpu_rule_tests(oracle, case):
    """Every built-in-scene case of tests/test_parity_rule_gpu.py (test_ppm_… / test_bdpt_… / test_path_tracing_every_pixel_in_tolerance_
    or_explained: same scene variant, size, photon launch, iterations, path length and radius), oracle against the reference's own
    programs: integer results identical, images within 1e-5."""
    _same(render("oracle", *case), render("ref", *case))


@pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("case", CASES + SYNTH_CASES, ids=[c[0] for c in CASES + SYNTH_CASES])
def test_golden_fixtures_are_what_the_reference_programs_produce(case):
    """The committed fixtures are not stale: the reference's own programs (oracle/_ref, built from /root/reference by oracle/refshim)
    reproduce every one of them now — integer results identical, images within 1e-5 (the light-tracing splat order)."""
    gold = dict(np.load(os.path.join(GOLDEN, case[0] + ".npz")))
    got = render("ref", *case[1:]) if case in CASES else render_synthetic("ref", *case[1:])
    _same(got, gold)
