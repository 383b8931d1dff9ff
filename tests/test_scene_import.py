"""Collada import: Scene::createFromFile restated for .dae files (host_scene_import.cpp) — node
transforms, triangulation, generated normals / tangents, the reference's material classification, lights, emitter quads, AABB
and camera rules (RenderEngine/scene/Scene.cpp:73-311, 469-485), then the imported scene through the renderer."""
import ctypes as C
import os

import numpy as np
import pytest

import collada_fixture as cf
import oracle_lib as ol
import util
from oppositerenderer_b200 import capi


def _mesh(d, i):
    m = d.meshes[i]
    v = np.ctypeslib.as_array(m.vertices, (m.numVertices, 3)).copy()
    f = np.ctypeslib.as_array(m.indices, (m.numTriangles, 3)).copy()
    n = np.ctypeslib.as_array(m.normals, (m.numVertices, 3)).copy() if m.normals else None
    uv = np.ctypeslib.as_array(m.uvs, (m.numVertices, 2)).copy() if m.uvs else None
    t = np.ctypeslib.as_array(m.tangents, (m.numVertices, 3)).copy() if m.tangents else None
    return v, f, n, uv, t, m.material


def test_collada_room_import(product, tmp_path):
    path, tex = cf.write_room(str(tmp_path))
    scene = product.Scene.createFromFile(path)
    d = scene.getSceneDescription().contents
    assert scene.getSceneName() == os.path.realpath(path)
    assert capi.load_library().opr_scene_desc_check(scene.getSceneDescription(), None, 0) == 0     # what initScene checks first
    # ---- material classification, Scene.cpp:178-265, in <library_materials> order
    kinds = [d.materials[i].kind for i in range(d.numMaterials)]
    assert kinds == [capi.MAT_DIFFUSE, capi.MAT_DIFFUSE, capi.MAT_DIFFUSE_EMITTER, capi.MAT_GLASS, capi.MAT_MIRROR, capi.MAT_TEXTURE]
    assert list(d.materials[1].Kd) == pytest.approx([0.8, 0.1, 0.1]) and d.materials[3].indexOfRefraction == pytest.approx(1.5)
    assert list(d.materials[3].Kr) == [1, 1, 1] and list(d.materials[4].Kr) == pytest.approx([0.9, 0.9, 0.9])
    assert d.materials[5].diffuseTexture == 0 and d.materials[5].normalTexture == 1 and d.numTextures == 2
    t0 = np.ctypeslib.as_array(d.textures[0].rgba, (8, 8, 4))
    assert (t0[..., 0] == tex[..., 2]).all() and (t0[..., 2] == tex[..., 0]).all() and (t0[..., 3] == 1).all()      # Image.cpp: BGR->RGB, alpha 1
    # ---- geometry: 9 instances; quads triangulated to 2, the box to 12, the pentagon to 3 triangles
    assert d.numMeshes == 9 and scene.getNumMeshes() == 9 and scene.getNumTriangles() == 7 * 2 + 12 + 3
    v, f, n, uv, t, mat = _mesh(d, 0)            # floor: scale (2, 1, 2)
    assert mat == 5 and sorted(map(tuple, v.tolist())) == sorted([(-2, 0, -2), (2, 0, -2), (2, 0, 2), (-2, 0, 2)])
    assert np.allclose(n, [0, 1, 0]) and uv is not None and t is not None and np.allclose(np.abs((t * n).sum(1)), 0, atol=1e-6)
    fn = np.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]])
    assert (fn[:, 1] > 0).all()                   # winding preserved: faces look up into the room
    v, f, n, uv, t, mat = _mesh(d, 1)            # ceiling: translate(0,4,0) * rotate(x,180) * scale
    assert np.allclose(v[:, 1], 4) and np.allclose(n, [0, -1, 0], atol=1e-6)
    v, f, n, uv, t, mat = _mesh(d, 3)            # left wall through an explicit <matrix>
    assert mat == 1 and np.allclose(v[:, 0], -2) and np.allclose(n, [1, 0, 0], atol=1e-6) and np.allclose(sorted(v[:, 1]), [0, 0, 4, 4])
    v, f, n, uv, t, mat = _mesh(d, 6)            # glass box: no normals in the file -> GenSmoothNormals (per position, area weighted)
    assert mat == 3 and v.shape == (8, 3) and f.shape == (12, 3) and uv is None
    c = np.array([-0.8, 0.61, 0.3])
    assert np.allclose(np.linalg.norm(n, axis=1), 1, atol=1e-5) and ((n * (v - c)).sum(1) > 0).all()
    v, f, n, uv, t, mat = _mesh(d, 8)            # pentagon under two nested nodes, fan-triangulated, generated normals
    assert f.shape == (3, 3) and np.allclose(v[:, 1], 0.3) and np.allclose(np.abs(n), [0, 1, 0], atol=1e-6)
    assert np.allclose(v[0], [0.8 + 0.5, 0.3, -0.5], atol=1e-6)
    # ---- lights: the point light, then the emitter quad twice (Scene.cpp:141 and :508)
    L = scene.getSceneLights()
    assert [l.lightType for l in L] == [capi.LIGHT_POINT, capi.LIGHT_AREA, capi.LIGHT_AREA]
    assert list(L[0].position) == pytest.approx([1.2, 3.0, -1.0]) and list(L[0].power) == pytest.approx([30, 30, 28])
    assert bytes(memoryview(L[1]).cast("B")) == bytes(memoryview(L[2]).cast("B"))
    lv, lf, _, _, _, lmat = _mesh(d, 5)
    assert lmat == 2 and list(L[1].position) == pytest.approx(lv[lf[0, 0]].tolist())
    assert list(L[1].v1) == pytest.approx((lv[lf[0, 1]] - lv[lf[0, 0]]).tolist()) and list(L[1].v2) == pytest.approx((lv[lf[0, 2]] - lv[lf[0, 0]]).tolist())
    assert L[1].area_or_angle == pytest.approx(1.0) and d.materials[2].inverseArea == pytest.approx(1.0) and list(L[1].normal) == pytest.approx([0, -1, 0], abs=1e-6)
    once = product.Scene.createFromFile(path, single_emitter_lights=True)
    assert len(once.getSceneLights()) == 2
    # ---- AABB from the FIRST vertex of every face (Scene.cpp:151-153), not the true bounds
    firsts = np.concatenate([_mesh(d, i)[0][_mesh(d, i)[1][:, 0]] for i in range(d.numMeshes)])
    mn, mx = scene.getSceneAABB()
    assert np.allclose(mn, firsts.min(0)) and np.allclose(mx, firsts.max(0))
    # ---- camera: eye, lookAt = eye + dir, fov = xfov_rad * 365 / (2 pi) (Scene.cpp:469-485), KeepHorizontal
    cam = scene.getDefaultCamera().pod
    assert list(cam.eye) == pytest.approx([0, 2, -5.5]) and list(cam.lookat) == pytest.approx([0, 2, -4.5], abs=1e-6)
    assert cam.hfov == pytest.approx(50.0 * 365.0 / 360.0, rel=1e-6) and cam.aspectRatioMode == 1
    assert scene.getSceneInitialPPMRadiusEstimate() > 0 and scene.warnings() == ""


def test_collada_up_axis_and_errors(product, tmp_path):
    path, _ = cf.write_room(str(tmp_path / "z"), up_axis="Z_UP")
    zscene = product.Scene.createFromFile(path)          # the handle owns the arrays: keep it alive
    d = zscene.getSceneDescription().contents
    v = _mesh(d, 1)[0]                           # ceiling at y = 4 in file space -> (x, z, -y)
    assert np.allclose(v[:, 2], -4, atol=1e-6)
    with pytest.raises(RuntimeError, match=r"The file that was supplied \(.*\) does not exist"):
        product.Scene.createFromFile(str(tmp_path / "nope.dae"))
    bad = tmp_path / "bad.dae"
    bad.write_text("<COLLADA><library_geometries></COLLADA>")
    with pytest.raises(RuntimeError, match="An error occurred in Assimp during reading of this file"):
        product.Scene.createFromFile(str(bad))
    notex = tmp_path / "z" / "floor.tga"
    os.remove(notex)
    with pytest.raises(RuntimeError, match="does not exist"):
        product.Scene.createFromFile(path)


def test_malformed_collada_never_reads_out_of_bounds(product, tmp_path):
    """Regression for two fuzz findings (AddressSanitizer on host_scene_import.cpp): a negative <input offset> indexed in front of
    the index tuple, and a negative <vcount> entry wrapped the bounds check of the polygon split.  Every mutation must either load
    (with the damaged primitive dropped or clamped) or fail with a message — and the arrays it returns must be self-consistent."""
    import random
    import re
    path, _ = cf.write_room(str(tmp_path / "m"))
    src = open(path).read()
    rng = random.Random(7)
    cases = [re.sub(r'offset="1"', 'offset="-1"', src, count=1), re.sub(r'offset="0"', 'offset="-7"', src, count=2),
             re.sub(r'offset="1"', 'offset="100000"', src, count=1),
             re.sub(r"<vcount>(\d+) ", lambda m: "<vcount>-3 ", src, count=1), re.sub(r"<vcount>(\d+) ", lambda m: "<vcount>4611686018427387904 ", src, count=1)]
    for _ in range(40):                                  # numbers replaced by hostile values anywhere in the file
        t = src
        for _ in range(rng.randrange(1, 4)):
            ms = list(re.finditer(r"-?\d+", t))
            m = rng.choice(ms)
            t = t[:m.start()] + str(rng.choice([-1, 99999999, 0, 2 ** 31, 7, -5, 1000])) + t[m.end():]
        cases.append(t)
    loaded = 0
    for i, text in enumerate(cases):
        f = tmp_path / "m" / ("case%d.dae" % i)
        f.write_text(text)
        try:
            scene = product.Scene.createFromFile(str(f))
        except RuntimeError as e:
            assert str(e)
            continue
        loaded += 1
        d = scene.getSceneDescription().contents
        for k in range(d.numMeshes):
            if d.meshes[k].numTriangles == 0:
                continue
            v, f = _mesh(d, k)[:2]
            assert f.min() >= 0 and f.max() < len(v)
            assert np.isfinite(v).all()
    assert loaded >= 10


def test_oracle_renders_the_imported_scene(product, oracle, tmp_path):
    """CPU: the imported description is a valid scene for the oracle (all six material kinds but glossy, both light types)."""
    path, _ = cf.write_room(str(tmp_path))
    scene = product.Scene.createFromFile(path)
    o = oracle.OracleRenderer("oracle", (32, 32), 10, util.SEED)
    o.init_scene(scene.getSceneDescription())
    cam = scene.getDefaultCamera()
    cam.setAspectRatio(1.0)
    o.render_next_iteration(0, 0, 0.3, oracle.make_details(cam.pod, capi.PROGRESSIVE_PHOTON_MAPPING, 24, 24))
    out = o.output()
    assert np.isfinite(out).all() and out.mean() > 0
    flags = o.hitpoints()["flags"]
    assert ((flags & ol.PRD_HIT_NON_SPECULAR) != 0).mean() > 0.5 and ((flags & ol.PRD_MISS) != 0).mean() < 0.2


@pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("method", [2, 1, 0])
def test_oracle_matches_host_compiled_reference_on_the_imported_scene(product, oracle, tmp_path, method):
    """CPU: the imported room (glass, mirror, textured + normal-mapped, diffuse and emitter meshes; area + point light) rendered by the
    oracle and by the reference's own programs (oracle/_ref) with the settings of test_imported_scene_parity: ray counts, photon keys,
    offsets and hit flags identical, images within 1e-5.  This pins the oracle on the inputs of that GPU test."""
    path, _ = cf.write_room(str(tmp_path))
    scene = product.Scene.createFromFile(path)
    cam = scene.getDefaultCamera()
    cam.setAspectRatio(1.0)
    size, pl, iters = 48, (64, 64), 2
    out = {}
    for which in ("oracle", "ref"):
        r = oracle.OracleRenderer(which, pl, 8, util.SEED, threads=1 if (which == "ref" and method == 1) else 0)
        r.init_scene(scene.getSceneDescription())
        det = oracle.make_details(cam.pod, method, size, size)
        for i in range(iters):
            r.render_next_iteration(i, i, 0.25, det)
        out[which] = dict(output=r.output(), rays=np.array(r.ray_counts(), np.uint64))
        if method == 2:
            out[which].update(photon_keys=r.photon_keys(), offsets=r.offsets(), hitpoint_flags=r.hitpoints()["flags"])
    a_, b_ = out["oracle"], out["ref"]
    assert b_["output"].sum() > 0
    assert (a_["rays"] == b_["rays"]).all(), (a_["rays"], b_["rays"])
    assert np.allclose(a_["output"], b_["output"], rtol=1e-5, atol=2e-5 * float(b_["output"].max()))
    if method == 2:
        for k in ("photon_keys", "offsets", "hitpoint_flags"):
            assert (a_[k] == b_[k]).all(), k


@pytest.mark.gpu
@pytest.mark.parametrize("method", [2, 1, 0])
def test_imported_scene_parity(product, oracle, tmp_path, method):
    """GPU: PPM, BDPT and PT of the imported room against the oracle on identical paths, under the per-pixel rule of
    tests/parity_rules.py (every pixel in tolerance or accounted for by a measured divergence)."""
    import parity_rules as pr
    path, _ = cf.write_room(str(tmp_path))
    scene = product.Scene.createFromFile(path)
    cam = scene.getDefaultCamera()
    cam.setAspectRatio(1.0)
    size, pl, iters = 48, (64, 64), 2
    r = product.OptixRenderer(photon_launch=pl, vcm_max_path_length=8, seed=util.SEED)
    r.initialize(product.ComputeDevice(0))
    r.initScene(scene)
    o = oracle.OracleRenderer("oracle", pl, 8, util.SEED)
    o.init_scene(scene.getSceneDescription())
    det = product.RenderServerRenderRequestDetails(cam, scene.getSceneName(), method, size, size, 2.0 / 3.0)
    odet = oracle.make_details(cam.pod, method, size, size)
    res = pr.run_lockstep("imported room, method %d" % method, r, o, det, odet, method, size, size, pl, [0.25] * iters, cam.pod)
    assert res["rel_l1"] < (6e-2 if method == 0 else 3e-2), res
