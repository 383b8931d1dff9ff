"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU, exports every symbol the headers
declare, fails loudly on compute calls without a device, and its host-side helpers (Camera, Light, built-in
scenes, radius schedule) agree with the oracle's independent restatement of the same reference code."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oppositerenderer_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "opr.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(opr_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(product):
    lib = capi.load_library()
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), "include/opr.h declares %s but the library does not export it" % s


def test_no_cpu_fallback(product):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = product.OptixRenderer()
    with pytest.raises(RuntimeError, match="no CPU path"):
        r.initialize(product.ComputeDevice(0))
    lib = capi.load_library()
    st = (C.c_uint32 * 6)()
    assert lib.opr_rng_init(st, 1, 0, 1) != 0
    assert b"no CPU path" in lib.opr_last_error()


def test_state_rules(product):
    # OptixRenderer.cpp:438-441 / :513-516: initScene and renderNextIteration require initialize()
    r = product.OptixRenderer()
    scene = product.Cornell()
    with pytest.raises(RuntimeError, match="Cannot initialize scene before OptixRenderer"):
        r.initScene(scene)
    det = product.RenderServerRenderRequestDetails(scene.getDefaultCamera(), "Cornell", product.RenderMethod.PROGRESSIVE_PHOTON_MAPPING, 8, 8, 2 / 3)
    with pytest.raises(RuntimeError, match="Traced before OptixRenderer was initialized"):
        r.renderNextIteration(0, 0, 1.0, True, det)


def _bytes(x):
    return bytes(memoryview(x).cast("B"))


@pytest.mark.parametrize("name,config", [("Cornell", 0), ("CornellSmall", 0), ("CornellSmall", (1 << 0) | (1 << 1) | (1 << 9)),
                                         ("CornellSmall", (1 << 3) | (1 << 7) | (1 << 11)), ("CornellSmall", (1 << 5) | (1 << 8) | (1 << 12) | (1 << 13)),
                                         ("CornellSmall", (1 << 2) | (1 << 6) | (1 << 10))])
def test_builtin_scene_tables_match_oracle(product, oracle, name, config):
    lib = capi.load_library()
    h = lib.opr_builtin_scene_create(name.encode(), config)
    a = lib.opr_builtin_scene_desc(h).contents
    osc = oracle.BuiltinScene(name, config)
    b = osc.desc.contents
    assert (a.numParallelograms, a.numSpheres, a.numMaterials, a.numLights) == (b.numParallelograms, b.numSpheres, b.numMaterials, b.numLights)
    assert list(a.aabbMin) == list(b.aabbMin) and list(a.aabbMax) == list(b.aabbMax)
    for i in range(a.numLights):
        assert _bytes(a.lights[i]) == _bytes(b.lights[i])

    def resolved(d, i):  # geometry with its material record (material ids may be numbered differently)
        p = d.parallelograms[i]
        return _bytes(p)[:36], _bytes(d.materials[p.material])

    assert [resolved(a, i) for i in range(a.numParallelograms)] == [resolved(b, i) for i in range(b.numParallelograms)]
    for i in range(a.numSpheres):
        assert _bytes(a.spheres[i])[:16] == _bytes(b.spheres[i])[:16]
        assert _bytes(a.materials[a.spheres[i].material]) == _bytes(b.materials[b.spheres[i].material])
    ca, cb = capi.opr_camera(), osc.camera()
    lib.opr_builtin_scene_camera(h, C.byref(ca))
    assert _bytes(ca) == _bytes(cb)
    lib.opr_builtin_scene_destroy(h)


def test_camera_aspect_ratio_and_radius(product, oracle):
    lib, olib = capi.load_library(), oracle.load("oracle")
    for ratio in (1.0, 16 / 9, 0.5):
        for mode in (0, 1, 2):
            a = product.Camera((1.25, 1.25, -2.85), (1.25, 1.25, 0), (0, 1, 0), 45, 45, 0.1, mode).pod
            b = capi.opr_camera.from_buffer_copy(_bytes(a))
            lib.opr_camera_set_aspect_ratio(C.byref(a), ratio)
            olib.ora_camera_set_aspect_ratio(C.byref(b), ratio)
            assert _bytes(a) == _bytes(b)
    mn, mx = (C.c_float * 3)(-5, -5, -5), (C.c_float * 3)(561, 553.85, 564.2)
    assert lib.opr_initial_ppm_radius(mn, mx) == olib.ora_initial_ppm_radius(mn, mx)
    r0 = 7.5
    sched = oracle.ppm_radius_schedule(r0, 40)
    for i in (0, 1, 7, 39):
        assert lib.opr_ppm_radius_for_iteration(r0, i, 2.0 / 3.0) == sched[i]


def test_light_constructors(product, oracle):
    olib = oracle.load("oracle")
    a = product.Light((0.5e6, 0.4e6, 0.2e6), (343.0, 548.7999, 227.0), (0, 0, 105.0), (-130.0, 0, 0)).pod
    b = capi.opr_light()
    f3 = lambda *v: (C.c_float * 3)(*v)
    olib.ora_make_area_light(C.byref(b), f3(0.5e6, 0.4e6, 0.2e6), f3(343.0, 548.7999, 227.0), f3(0, 0, 105.0), f3(-130.0, 0, 0))
    assert _bytes(a) == _bytes(b)
    assert list(a.normal) == [0.0, -1.0, 0.0] or np.allclose(list(a.normal), [0, -1, 0])
    a = product.Light((30, 30, 30), (1.25, 2.25, 1.25)).pod
    olib.ora_make_point_light(C.byref(b), f3(30, 30, 30), f3(1.25, 2.25, 1.25))
    assert _bytes(a) == _bytes(b)


def test_scene_description_check_rejects_what_the_device_would_follow(product):
    """opr_scene_desc_check (run by opr_init_scene before anything is uploaded): material kinds, material / texture / vertex indices
    and light types of a caller-built description — the reference trusts its own scene classes, a C ABI cannot."""
    from oppositerenderer_b200.renderer import make_material, SceneDescription
    lib = capi.load_library()
    msg = C.create_string_buffer(200)

    def check(scene):
        rc = lib.opr_scene_desc_check(scene.getSceneDescription(), msg, 200)
        return rc, msg.value.decode()

    verts = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    light = product.Light((1, 1, 1), (0, 5, 0))

    def scene(indices=(0, 1, 2), material=0, kind=capi.MAT_DIFFUSE, n_mats=1):
        mats = [make_material(kind, Kd=(0.5, 0.5, 0.5)) for _ in range(n_mats)]
        return SceneDescription("t", (0, 0, 0), (1, 1, 1), mats, [light], meshes=[dict(vertices=verts, indices=np.array([indices], np.int32), material=material)])

    assert check(scene())[0] == 0
    rc, text = check(scene(indices=(0, 1, 3)))
    assert rc == 1 and "vertex index 3 out of range" in text
    rc, text = check(scene(indices=(0, -1, 2)))
    assert rc == 1 and "out of range" in text
    rc, text = check(scene(material=1))
    assert rc == 1 and "material 1 does not exist" in text
    rc, text = check(scene(kind=17))
    assert rc == 1 and "unknown kind 17" in text
    rc, text = check(scene(kind=capi.MAT_TEXTURE))
    assert rc == 1 and "without a diffuse texture" in text
    assert lib.opr_scene_desc_check(None, msg, 200) == 1
    for name in ("Cornell", "CornellSmall"):                 # the built-in tables are sound
        h = lib.opr_builtin_scene_create(name.encode(), 0)
        assert lib.opr_scene_desc_check(lib.opr_builtin_scene_desc(h), msg, 200) == 0, msg.value
        lib.opr_builtin_scene_destroy(h)


@pytest.mark.parametrize("w,h", [(1920, 1080), (3840, 2160), (256, 256), (64, 64), (8, 4), (96, 54), (40, 32), (1, 1), (33, 7)])
def test_queue_slot_order_is_a_permutation_of_the_pixels(product, w, h):
    """Primary rays and direct-light samples take their pixels in 8 x 4 tiles (32 consecutive slots = one tile) when the image divides
    into tiles, else row by row; either way every pixel must be named by exactly one slot."""
    import numpy as np
    from oppositerenderer_b200 import capi
    lib = capi.load_library()
    n = w * h
    step = 1 if n <= 300000 else 0
    if step:
        px = np.array([lib.opr_debug_slot_pixel(s, w, h) for s in range(n)], np.int64)
        assert np.array_equal(np.sort(px), np.arange(n))
    else:       # large images: every slot of a sample of whole tile rows (a tile row is a closed set under the mapping) + closed form
        rows = [0, 1, (h // 4) // 2, h // 4 - 1]
        for tr in rows:
            s0 = tr * (w // 8) * 32
            px = np.array([lib.opr_debug_slot_pixel(s, w, h) for s in range(s0, s0 + (w // 8) * 32)], np.int64)
            assert np.array_equal(np.sort(px), np.arange(tr * 4 * w, (tr + 1) * 4 * w))
    if w % 8 == 0 and h % 4 == 0:
        first = [lib.opr_debug_slot_pixel(s, w, h) for s in range(min(32, n))]
        assert first == [(l >> 3) * w + (l & 7) for l in range(min(32, n))]          # slot l of tile 0 = pixel (l & 7, l >> 3)
    else:
        assert all(lib.opr_debug_slot_pixel(s, w, h) == s for s in range(0, n, max(1, n // 97)))
