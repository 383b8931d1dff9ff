"""Host-side I/O either side of the path: display normalisation + BMP export
(Gui/gui/RenderWidget.cpp:187-199, 243-317), texture image loading (RenderEngine/util/Image.cpp:17-140) and the
Standalone driver loop (Standalone/StandaloneRenderManager.cpp:75-140) through the opr_render tool."""
import ctypes as C
import os
import struct
import subprocess

import numpy as np
import pytest

from oppositerenderer_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PF, PB = C.POINTER(C.c_float), C.POINTER(C.c_uint8)


def _display_reference(sum_rgb, iterations, gamma):
    """numpy restatement of RenderWidget.cpp:297-315: bgr = clamp(pow(rgb * invIterations, invGamma) * 255) truncated to a byte."""
    inv_gamma, inv_it = np.float32(1.0) / np.float32(gamma), np.float32(1.0) / np.float32(iterations)
    with np.errstate(invalid="ignore"):
        v = np.power(sum_rgb[..., ::-1].astype(np.float32) * inv_it, inv_gamma) * np.float32(255.0)
    v = np.where(np.isnan(v), 0.0, v)           # std::max(0.f, NaN) == 0.f
    return np.minimum(255.0, np.maximum(0.0, v)).astype(np.uint8)


def test_display_and_bmp_match_the_reference_formula(product, tmp_path):
    lib = capi.load_library()
    rng = np.random.default_rng(3)
    w, h, its = 37, 21, 12.0            # width*3 not a multiple of 4: the reference writes no row padding, neither do we
    img = (rng.random((h, w, 3), dtype=np.float32) ** 3 * 30.0).astype(np.float32)
    img[0, 0] = [0.0, 1e9, -1.0]
    out = np.zeros((h, w, 3), np.uint8)
    assert lib.opr_display_bgr8(img.ctypes.data_as(PF), w, h, its, 2.2, out.ctypes.data_as(PB)) == 0
    ref = _display_reference(img, its, 2.2)
    assert np.abs(out.astype(int) - ref.astype(int)).max() <= 1 and (out == ref).mean() > 0.999
    assert tuple(out[0, 0]) == (0, 255, 0)
    path = str(tmp_path / "a.bmp")
    assert lib.opr_save_bmp(path.encode(), img.ctypes.data_as(PF), w, h, its, 2.2) == 0
    raw = open(path, "rb").read()
    assert raw[:2] == b"BM" and len(raw) == 54 + w * h * 3
    size, res, off, hsz, bw, bh, planes, bpp, comp, isz, hres, vres, pal, imp = struct.unpack("<IIIIiihhIIIIII", raw[2:54])
    assert (size, res, off, hsz, bw, bh, planes, bpp, comp, isz, hres, vres, pal, imp) == (54 + w * h * 3, 0, 54, 40, w, h, 1, 24, 0, w * h * 3, 2953, 2953, 0, 0)
    assert (np.frombuffer(raw[54:], np.uint8).reshape(h, w, 3) == out).all()
    assert lib.opr_display_bgr8(img.ctypes.data_as(PF), w, h, 0.0, 2.2, out.ctypes.data_as(PB)) != 0


def _write_tga(path, pixels_bgr, img_type=2, origin_top=False, id_len=3):
    h, w, d = pixels_bgr.shape
    hdr = bytes([id_len, 0, img_type]) + bytes(5) + struct.pack("<HHHH", 0, 0, w, h) + bytes([8 * d, 0x20 if origin_top else 0])
    open(path, "wb").write(hdr + bytes(id_len) + pixels_bgr.tobytes())


def _load(lib, path):
    w, h = C.c_int32(), C.c_int32()
    rc = lib.opr_image_info(path.encode(), C.byref(w), C.byref(h))
    if rc != 0:
        raise RuntimeError(lib.opr_io_last_error().decode())
    px = np.zeros((h.value, w.value, 4), np.uint8)
    assert lib.opr_image_read(path.encode(), px.ctypes.data_as(PB), px.nbytes) == 0
    return px


def test_tga_loading_follows_image_cpp(product, tmp_path):
    """Image.cpp:88-112 — types 2/3 uncompressed, BGR->RGB (libtga TGA_RGB), rows kept in file order, alpha byte = 1."""
    lib = capi.load_library()
    rng = np.random.default_rng(5)
    for depth in (3, 4):
        bgr = rng.integers(0, 256, (5, 7, depth), dtype=np.uint8)
        p = str(tmp_path / ("t%d.tga" % depth))
        _write_tga(p, bgr)
        px = _load(lib, p)
        assert px.shape == (5, 7, 4)
        assert (px[..., 0] == bgr[..., 2]).all() and (px[..., 1] == bgr[..., 1]).all() and (px[..., 2] == bgr[..., 0]).all()
        assert (px[..., 3] == 1).all()
    p = str(tmp_path / "rle.tga")
    _write_tga(p, rng.integers(0, 256, (2, 2, 3), dtype=np.uint8), img_type=10)
    with pytest.raises(RuntimeError, match="does not support type '10'"):
        _load(lib, p)
    with pytest.raises(RuntimeError, match="does not exist"):
        _load(lib, str(tmp_path / "missing.tga"))
    small = np.zeros((2, 2, 4), np.uint8)
    p = str(tmp_path / "t3.tga")
    assert lib.opr_image_read(p.encode(), small.ctypes.data_as(PB), small.nbytes) != 0      # 7x5 image, 2x2 destination


def test_bmp_round_trip_through_the_writer(product, tmp_path):
    lib = capi.load_library()
    rng = np.random.default_rng(7)
    w, h = 8, 6                          # width*3 % 4 == 0: a well-formed file
    img = rng.random((h, w, 3), dtype=np.float32)
    path = str(tmp_path / "r.bmp")
    assert lib.opr_save_bmp(path.encode(), img.ctypes.data_as(PF), w, h, 1.0, 1.0) == 0
    px = _load(lib, path)
    expect = _display_reference(img, 1.0, 1.0)[::-1, :, ::-1]     # the loader presents rows top-down and RGB
    assert (px[..., :3] == expect).all() and (px[..., 3] == 255).all()


def test_texture_file_feeds_a_texture_material(product, tmp_path):
    """A TGA on disk -> RGBA8 array -> opr_texture of a SceneDescription (what material/Texture.cpp:109-131 does with Image)."""
    import oppositerenderer_b200 as opr
    from oppositerenderer_b200.renderer import make_material, SceneDescription
    lib = capi.load_library()
    bgr = np.full((4, 4, 3), (10, 128, 250), np.uint8)
    p = str(tmp_path / "albedo.tga")
    _write_tga(p, bgr)
    tex = _load(lib, p)
    quad = dict(vertices=np.array([[0, 0, 0], [1, 0, 0], [1, 0, 1], [0, 0, 1]], np.float32), normals=np.tile([0, 1, 0], (4, 1)).astype(np.float32),
                uvs=np.array([[0, 0], [1, 0], [1, 1], [0, 1]], np.float32), tangents=None, bitangents=None, indices=np.array([[0, 2, 1], [0, 3, 2]], np.int32), material=0)
    light = opr.Light((1, 1, 1), (0.4, 1, 0.4), (0.2, 0, 0), (0, 0, 0.2))
    s = SceneDescription("tex", (0, 0, 0), (1, 1, 1), [make_material(capi.MAT_TEXTURE, diffuseTexture=0)], [light], meshes=[quad], textures=[tex])
    d = s.getSceneDescription().contents
    assert d.numTextures == 1 and d.textures[0].width == 4 and d.textures[0].rgba[0] == 250 and d.textures[0].rgba[2] == 10 and d.textures[0].rgba[3] == 1


@pytest.mark.gpu
def test_opr_render_tool_runs_the_driver_loop(product, tmp_path):
    """tools/opr_render.cpp = StandaloneRenderManager loop in C++ (double radius recurrence, read-back every 5th iteration, BMP
    export); its sum buffer equals the Python mirror's and its BMP equals the display formula applied to that buffer."""
    exe = os.path.join(ROOT, "oppositerenderer_b200", "csrc", "opr_render")
    bmp, raw = str(tmp_path / "o.bmp"), str(tmp_path / "o.f32")
    out = subprocess.run([exe, "--scene", "CornellSmall", "--method", "bdpt", "--size", "48x40", "--iterations", "7", "--out", bmp, "--raw", raw], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    a = np.fromfile(raw, np.float32).reshape(40, 48, 3)
    scene = product.CornellSmall()
    cam = scene.getDefaultCamera()
    cam.setAspectRatio(48 / 40)
    r = product.OptixRenderer()
    r.initialize(product.ComputeDevice(0))
    r.initScene(scene)
    det = product.RenderServerRenderRequestDetails(cam, scene.getSceneName(), product.RenderMethod.VCM_BIDIRECTIONAL_PATH_TRACING, 48, 40, 2.0 / 3.0)
    radii = product.ppm_radius_schedule(scene.getSceneInitialPPMRadiusEstimate(), 7)
    for i in range(7):
        r.renderNextIteration(i, i, radii[i], True, det)
    b = r.getOutputBuffer()
    assert np.allclose(a, b, rtol=1e-4, atol=1e-5 * float(b.max()))      # the light-tracing splat order is not fixed (atomics)
    data = np.frombuffer(open(bmp, "rb").read()[54:], np.uint8).reshape(40, 48, 3)
    ref = _display_reference(a, 7.0, 2.2)
    assert np.abs(data.astype(int) - ref.astype(int)).max() <= 1


@pytest.mark.gpu
def test_opr_render_tool_distributed_over_one_gpu_equals_standalone(product, tmp_path):
    """--gpus 1 runs the Client/Server arrangement (DistributedRenderManager -> opr_group_*: hand-out loop, worker thread, reduce) on one
    GPU: same iterations, same radii (closed form == recurrence in float), same seed -> the same sum buffer as the standalone loop."""
    exe = os.path.join(ROOT, "oppositerenderer_b200", "csrc", "opr_render")
    raws = []
    for extra in ([], ["--gpus", "1"]):
        raw = str(tmp_path / ("o%d.f32" % len(raws)))
        out = subprocess.run([exe, "--scene", "Cornell", "--method", "ppm", "--size", "48x40", "--iterations", "6", "--out", str(tmp_path / "o.bmp"), "--raw", raw] + extra,
                             capture_output=True, text=True)
        assert out.returncode == 0, out.stdout + out.stderr
        raws.append(np.fromfile(raw, np.float32))
    assert raws[0].sum() > 0 and np.allclose(raws[0], raws[1], rtol=1e-5, atol=1e-6 * float(raws[0].max()))
