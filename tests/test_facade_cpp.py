"""The header-only C++ façade (include/OptixRenderer.h: the reference's class / method / enum names over the C ABI),
driven the way Standalone/StandaloneRenderManager.cpp drives the reference's OptixRenderer."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "oppositerenderer_b200", "csrc")


@pytest.fixture(scope="module")
def facade_exe(product, tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("facade") / "facade_check")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "facade_check.cpp"),
                           "-L", LIBDIR, "-loppositerenderer_b200", "-Wl,-rpath," + LIBDIR, "-o", exe])
    return exe


def test_facade_state_rules_and_pods(facade_exe):
    out = subprocess.run([facade_exe, "rules"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "ok initScene before initialize" in out.stdout and "ok render before initialize" in out.stdout


@pytest.mark.gpu
def test_facade_render_equals_python_mirror(facade_exe, product, tmp_path):
    """Same C ABI underneath: the C++ driver loop (double-precision radius recurrence) and the Python mirror (closed form)
    must produce the same sum buffer."""
    W, H, N = 40, 32, 3
    raw = str(tmp_path / "out.f32")
    out = subprocess.run([facade_exe, "render", str(W), str(H), str(N), raw], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    a = np.fromfile(raw, np.float32).reshape(H, W, 3)
    scene = product.Cornell()
    cam = scene.getDefaultCamera()
    cam.setAspectRatio(W / H)
    r = product.OptixRenderer(photon_launch=(64, 64), seed=1645301512)
    r.initialize(product.ComputeDevice(0))
    r.initScene(scene)
    det = product.RenderServerRenderRequestDetails(cam, scene.getSceneName(), product.RenderMethod.PROGRESSIVE_PHOTON_MAPPING, W, H, 2.0 / 3.0)
    radii = product.ppm_radius_schedule(scene.getSceneInitialPPMRadiusEstimate(), N)
    for i in range(N):
        r.renderNextIteration(i, i, radii[i], True, det)
    b = r.getOutputBuffer()
    assert a.sum() > 0
    # photon stores race only in order, never in content; sums are order-independent per pixel here
    assert np.allclose(a, b, rtol=1e-5, atol=1e-6 * float(b.max()))


@pytest.mark.gpu
def test_facade_group_of_one_equals_single_renderer(facade_exe, tmp_path):
    """OptixRendererGroup (opr_group_*: hand-out loop, worker thread per GPU, reduce) over ONE GPU must reproduce the single
    renderer bit for bit (rank 0 keeps the base seed; the radii of the closed form round to the same floats as the recurrence)."""
    W, H, N = 40, 32, 3
    one, grp = str(tmp_path / "one.f32"), str(tmp_path / "grp.f32")
    for args in (["render", str(W), str(H), str(N), one], ["group", str(W), str(H), str(N), grp, "1"]):
        out = subprocess.run([facade_exe] + args, capture_output=True, text=True)
        assert out.returncode == 0, (args, out.returncode, out.stdout + out.stderr)
    a, b = np.fromfile(one, np.float32), np.fromfile(grp, np.float32)
    assert a.sum() > 0 and np.allclose(a, b, rtol=1e-5, atol=1e-6 * float(a.max()))


@pytest.mark.gpu
def test_facade_group_of_two_gpus(facade_exe, product, tmp_path):
    """Two GPUs from one C++ process: 4 global iterations handed out over 2 GPUs and reduced with NCCL from C++ equal the sum of two
    Python-mirror renderers given the per-rank seeds and the partition's iterations."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from oppositerenderer_b200 import distributed as D
    W, H, N = 40, 32, 4
    raw = str(tmp_path / "grp2.f32")
    out = subprocess.run([facade_exe, "group", str(W), str(H), str(N), raw, "2"], capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stdout + out.stderr)
    a = np.fromfile(raw, np.float32).reshape(H, W, 3)
    scene = product.Cornell()
    cam = scene.getDefaultCamera()
    cam.setAspectRatio(W / H)
    det = product.RenderServerRenderRequestDetails(cam, scene.getSceneName(), product.RenderMethod.PROGRESSIVE_PHOTON_MAPPING, W, H, 2.0 / 3.0)
    expect = np.zeros((H, W, 3), np.float32)
    for rank in range(2):
        r = product.OptixRenderer(photon_launch=(64, 64), seed=D.rank_seed(1645301512, rank))
        r.initialize(product.ComputeDevice(rank))
        r.initScene(scene)
        D.render_partition(r, D.IterationPartition(rank, 2, scene.getSceneInitialPPMRadiusEstimate()), det, 0, N // 2)
        expect = expect + r.getOutputBuffer()
    assert np.allclose(a, expect, rtol=1e-5, atol=1e-6 * float(expect.max()))
