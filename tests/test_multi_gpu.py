"""Two ranks on two B200s over NCCL (skipped with fewer than two GPUs): the iteration partition drives the CUDA renderer and the
running-sum buffers are reduced (i) by the C++ library itself — opr_comm_init_rank + opr_reduce_output, NCCL called from C++
behind the C ABI — and (ii) through torch.distributed on a zero-copy device view (the Python mirror); both totals on rank 0 equal
the host-side sum of the two ranks' buffers bit for bit (SURVEY.md §8e parity caveat (i)), and a second reduce gives the same
total (the running sums are never a reduce target).  A third test drives both GPUs from ONE process through opr_group_*."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SIZE, PL, LOCAL_ITERS, SEED = 64, (128, 128), 3, 1645301512


def _worker(rank, world, port, outdir):
    import torch
    import torch.distributed as dist
    import oppositerenderer_b200 as opr
    from oppositerenderer_b200 import capi, distributed as D
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    scene = opr.CornellSmall()
    cam = scene.getDefaultCamera()
    cam.setAspectRatio(1.0)
    r = opr.OptixRenderer(photon_launch=PL, seed=D.rank_seed(SEED, rank))
    r.initialize(opr.ComputeDevice(rank))
    r.initScene(scene)
    det = opr.RenderServerRenderRequestDetails(cam, scene.getSceneName(), opr.RenderMethod.PROGRESSIVE_PHOTON_MAPPING, SIZE, SIZE, 2.0 / 3.0)
    part = D.IterationPartition(rank, world, 0.05)
    D.render_partition(r, part, det, 0, LOCAL_ITERS)
    np.save(os.path.join(outdir, "rank%d.npy" % rank), r.getOutputBuffer())
    dev = torch.device("cuda", rank)
    # (i) the product's own collective: the 128-byte NCCL id travels over torch.distributed (plumbing), the reduce is C++
    ids = [opr.renderer.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    r.comm_init_rank(ids[0], rank, world)
    for k in range(2):
        total = np.zeros((SIZE, SIZE, 3), np.float32)
        r.reduce_output(0, total if rank == 0 else None)
        if rank == 0:
            np.save(os.path.join(outdir, "reduced_capi%d.npy" % k), total)
    r.synchronize()
    # (ii) the Python mirror over torch.distributed
    stream = torch.cuda.ExternalStream(capi.load_library().opr_get_stream(r._h), device=dev)
    with torch.cuda.stream(stream):
        total_t = D.reduce_sum_to_root(D.output_as_tensor(r, dev), dst=0)
    torch.cuda.synchronize()
    if rank == 0:
        np.save(os.path.join(outdir, "reduced.npy"), total_t.cpu().numpy())
    np.save(os.path.join(outdir, "after%d.npy" % rank), r.getOutputBuffer())   # the renderer's own sum is untouched by either reduce
    dist.barrier()
    r.comm_destroy()
    dist.destroy_process_group()


def test_nccl_reduce_of_the_sum_buffers(product, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b, red = np.load(tmp_path / "rank0.npy"), np.load(tmp_path / "rank1.npy"), np.load(tmp_path / "reduced.npy")
    assert a.sum() > 0 and b.sum() > 0 and not np.array_equal(a, b)
    assert (red.view(np.uint32) == (a + b).view(np.uint32)).all()
    for k in range(2):
        assert (np.load(tmp_path / ("reduced_capi%d.npy" % k)).view(np.uint32) == (a + b).view(np.uint32)).all()
    assert np.array_equal(np.load(tmp_path / "after0.npy"), a) and np.array_equal(np.load(tmp_path / "after1.npy"), b)


def test_group_of_two_gpus_in_one_process(product):
    """opr_group_*: one process, two GPUs, the hand-out loop and the reduce both inside the C++ library; the total equals the sum
    of two single renderers given the same seeds, iterations and radii."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import oppositerenderer_b200 as opr
    from oppositerenderer_b200 import distributed as D
    from oppositerenderer_b200.renderer import RendererGroup
    scene = opr.CornellSmall()
    cam = scene.getDefaultCamera()
    cam.setAspectRatio(1.0)
    det = opr.RenderServerRenderRequestDetails(cam, scene.getSceneName(), opr.RenderMethod.PROGRESSIVE_PHOTON_MAPPING, SIZE, SIZE, 2.0 / 3.0)
    g = RendererGroup([0, 1], photon_launch=PL, seed=SEED)
    assert g.size() == 2
    g.initScene(scene)
    g.renderNextIterations(0, 2 * LOCAL_ITERS, 0.05, 2.0 / 3.0, det)
    total = g.reduceOutput(np.zeros((SIZE, SIZE, 3), np.float32))
    again = g.reduceOutput(np.zeros((SIZE, SIZE, 3), np.float32))
    g.close()
    expect = np.zeros((SIZE, SIZE, 3), np.float32)
    for rank in range(2):
        r = opr.OptixRenderer(photon_launch=PL, seed=D.rank_seed(SEED, rank))
        r.initialize(opr.ComputeDevice(rank))
        r.initScene(scene)
        D.render_partition(r, D.IterationPartition(rank, 2, 0.05), det, 0, LOCAL_ITERS)
        expect = expect + r.getOutputBuffer() if rank else r.getOutputBuffer().copy()
        r.close()
    assert (total.view(np.uint32) == expect.view(np.uint32)).all()
    assert np.array_equal(total, again)
