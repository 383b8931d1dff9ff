"""Two ranks on two B200s over NCCL (skipped with fewer than two GPUs): the iteration partition drives the CUDA renderer, the
running-sum buffers are reduced in place through the zero-copy device view on the renderer's own stream, and the result on
rank 0 equals the host-side sum of the two ranks' buffers bit for bit (SURVEY.md §8e parity caveat (i))."""
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
    stream = torch.cuda.ExternalStream(capi.load_library().opr_get_stream(r._h), device=dev)
    with torch.cuda.stream(stream):
        D.reduce_sum_to_root(D.output_as_tensor(r, dev), dst=0)
    if rank == 0:
        np.save(os.path.join(outdir, "reduced.npy"), r.getOutputBuffer())      # the reduce wrote into the renderer's own sum buffer
    dist.barrier()
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
