"""World-size-2 `gloo` tests of the multi-GPU host logic: iteration partition, closed-form radius,
per-rank seeds and the single reduce of the running-sum buffers.  The renderer behind the partition is the CPU oracle
here (there is no GPU in this container); on the GPU box the same module drives the CUDA renderer over NCCL (bench.py)."""
import os
import socket
import sys

import numpy as np
import pytest

import oracle_lib as ol
from oppositerenderer_b200 import capi, distributed as D

SIZE, PL, LOCAL_ITERS, SEED = 24, (16, 16), 2, 1645301512


class _OracleAsRenderer:
    """OptixRenderer call signature over the oracle, so render_partition can drive it."""

    def __init__(self, seed):
        self.o = ol.OracleRenderer("oracle", PL, 10, seed)
        self.scene = ol.BuiltinScene("Cornell")
        self.o.init_scene(self.scene.desc)
        self.calls = []

    def renderNextIteration(self, it, local, radius, createOutput, details):
        self.calls.append((it, local, radius))
        self.o.render_next_iteration(it, local, radius, details)

    def getOutputBuffer(self, data=None):
        return self.o.output()


def _render_rank(rank, world):
    r = _OracleAsRenderer(D.rank_seed(SEED, rank))
    det = ol.make_details(r.scene.camera(1.0), capi.PROGRESSIVE_PHOTON_MAPPING, SIZE, SIZE)
    part = D.IterationPartition(rank, world, r.scene.initial_radius())
    D.render_partition(r, part, det, 0, LOCAL_ITERS)
    return r.getOutputBuffer(), r.calls


def _worker(rank, world, port, outdir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    out, calls = _render_rank(rank, world)
    np.save(os.path.join(outdir, "rank%d.npy" % rank), out)
    np.save(os.path.join(outdir, "calls%d.npy" % rank), np.array(calls, np.float64))
    t = torch.from_numpy(out.copy())
    D.reduce_sum_to_root(t, dst=0)
    if rank == 0:
        np.save(os.path.join(outdir, "reduced.npy"), t.numpy())
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_partition_covers_every_iteration_once():
    for world in (1, 2, 4, 8):
        seen = {}
        for rank in range(world):
            p = D.IterationPartition(rank, world, 7.5)
            for it, local, radius in p.steps(0, 5):
                assert it not in seen
                seen[it] = (rank, local)
                assert D.owner_of(it, world) == (rank, local)
        assert sorted(seen) == list(range(5 * world))
    with pytest.raises(ValueError):
        D.IterationPartition(2, 2, 1.0)


def test_closed_form_radius_matches_the_driver_recurrence(product):
    # Standalone/StandaloneRenderManager.cpp:105-107 carried sequentially in double vs the per-iteration closed form
    seq = ol.ppm_radius_schedule(7.4783, 200)
    for i in (0, 1, 2, 7, 63, 199):
        assert D.radius_for_iteration(7.4783, i) == pytest.approx(seq[i], rel=1e-12)
        assert np.float32(D.radius_for_iteration(7.4783, i)) == np.float32(seq[i])


def test_rank_seeds_are_distinct():
    seeds = {D.rank_seed(SEED, r) for r in range(8)}
    assert len(seeds) == 8 and D.rank_seed(SEED, 0) == SEED


def test_two_rank_gloo_reduce_equals_sum_of_rank_buffers(oracle, product, tmp_path):
    """(i) reduce(sum) over gloo == CPU sum of the two per-rank buffers, bit for bit; (ii) the two ranks rendered global
    iterations {0,2} and {1,3} with the sequential schedule's radii; (iii) different seeds give different images."""
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    a, b = np.load(tmp_path / "rank0.npy"), np.load(tmp_path / "rank1.npy")
    red = np.load(tmp_path / "reduced.npy")
    assert (red.view(np.uint32) == (a + b).view(np.uint32)).all()
    assert not np.array_equal(a, b)
    seq = ol.ppm_radius_schedule(ol.BuiltinScene("Cornell").initial_radius(), world * LOCAL_ITERS)
    for rank in range(world):
        calls = np.load(tmp_path / ("calls%d.npy" % rank))
        assert [int(c[0]) for c in calls] == [rank, rank + world]
        assert [int(c[1]) for c in calls] == [0, 1]
        for c in calls:
            assert np.float32(c[2]) == np.float32(seq[int(c[0])])
    # the in-process rendering of rank 1's partition reproduces what the spawned rank produced (determinism of the partition)
    again, _ = _render_rank(1, world)
    assert (again.view(np.uint32) == b.view(np.uint32)).all()
    # normalised image = sum / total iterations is a plausible Cornell render (non-negative, not empty)
    img = red / (world * LOCAL_ITERS)
    assert (img >= 0).all() and img.mean() > 0


if __name__ == "__main__":
    sys.exit(pytest.main([__file__, "-q"]))
