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
    total = D.reduce_sum_to_root(t, dst=0)
    assert (t.numpy().view(np.uint32) == out.view(np.uint32)).all()      # the local running sum is never a reduce target
    again = D.reduce_sum_to_root(t, dst=0)                                # ... so a second request during a progressive render gives the same total
    if rank == 0:
        assert (again.numpy().view(np.uint32) == total.numpy().view(np.uint32)).all()
        np.save(os.path.join(outdir, "reduced.npy"), total.numpy())
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


def test_rank_seed_ranges_are_disjoint_over_the_whole_state_buffer(product):
    """State i is seeded with seed + i (OptixRenderer_SpatialHash.cu:310-347): two ranks share random streams unless their seeds
    differ by more than the state count (1920x1080 needs 2.07 M states, 3840x2160 8.3 M).  Python and C++ hold the same rule."""
    lib = capi.load_library()
    ns = 3840 * 2160
    ranges = []
    for r in range(D.MAX_RANKS):
        s = D.rank_seed(SEED, r)
        assert s == lib.opr_group_rank_seed(SEED, r)
        ranges.append((s, s + ns))
    for i, (a0, a1) in enumerate(ranges):
        for b0, b1 in ranges[i + 1:]:
            for shift in (-(1 << 32), 0, 1 << 32):          # seeds wrap modulo 2^32
                assert a1 <= b0 + shift or b1 + shift <= a0
    D.check_state_count(ns)
    with pytest.raises(ValueError):
        D.check_state_count(D.SEED_STRIDE + 1)
    with pytest.raises(ValueError):
        D.rank_seed(SEED, D.MAX_RANKS)
    # the first states of two ranks really differ (oracle XORWOW seeding, bit-exact with the CUDA path)
    a, b = np.zeros(6 * 4, np.uint32), np.zeros(6 * 4, np.uint32)
    o = ol.load("oracle")
    o.ora_rng_init(ol.uptr(a), D.rank_seed(SEED, 0), 2_000_000, 4)
    o.ora_rng_init(ol.uptr(b), D.rank_seed(SEED, 2), 2_000_000 - 2 * 104729, 4)   # the old stride made exactly these states equal
    assert not np.array_equal(a, b)


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
