"""Iteration partitioning across the GPUs of one box — the replacement of the reference's Gigabit client/server
distribution (Client/DistributedApplication.cpp:96-122, Server/server/RenderServerRenderer.cpp:154-172).

The reference's client hands out global iteration numbers with their radii to render servers and adds the returned
W*H float3 buffers (Client/client/RenderResultPacketReceiver + Gui/gui/RenderWidget.cpp:120,197).  Here every rank
(one process per GPU, scene + BVH replicated) renders the global iterations ``rank, rank + G, rank + 2G, ...`` with
``localIterationNumber = i // G``; the radius of global iteration ``i`` is a pure function of ``i`` (closed form of
Standalone/StandaloneRenderManager.cpp:105-107, ``opr_ppm_radius_for_iteration``), so no radius is exchanged.  The
only data-path collective is ONE reduce (sum) of the running-sum buffers (12*W*H bytes per rank) onto rank 0 per
output request — ``torch.distributed`` NCCL over NVLink on the GPU box, gloo in the CPU tests.

Each rank needs its own RNG stream: the reference's servers got that implicitly from wall-clock seeds
(OptixRenderer_SpatialHash.cu:322); ``rank_seed`` makes it explicit and reproducible.
"""
from . import capi

# State i of a renderer is seeded with `seed + i` (rng_init_kernel, OptixRenderer_SpatialHash.cu:310-347), i < rsW * rsH.  Two
# ranks must therefore differ by MORE than the whole state buffer, or rank r's state j is rank 0's state j + r * stride and both
# walk the same photon paths.  2^28 states (16384 x 16384) per rank, 16 ranks: disjoint seed ranges modulo 2^32.
# The same rule lives in C++ (opr_group_rank_seed, csrc/group.cu); tests/test_distributed_cpu.py keeps them equal.
SEED_STRIDE = 1 << 28
MAX_RANKS = (1 << 32) // SEED_STRIDE


def rank_seed(base_seed, rank):
    if not (0 <= int(rank) < MAX_RANKS):
        raise ValueError("rank %d: disjoint per-rank RNG seed ranges exist for %d ranks" % (rank, MAX_RANKS))
    return (int(base_seed) + int(rank) * SEED_STRIDE) & 0xFFFFFFFF


def check_state_count(num_states):
    """Raise unless a state buffer of `num_states` entries keeps the seed ranges of different ranks disjoint."""
    if int(num_states) > SEED_STRIDE:
        raise ValueError("%d RNG states exceed the per-rank seed stride %d: ranks would share states" % (num_states, SEED_STRIDE))


def global_iteration(rank, world, local_iteration):
    """Global iteration number rendered by `rank` as its `local_iteration`-th step."""
    return int(local_iteration) * int(world) + int(rank)


def owner_of(iteration, world):
    """(rank, localIterationNumber) that renders global iteration `iteration`."""
    return int(iteration) % int(world), int(iteration) // int(world)


def radius_for_iteration(r0, iteration, alpha=2.0 / 3.0):
    """r_i, with r_{i+1}^2 = r_i^2 (i + alpha)/(i + 1) carried in host double exactly like the driver loop."""
    return float(capi.load_library().opr_ppm_radius_for_iteration(float(r0), int(iteration), float(alpha)))


class IterationPartition:
    """The (global iteration, local iteration, radius) triples of one rank."""

    def __init__(self, rank, world, initial_radius, alpha=2.0 / 3.0):
        if not (0 <= rank < world):
            raise ValueError("rank %d outside world of %d" % (rank, world))
        self.rank, self.world, self.r0, self.alpha = int(rank), int(world), float(initial_radius), float(alpha)

    def step(self, local_iteration):
        it = global_iteration(self.rank, self.world, local_iteration)
        return it, int(local_iteration), radius_for_iteration(self.r0, it, self.alpha)

    def steps(self, first_local, count):
        for k in range(count):
            yield self.step(first_local + k)


def render_partition(renderer, partition, details, first_local, count, readback=None):
    """Run `count` local iterations of `partition` on `renderer` (anything with the OptixRenderer call signature)."""
    for it, local, radius in partition.steps(first_local, count):
        renderer.renderNextIteration(it, local, radius, True, details)
        if readback is not None:
            renderer.getOutputBuffer(readback)


def reduce_sum_to_root(tensor, dst=0, group=None, out=None):
    """One reduce (sum) of the per-rank running-sum buffers onto `dst`.  The renderer's accumulator is NEVER the target: the sum
    lands in `out` (or a fresh clone), so the local running sum stays this rank's own and the call can be repeated during a
    progressive render (a reduce in place would fold the other ranks' earlier sums into rank 0's accumulator and count them again
    at the next reduce).  Returns the tensor that holds the total on `dst` (undefined content elsewhere).  No-op copy in a world of
    one.  The caller divides by the TOTAL number of iterations (all ranks) afterwards, like RenderWidget.cpp:197."""
    import torch.distributed as dist
    if out is None:
        out = tensor.clone()
    else:
        out.copy_(tensor)
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return out
    dist.reduce(out, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return out


def output_as_tensor(renderer, device):
    """Zero-copy torch view (H, W, 3 float32) of the renderer's running-sum buffer in HBM."""
    import torch
    h, w = renderer.getHeight(), renderer.getWidth()

    class _Arr:
        __cuda_array_interface__ = {"shape": (h, w, 3), "typestr": "<f4", "data": (renderer.output_device_ptr(), False), "version": 2}
    return torch.as_tensor(_Arr(), device=device)
