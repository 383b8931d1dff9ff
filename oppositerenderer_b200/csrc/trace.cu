// trace.cu — wavefront closest-hit / any-hit traversal kernels over SoA ray buffers.  Every rtTrace call
// site of the reference (RayGeneratorPPM.cu:53, PhotonGenerator.cu:120, Diffuse.cu:134,
// Mirror.cu:60,75, Glass.cu:135,203, Texture.cu:174, light.h:69, vcm.h:56, VCMLightPass.cu:76,
// VCMCameraPass.cu:64) becomes a push into a ray queue that these kernels drain.
//
// Algorithmic bytes: closest-hit 48 B/ray (32 B ray in, 16 B hit out) + 4 B queue index; any-hit 36 B/ray.
#include "renderer.h"

namespace opr {

constexpr int TRACE_THREADS = 128;

// Closest hit for the rays [0, *count) of queue `cur`; each hit ray is appended to the queue of its
// material kind (warp-aggregated: one atomicAdd per distinct kind per warp via __match_any_sync), so the
// shading kernels run sorted by material.
__global__ void __launch_bounds__(TRACE_THREADS) trace_closest_kernel(SceneView sc, const float4* __restrict__ rayO, const float4* __restrict__ rayD,
                                                                      const unsigned* __restrict__ count, float4* __restrict__ hits,
                                                                      unsigned* __restrict__ matQueue, unsigned* __restrict__ matCount,
                                                                      unsigned capacity, unsigned* __restrict__ stat)
{
    const unsigned n = *count;
    const unsigned lane = threadIdx.x & 31;
    for (unsigned base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
        const unsigned i = base + threadIdx.x;
        const bool active = i < n;
        int kind = -1;
        if (active) {
            const float4 o4 = rayO[i], d4 = rayD[i];
            HitRec hit;
            traverse<false>(sc, f3(o4), f3(d4), o4.w, d4.w, hit);
            hits[i] = make_float4(hit.t, __int_as_float(hit.prim), hit.u, hit.v);
            kind = hit.prim < 0 ? QUEUE_MISS : sc.materials[prim_material(sc, hit.prim)].kind;
        }
        const unsigned peers = __match_any_sync(0xffffffffu, kind);
        if (active) {
            const unsigned leader = __ffs(peers) - 1;
            unsigned slot = 0;
            if (lane == leader) slot = atomicAdd(&matCount[kind], __popc(peers));
            slot = __shfl_sync(peers, slot, leader);
            matQueue[(size_t)kind * capacity + slot + __popc(peers & ((1u << lane) - 1))] = i;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(stat, n);
}

void launch_trace_closest(const SceneView& sc, const RayQueues& q, int cur, LaunchCtx lc)
{
    int blocks = 148 * 16;
    trace_closest_kernel<<<blocks, TRACE_THREADS, 0, lc.st>>>(sc, q.rayO[cur], q.rayD[cur], q.counters + cur, q.hits, q.matQueue,
                                                              q.counters + 2, q.capacity, q.counters + 16 + STAT_CLOSEST_RAYS);
    ++*lc.launches;
}

void launch_trace_closest_bounce(const SceneView& sc, const RayQueues& q, int bounce, LaunchCtx lc)
{
    const int cur = bounce & 1;
    trace_closest_kernel<<<148 * 16, TRACE_THREADS, 0, lc.st>>>(sc, q.rayO[cur], q.rayD[cur], q.counters + CTR_RAYCOUNT + bounce, q.hits, q.matQueue,
                                                                q.counters + CTR_MATCOUNT + 8 * bounce, q.capacity, q.counters + CTR_STATS + STAT_CLOSEST_RAYS);
    ++*lc.launches;
}

// Raw entry points over caller-supplied SoA rays (C-ABI opr_trace_closest / opr_trace_any)
__global__ void __launch_bounds__(TRACE_THREADS) trace_closest_raw_kernel(SceneView sc, const float4* __restrict__ rayO, const float4* __restrict__ rayD,
                                                                          unsigned n, float4* __restrict__ hits)
{
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 o4 = rayO[i], d4 = rayD[i];
        HitRec hit;
        traverse<false>(sc, f3(o4), f3(d4), o4.w, d4.w, hit);
        hits[i] = make_float4(hit.t, __int_as_float(hit.prim), hit.u, hit.v);
    }
}

__global__ void __launch_bounds__(TRACE_THREADS) trace_any_raw_kernel(SceneView sc, const float4* __restrict__ rayO, const float4* __restrict__ rayD,
                                                                      unsigned n, int* __restrict__ occluded)
{
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 o4 = rayO[i], d4 = rayD[i];
        HitRec hit;
        occluded[i] = traverse<true>(sc, f3(o4), f3(d4), o4.w, d4.w, hit) ? 1 : 0;
    }
}

void launch_trace_closest_raw(const SceneView& sc, const float4* rayO, const float4* rayD, unsigned n, float4* hits, LaunchCtx lc)
{
    if (!n) return;
    int blocks = (int)std::min<unsigned>((n + TRACE_THREADS - 1) / TRACE_THREADS, 148 * 16);
    trace_closest_raw_kernel<<<blocks, TRACE_THREADS, 0, lc.st>>>(sc, rayO, rayD, n, hits);
    ++*lc.launches;
}

void launch_trace_any_raw(const SceneView& sc, const float4* rayO, const float4* rayD, unsigned n, int* occluded, LaunchCtx lc)
{
    if (!n) return;
    int blocks = (int)std::min<unsigned>((n + TRACE_THREADS - 1) / TRACE_THREADS, 148 * 16);
    trace_any_raw_kernel<<<blocks, TRACE_THREADS, 0, lc.st>>>(sc, rayO, rayD, n, occluded);
    ++*lc.launches;
}

} // namespace opr
