// trace.cu — wavefront closest-hit / any-hit traversal kernels over SoA ray buffers.  Every rtTrace call
// site of the reference (RayGeneratorPPM.cu:53, PhotonGenerator.cu:120, Diffuse.cu:134,
// Mirror.cu:60,75, Glass.cu:135,203, Texture.cu:174, light.h:69, vcm.h:56, VCMLightPass.cu:76,
// VCMCameraPass.cu:64) becomes a push into a ray queue that these kernels drain.
//
// Algorithmic bytes: closest-hit 48 B/ray (32 B ray in, 16 B hit out) + 4 B queue index; any-hit 36 B/ray.
#include "renderer.h"
#include <cstdlib>

namespace opr {

constexpr int TRACE_THREADS = 128;

// Traversal stack of the wavefront kernels: the first OPR_SSTACK entries of every lane live in shared memory, laid out
// [entry][thread] so that a lane's bank is its thread index whatever its depth — a push or a pop of a warp whose lanes sit at 17
// different depths is two shared-memory wavefronts (8-byte entries), where the same access to the per-thread local-memory stack
// touched two 128-byte lines per DISTINCT depth, on the L1 data pipe that already bounds incoherent traversal (profiles/).  Deeper
// entries (rare: the bound is 3 * levels + 4) spill to the local array.  0 = everything in local memory.
#ifndef OPR_SSTACK
#define OPR_SSTACK 8
#endif
#if OPR_SSTACK > 0
#undef OPR_STACK_ST
#undef OPR_STACK_LD
#define OPR_STACK_ST(I, V) { if ((I) < OPR_SSTACK) sstack[(I) * TRACE_THREADS] = (V); else stack[(I) - OPR_SSTACK] = (V); }
#define OPR_STACK_LD(DST, I) { if ((I) < OPR_SSTACK) DST = sstack[(I) * TRACE_THREADS]; else DST = stack[(I) - OPR_SSTACK]; }
#define OPR_STACK_DECL                                                                    \
    __shared__ int2 sstackMem[OPR_SSTACK * TRACE_THREADS];                               \
    int2* const sstack = sstackMem + threadIdx.x;                                        \
    int2 stack[TRAVERSAL_STACK - OPR_SSTACK];
#else
#define OPR_STACK_DECL int2 stack[TRAVERSAL_STACK];
#endif

// Closest hit for the rays [0, *count) of a queue; each finished ray is appended to the queue of its material kind
// (warp-aggregated: one atomicAdd per distinct kind per warp via __match_any_sync), so the shading kernels run sorted by
// material.
//
// Persistent warps with dynamic ray fetch: a warp keeps 32 per-lane traversal states; when at least `refillIdle` lanes
// have finished their ray (or all have), the finished rays are retired (hit record + material queue) and the idle lanes
// pull fresh rays from the shared cursor, so that warps stay full while ray lengths differ by an order of magnitude
// (incoherent photon bounces).  The node / leaf steps are the warp-lockstep ones of scene.cuh.
#ifndef OPR_TRACE_MINBLOCKS
#define OPR_TRACE_MINBLOCKS 1
#endif
template <bool QN>
__global__ void __launch_bounds__(TRACE_THREADS, OPR_TRACE_MINBLOCKS) trace_closest_kernel(SceneView sc, const float4* __restrict__ rayO, const float4* __restrict__ rayD,
                                                                      const unsigned* __restrict__ count, float4* __restrict__ hits,
                                                                      unsigned* __restrict__ matQueue, unsigned* __restrict__ matCount,
                                                                      unsigned capacity, unsigned* __restrict__ stat, unsigned* __restrict__ fetchCursor,
                                                                      unsigned refillIdle)
{
    const unsigned FULL = 0xffffffffu, NONE = 0xffffffffu;
    const unsigned n = *count;
    const unsigned lane = threadIdx.x & 31;
    float3 o = f3(0.f), d = f3(0.f), idir = f3(0.f), ood = f3(0.f);
    float tmin = 0.f;
    HitRec hit; hit.t = 0.f; hit.prim = -1; hit.u = hit.v = 0.f;
    OPR_STACK_DECL
    int sp = 0;
    int cur = TRAVERSAL_DONE;
    unsigned ray = NONE;
    bool exhausted = false;   // warp-uniform
    for (;;) {
        const unsigned idleMask = __ballot_sync(FULL, cur == TRAVERSAL_DONE);
        if (idleMask == FULL || (!exhausted && (unsigned)__popc(idleMask) >= refillIdle)) {
            // ---- retire finished rays
            const bool finished = (cur == TRAVERSAL_DONE) && (ray != NONE);
            int kind = -1;
            if (finished) {
                const int prim = hit.prim >= 0 ? __ldg(&sc.leafPrim[hit.prim]) : -1;
                hits[ray] = make_float4(hit.t, __int_as_float(prim), hit.u, hit.v);
                kind = prim < 0 ? QUEUE_MISS : sc.materials[prim_material(sc, prim)].kind;
            }
            const unsigned peers = __match_any_sync(FULL, kind);
            if (finished) {
                const unsigned leader = __ffs(peers) - 1;
                unsigned slot = 0;
                if (lane == leader) slot = atomicAdd(&matCount[kind], __popc(peers));
                slot = __shfl_sync(peers, slot, leader);
                matQueue[(size_t)kind * capacity + slot + __popc(peers & ((1u << lane) - 1))] = ray;
                ray = NONE;
            }
            // ---- refill idle lanes
            if (!exhausted) {
                const unsigned want = __popc(idleMask);
                unsigned base = 0;
                if (lane == 0) base = atomicAdd(fetchCursor, want);
                base = __shfl_sync(FULL, base, 0);
                if (cur == TRAVERSAL_DONE) {
                    const unsigned idx = base + __popc(idleMask & ((1u << lane) - 1));
                    if (idx < n) {
                        const float4 o4 = rayO[idx], d4 = rayD[idx];
                        o = f3(o4); d = f3(d4); tmin = o4.w;
                        ray_slab_setup<QN>(sc, o, d, idir, ood);
                        hit.t = d4.w; hit.prim = -1; hit.u = hit.v = 0.f;
                        sp = 0; cur = sc.rootRef; ray = idx;
                    }
                }
                if (base + want >= n) exhausted = true;
            }
            if (__all_sync(FULL, cur == TRAVERSAL_DONE)) break;
        }
        OPR_NODE_PHASE(FULL, false, QN)
        if (cur < 0 && cur != TRAVERSAL_DONE) {
            [[maybe_unused]] bool found = false;
            OPR_LEAF_STEP(found)      // a closest-hit query keeps going after an accepted primitive: hit.t only shrinks the search
            OPR_POP(false)
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(stat, n);
}

static unsigned trace_blocks_per_sm()
{
    static int v = -1;
    if (v < 0) { const char* e = getenv("OPR_TRACE_BLOCKS_PER_SM"); v = e ? atoi(e) : 8; if (v < 1) v = 1; if (v > 16) v = 16; }
    return (unsigned)v;
}

static unsigned trace_refill_idle()
{
    static int v = -1;
    if (v < 0) { const char* e = getenv("OPR_TRACE_REFILL"); v = e ? atoi(e) : 24; if (v < 1) v = 1; if (v > 32) v = 32; }
    return (unsigned)v;
}

// legacy ping-pong counters ([cur] ray count, [2..9) material counts) used by the VCM passes
void launch_trace_closest(const SceneView& sc, const RayQueues& q, int cur, LaunchCtx lc, bool quantized)
{
    OPR_CUDA_CHECK(cudaMemsetAsync(q.counters + CTR_FETCH_LEGACY, 0, sizeof(unsigned), lc.st));
    auto k = quantized ? trace_closest_kernel<true> : trace_closest_kernel<false>;
    k<<<numSMs() * trace_blocks_per_sm(), TRACE_THREADS, 0, lc.st>>>(sc, q.rayO[cur], q.rayD[cur], q.counters + cur, q.hits, q.matQueue,
                                                               q.counters + 2, q.capacity, q.counters + CTR_STATS + STAT_CLOSEST_RAYS,
                                                               q.counters + CTR_FETCH_LEGACY, trace_refill_idle());
    ++*lc.launches;
}

void launch_trace_closest_bounce(const SceneView& sc, const RayQueues& q, int bounce, LaunchCtx lc, bool quantized, bool wholeTiles)
{
    // wholeTiles (primary rays, whose 32 consecutive slots are an 8 x 4-pixel tile): a warp takes new rays only when all of its lanes are done, so
    // it always holds one tile (primary trace 0.381 -> 0.363 ms; the incoherent traces lose 12 % with the same rule and keep the threshold)
    const int cur = bounce & 1;
    auto k = quantized ? trace_closest_kernel<true> : trace_closest_kernel<false>;
    k<<<numSMs() * trace_blocks_per_sm(), TRACE_THREADS, 0, lc.st>>>(sc, q.rayO[cur], q.rayD[cur], q.counters + CTR_RAYCOUNT + bounce, q.hits, q.matQueue,
                                                               q.counters + CTR_MATCOUNT + 8 * bounce, q.capacity, q.counters + CTR_STATS + STAT_CLOSEST_RAYS,
                                                               q.counters + CTR_FETCH + bounce, wholeTiles ? 32u : trace_refill_idle());
    ++*lc.launches;
}

// Any-hit for a compacted queue of shadow rays (ppm/DirectRadianceEstimation.cu:79-83: any intersection in (tmin, tmax)
// zeroes the sample).  Ray i = (origin = hit point of pixel idx >> 2, direction + tmax = shadowD[i]); an occluded ray
// clears sampleFactor[idx].  Same persistent / dynamic-fetch / warp-lockstep structure as trace_closest_kernel.
template <bool QN>
__global__ void __launch_bounds__(TRACE_THREADS, OPR_TRACE_MINBLOCKS) trace_shadow_queue_kernel(SceneView sc, const float4* __restrict__ hitPos, const float4* __restrict__ shadowD,
                                                                           const unsigned* __restrict__ shadowIdx, const unsigned* __restrict__ count,
                                                                           float* __restrict__ sampleFactor, unsigned* __restrict__ fetchCursor,
                                                                           unsigned* __restrict__ stat, unsigned refillIdle)
{
    const unsigned FULL = 0xffffffffu;
    const unsigned n = *count;
    const unsigned lane = threadIdx.x & 31;
    float3 o = f3(0.f), d = f3(0.f), idir = f3(0.f), ood = f3(0.f);
    const float tmin = 0.0001f;
    HitRec hit; hit.t = 0.f; hit.prim = -1; hit.u = hit.v = 0.f;
    OPR_STACK_DECL
    int sp = 0;
    int cur = TRAVERSAL_DONE;
    unsigned idx = 0;
    bool exhausted = false;
    for (;;) {
        const unsigned idleMask = __ballot_sync(FULL, cur == TRAVERSAL_DONE);
        if (idleMask == FULL || (!exhausted && (unsigned)__popc(idleMask) >= refillIdle)) {
            if (!exhausted) {
                const unsigned want = __popc(idleMask);
                unsigned base = 0;
                if (lane == 0) base = atomicAdd(fetchCursor, want);
                base = __shfl_sync(FULL, base, 0);
                if (cur == TRAVERSAL_DONE) {
                    const unsigned i = base + __popc(idleMask & ((1u << lane) - 1));
                    if (i < n) {
                        idx = shadowIdx[i];
                        const float4 d4 = shadowD[i];
                        o = f3(hitPos[idx >> 2]); d = f3(d4);
                        ray_slab_setup<QN>(sc, o, d, idir, ood);
                        hit.t = d4.w; hit.prim = -1;
                        sp = 0; cur = sc.rootRef;
                    }
                }
                if (base + want >= n) exhausted = true;
            }
            if (__all_sync(FULL, cur == TRAVERSAL_DONE)) break;
        }
        OPR_NODE_PHASE(FULL, true, QN)
        if (cur < 0 && cur != TRAVERSAL_DONE) {
            bool found = false;
            OPR_LEAF_STEP(found)
            if (found) { sampleFactor[idx] = 0.f; cur = TRAVERSAL_DONE; sp = 0; }
            else OPR_POP(true)
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(stat, n);
}

void launch_trace_shadow_queue(const SceneView& sc, const float4* hitPos, const float4* shadowD, const unsigned* shadowIdx, const unsigned* count,
                               float* sampleFactor, unsigned* fetchCursor, unsigned* stat, LaunchCtx lc, bool quantized)
{
    auto k = quantized ? trace_shadow_queue_kernel<true> : trace_shadow_queue_kernel<false>;
    k<<<numSMs() * trace_blocks_per_sm(), TRACE_THREADS, 0, lc.st>>>(sc, hitPos, shadowD, shadowIdx, count, sampleFactor, fetchCursor, stat, trace_refill_idle());
    ++*lc.launches;
}

// Any-hit for the deferred connections of the BDPT passes (vcm.cu): segment i = (connO[i].xyz, connD[i].xyz) over
// (EPS_RAY, connO[i].w); when nothing occludes it (isOccluded, vcm/vcm.h:43-59) its contribution connC[i] is added to
// target[stride * index] (camera-path colour accumulators, or the image for light-tracing splats: vcm.h:159).
template <bool QN>
__global__ void __launch_bounds__(TRACE_THREADS, OPR_TRACE_MINBLOCKS) trace_connect_queue_kernel(SceneView sc, const float4* __restrict__ connO, const float4* __restrict__ connD,
                                                                            const float4* __restrict__ connC, const unsigned* __restrict__ count,
                                                                            float* __restrict__ target, unsigned stride, unsigned* __restrict__ fetchCursor,
                                                                            unsigned* __restrict__ stat, unsigned refillIdle, unsigned* __restrict__ dbgCount)
{
    const unsigned FULL = 0xffffffffu, NONE = 0xffffffffu;
    const unsigned n = *count;
    const unsigned lane = threadIdx.x & 31;
    float3 o = f3(0.f), d = f3(0.f), idir = f3(0.f), ood = f3(0.f);
    const float tmin = EPS_RAY;
    HitRec hit; hit.t = 0.f; hit.prim = -1; hit.u = hit.v = 0.f;
    OPR_STACK_DECL
    int sp = 0;
    int cur = TRAVERSAL_DONE;
    unsigned pending = NONE;      // queue index of a finished, unoccluded segment whose contribution is not added yet
    bool exhausted = false;
    for (;;) {
        const unsigned idleMask = __ballot_sync(FULL, cur == TRAVERSAL_DONE);
        if (idleMask == FULL || (!exhausted && (unsigned)__popc(idleMask) >= refillIdle)) {
            if (cur == TRAVERSAL_DONE && pending != NONE) {
                const float4 c = connC[pending];
                const unsigned ti = __float_as_uint(connD[pending].w);
                float* t = target + (size_t)stride * ti;
                if (dbgCount) atomicAdd(dbgCount + ti, 1u);     // parity diagnostics only (null otherwise)
                atomicAdd(t, c.x); atomicAdd(t + 1, c.y); atomicAdd(t + 2, c.z);
                pending = NONE;
            }
            if (!exhausted) {
                const unsigned want = __popc(idleMask);
                unsigned base = 0;
                if (lane == 0) base = atomicAdd(fetchCursor, want);
                base = __shfl_sync(FULL, base, 0);
                if (cur == TRAVERSAL_DONE) {
                    const unsigned i = base + __popc(idleMask & ((1u << lane) - 1));
                    if (i < n) {
                        const float4 o4 = connO[i], d4 = connD[i];
                        o = f3(o4); d = f3(d4);
                        ray_slab_setup<QN>(sc, o, d, idir, ood);
                        hit.t = o4.w; hit.prim = -1;
                        sp = 0; cur = sc.rootRef; pending = i;
                    }
                }
                if (base + want >= n) exhausted = true;
            }
            if (__all_sync(FULL, cur == TRAVERSAL_DONE)) break;
        }
        OPR_NODE_PHASE(FULL, true, QN)
        if (cur < 0 && cur != TRAVERSAL_DONE) {
            bool found = false;
            OPR_LEAF_STEP(found)
            if (found) { pending = NONE; cur = TRAVERSAL_DONE; sp = 0; }
            else OPR_POP(true)
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(stat, n);
}

void launch_trace_connect_queue(const SceneView& sc, const float4* connO, const float4* connD, const float4* connC, const unsigned* count, float* target,
                                unsigned stride, unsigned* fetchCursor, unsigned* stat, LaunchCtx lc, unsigned blocksPerSM, unsigned* dbgCount, bool quantized)
{
    const unsigned blocks = numSMs() * (blocksPerSM ? blocksPerSM : trace_blocks_per_sm());
    auto k = quantized ? trace_connect_queue_kernel<true> : trace_connect_queue_kernel<false>;
    k<<<blocks, TRACE_THREADS, 0, lc.st>>>(sc, connO, connD, connC, count, target, stride, fetchCursor, stat, trace_refill_idle(), dbgCount);
    ++*lc.launches;
}

// Raw entry points over caller-supplied SoA rays (C-ABI opr_trace_closest / opr_trace_any)
template <bool QN>
__global__ void __launch_bounds__(TRACE_THREADS) trace_closest_raw_kernel(SceneView sc, const float4* __restrict__ rayO, const float4* __restrict__ rayD,
                                                                          unsigned n, float4* __restrict__ hits)
{
    for (unsigned base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
        const unsigned i = base + threadIdx.x;
        const bool active = i < n;
        const float4 o4 = active ? rayO[i] : make_float4(0.f, 0.f, 0.f, 0.f), d4 = active ? rayD[i] : make_float4(0.f, 0.f, 1.f, 0.f);
        HitRec hit;
        traverse_warp<false, QN>(sc, 0xffffffffu, active, f3(o4), f3(d4), o4.w, d4.w, hit);
        if (active) hits[i] = make_float4(hit.t, __int_as_float(hit.prim), hit.u, hit.v);
    }
}

template <bool QN>
__global__ void __launch_bounds__(TRACE_THREADS) trace_any_raw_kernel(SceneView sc, const float4* __restrict__ rayO, const float4* __restrict__ rayD,
                                                                      unsigned n, int* __restrict__ occluded)
{
    for (unsigned base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
        const unsigned i = base + threadIdx.x;
        const bool active = i < n;
        const float4 o4 = active ? rayO[i] : make_float4(0.f, 0.f, 0.f, 0.f), d4 = active ? rayD[i] : make_float4(0.f, 0.f, 1.f, 0.f);
        HitRec hit;
        const bool occ = traverse_warp<true, QN>(sc, 0xffffffffu, active, f3(o4), f3(d4), o4.w, d4.w, hit);
        if (active) occluded[i] = occ ? 1 : 0;
    }
}

void launch_trace_closest_raw(const SceneView& sc, const float4* rayO, const float4* rayD, unsigned n, float4* hits, LaunchCtx lc, bool quantized)
{
    if (!n) return;
    int blocks = (int)std::min<unsigned>((n + TRACE_THREADS - 1) / TRACE_THREADS, numSMs() * 16);
    auto k = quantized ? trace_closest_raw_kernel<true> : trace_closest_raw_kernel<false>;
    k<<<blocks, TRACE_THREADS, 0, lc.st>>>(sc, rayO, rayD, n, hits);
    ++*lc.launches;
}

void launch_trace_any_raw(const SceneView& sc, const float4* rayO, const float4* rayD, unsigned n, int* occluded, LaunchCtx lc, bool quantized)
{
    if (!n) return;
    int blocks = (int)std::min<unsigned>((n + TRACE_THREADS - 1) / TRACE_THREADS, numSMs() * 16);
    auto k = quantized ? trace_any_raw_kernel<true> : trace_any_raw_kernel<false>;
    k<<<blocks, TRACE_THREADS, 0, lc.st>>>(sc, rayO, rayD, n, occluded);
    ++*lc.launches;
}

} // namespace opr
