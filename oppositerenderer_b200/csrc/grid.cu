// grid.cu — photon uniform-grid build and the radius-bounded gather of Progressive Photon Mapping.
//
// Build (replaces thrust in RenderEngine/renderer/OptixRenderer_SpatialHash.cu:209-282):
//   photon_bbox      AABB + count of valid photons         (getPhotonsBoundingBox :123-128)
//   [host]           cell size / grid size in fp32         (:62-71, :42-50, :229-248) — on the host, as in the
//                    reference, so that powf/ceil round identically to the oracle (bit-exact keys)
//   grid_keys_hist   cell key per slot + histogram         (calculateHashCellsKernel :152-173)
//   radix sort       (key, slot) pairs, stable             (sortPhotonsByHash :193-196)
//   exclusive scan   histogram -> offset table             (createHashmapOffsetTable :202-207)
//   photon_reorder   one gather of the 36-byte payload into sorted order (thrust moved it every pass)
// Algorithmic bytes: 36N (bbox) + 16N (keys: pos + key + hist) + sort 8N r/w per pass + 72V (reorder) + 8C.
//
// Gather (replaces ppm/IndirectRadianceEstimation.cu:69-227, uniform-grid branch):
// algorithmic bytes 64 B hitpoint + 12 B out per pixel + 36 B per photon touched once + 8 B per cell row.
#include <cstdlib>
#include "renderer.h"
#include <cmath>

namespace opr {

// ---- float atomics through the ordered-int trick ---------------------------------------------------------
OPR_DEV void atomicMinFloat(float* addr, float v)
{
    v += 0.0f;   // canonicalise -0
    if (v >= 0.f) atomicMin((int*)addr, __float_as_int(v)); else atomicMax((unsigned*)addr, __float_as_uint(v));
}
OPR_DEV void atomicMaxFloat(float* addr, float v)
{
    v += 0.0f;
    if (v >= 0.f) atomicMax((int*)addr, __float_as_int(v)); else atomicMin((unsigned*)addr, __float_as_uint(v));
}

constexpr int BBOX_THREADS = 256;

// One thread per photon SLOT (coalesced 16-byte reads; the first version gave a thread the <= 4 deposits of one emitting thread: strided
// reads in a dependent loop: 0.08 -> 0.05 ms including the hand-over to the host).  A photon is valid when it was stored this iteration and fmaxf(power) > 0
// (PhotonToAABBConverter :112-119).  min / max / count do not depend on the order: the result is bit-identical.
__global__ void __launch_bounds__(BBOX_THREADS) photon_bbox_kernel(PhotonBuf photons, const unsigned* __restrict__ phState, unsigned numThreads,
                                                                   float* __restrict__ bbox)
{
    __shared__ float sLo[BBOX_THREADS / 32][3], sHi[BBOX_THREADS / 32][3];
    __shared__ unsigned sCount[BBOX_THREADS / 32];
    float3 lo = f3(FLT_MAX), hi = f3(-FLT_MAX);
    unsigned count = 0;
    const size_t numSlots = (size_t)numThreads * MAX_PHOTON_DEPOSITS;
    for (size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x; s < numSlots; s += (size_t)gridDim.x * blockDim.x) {
        const unsigned numStored = (__ldg(phState + s / MAX_PHOTON_DEPOSITS) >> 8) & 0xffu;
        if ((unsigned)(s % MAX_PHOTON_DEPOSITS) < numStored) {
            const float4 b = photons.b[s];
            if (fmaxf(fmaxf(b.x, b.y), b.z) > 0) {
                const float4 a = photons.a[s];
                lo = fmin3(lo, f3(a)); hi = fmax3(hi, f3(a));
                count++;
            }
        }
    }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo.x = fminf(lo.x, __shfl_xor_sync(0xffffffffu, lo.x, o)); lo.y = fminf(lo.y, __shfl_xor_sync(0xffffffffu, lo.y, o));
        lo.z = fminf(lo.z, __shfl_xor_sync(0xffffffffu, lo.z, o));
        hi.x = fmaxf(hi.x, __shfl_xor_sync(0xffffffffu, hi.x, o)); hi.y = fmaxf(hi.y, __shfl_xor_sync(0xffffffffu, hi.y, o));
        hi.z = fmaxf(hi.z, __shfl_xor_sync(0xffffffffu, hi.z, o));
        count += __shfl_xor_sync(0xffffffffu, count, o);
    }
    const unsigned w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { sLo[w][0] = lo.x; sLo[w][1] = lo.y; sLo[w][2] = lo.z; sHi[w][0] = hi.x; sHi[w][1] = hi.y; sHi[w][2] = hi.z; sCount[w] = count; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (unsigned i = 1; i < BBOX_THREADS / 32; i++) {
            lo = fmin3(lo, f3(sLo[i][0], sLo[i][1], sLo[i][2])); hi = fmax3(hi, f3(sHi[i][0], sHi[i][1], sHi[i][2])); count += sCount[i];
        }
        if (count) {
            atomicMinFloat(bbox + 0, lo.x); atomicMinFloat(bbox + 1, lo.y); atomicMinFloat(bbox + 2, lo.z);
            atomicMaxFloat(bbox + 3, hi.x); atomicMaxFloat(bbox + 4, hi.y); atomicMaxFloat(bbox + 5, hi.z);
            atomicAdd((unsigned*)(bbox + 6), count);
        }
    }
}

// Hands the seven scalars to the host by storing them into mapped pinned memory, and re-arms the device copy for the next iteration.
// A store over PCIe instead of a cudaMemcpyAsync: the copy engine may be busy with a 25 MB image read-back (opr_get_output_buffer_async),
// and the host waits for exactly these 28 bytes before it can launch the grid build.
__global__ void photon_bbox_publish_kernel(float* __restrict__ bbox, volatile float* hostOut)
{
    const int i = threadIdx.x;
    if (i < 7) {
        hostOut[i] = bbox[i];
        bbox[i] = i < 3 ? FLT_MAX : (i < 6 ? -FLT_MAX : 0.f);
    }
    __threadfence_system();
}

void launch_photon_bbox(const PhotonBuf& photons, const unsigned* phState, unsigned numThreads, float* bbox, LaunchCtx lc, float* hostOutMapped)
{
    const size_t numSlots = (size_t)numThreads * MAX_PHOTON_DEPOSITS;
    int blocks = (int)std::min<size_t>((numSlots + BBOX_THREADS - 1) / BBOX_THREADS, (size_t)numSMs() * 8);
    photon_bbox_kernel<<<blocks, BBOX_THREADS, 0, lc.st>>>(photons, phState, numThreads, bbox);
    ++*lc.launches;
    if (hostOutMapped) { photon_bbox_publish_kernel<<<1, 32, 0, lc.st>>>(bbox, hostOutMapped); ++*lc.launches; }
}

// ---- host scalars: getSmallestPossibleCellSize :62-71, calculateGridSize :42-50, :229-248 ---------------
static inline unsigned f2u_sat(float f) { if (!(f > 0.0f)) return 0u; if (f >= 4294967296.0f) return 0xffffffffu; return (unsigned)f; }

bool compute_grid_params(const float lo[3], const float hi[3], GridParams& g)
{
    // padAABB :135-141
    const float first[3] = {lo[0] - 0.0000001f, lo[1] - 0.0000001f, lo[2] - 0.0000001f};
    const float second[3] = {hi[0] + 0.0000001f, hi[1] + 0.0000001f, hi[2] + 0.0000001f};
    const float ext[3] = {second[0] - first[0], second[1] - first[1], second[2] - first[2]};
    float sceneVolume = ext[0] * ext[1] * ext[2];
    float minVolumePerCell = sceneVolume / PHOTON_GRID_MAX_SIZE;
    float smallestPossibleRadiusC = powf(minVolumePerCell, 1.0f / 3.0f);
    float inv = 1.0f / smallestPossibleRadiusC;                           // float3 / float multiplies by the reciprocal
    unsigned nc[3] = {f2u_sat(floorf(ext[0] * inv)), f2u_sat(floorf(ext[1] * inv)), f2u_sat(floorf(ext[2] * inv))};
    float r0 = ext[0] / nc[0], r1 = ext[1] / nc[1], r2 = ext[2] / nc[2];
    float smallestPossibleCellSize = fmaxf(fmaxf(r0, r1), r2);
    float cellSize = (float)(smallestPossibleCellSize + 0.001);           // float + double literal (:230)
    float invc = 1.0f / cellSize;
    unsigned gs[3] = {f2u_sat(ceilf(ext[0] * invc)), f2u_sat(ceilf(ext[1] * invc)), f2u_sat(ceilf(ext[2] * invc))};
    for (int i = 0; i < 3; i++) gs[i] = gs[i] < 1u ? 1u : gs[i];
    g.cellSize = cellSize;
    g.gridSize = make_uint3(gs[0], gs[1], gs[2]);
    g.origo = make_float3(first[0], first[1], first[2]);
    g.numCells = gs[0] * gs[1] * gs[2];
    return g.numCells <= PHOTON_GRID_MAX_SIZE;                            // :244-248
}

// ---- keys + histogram: calculateHashCellsKernel :152-173, getPhotonGridIndex(1D) ppm/PhotonGrid.h:19-28 ---
__global__ void __launch_bounds__(256) grid_keys_hist_kernel(PhotonBuf photons, const unsigned* __restrict__ phState, unsigned numSlots,
                                                             const GridParams* __restrict__ gp, unsigned* __restrict__ keys, unsigned* __restrict__ hist)
{
    const GridParams g = *gp;       // written by the host function that precedes this kernel in the stream (renderer.cu)
    const float invCell = 1.f / g.cellSize;
    const unsigned invalidHashCell = g.numCells + 1;
    for (unsigned s = blockIdx.x * blockDim.x + threadIdx.x; s < numSlots; s += gridDim.x * blockDim.x) {
        const unsigned t = s / MAX_PHOTON_DEPOSITS, k = s % MAX_PHOTON_DEPOSITS;
        const unsigned numStored = (phState[t] >> 8) & 0xffu;
        unsigned hashCell = invalidHashCell;
        if (k < numStored) {
            const float4 b = photons.b[s];
            if (fmaxf(fmaxf(b.x, b.y), b.z) > 0) {
                const float4 a = photons.a[s];
                // (position - origo) * (1/cellSize): a subtraction then a multiplication, nothing to contract
                const float px = (a.x - g.origo.x) * invCell, py = (a.y - g.origo.y) * invCell, pz = (a.z - g.origo.z) * invCell;
                const unsigned gx = (unsigned)floorf(px), gy = (unsigned)floorf(py), gz = (unsigned)floorf(pz);
                hashCell = gx + gy * g.gridSize.x + gz * g.gridSize.x * g.gridSize.y;
                if (hashCell < g.numCells) atomicAdd(hist + hashCell, 1u);
            }
        }
        keys[s] = hashCell;
    }
}

void launch_grid_keys(const PhotonBuf& photons, const unsigned* phState, unsigned numSlots, const GridParams* g, unsigned* keys, unsigned* hist, LaunchCtx lc)
{
    int blocks = (int)std::min<unsigned>((numSlots + 255) / 256, numSMs() * 16);
    grid_keys_hist_kernel<<<blocks, 256, 0, lc.st>>>(photons, phState, numSlots, g, keys, hist);
    ++*lc.launches;
}

__global__ void __launch_bounds__(256) photon_reorder_kernel(PhotonBuf src, const unsigned* __restrict__ sortedSlots, const GridParams* __restrict__ gp, PhotonBuf dst)
{
    const unsigned numValid = gp->numValid;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < numValid; i += gridDim.x * blockDim.x) {
        const unsigned s = sortedSlots[i];
        dst.a[i] = src.a[s]; dst.b[i] = src.b[s]; dst.c[i] = src.c[s];
    }
}

void launch_photon_reorder(const PhotonBuf& src, const unsigned* sortedSlots, unsigned maxValid, const GridParams* g, const PhotonBuf& dst, LaunchCtx lc)
{
    if (!maxValid) return;
    int blocks = (int)std::min<unsigned>((maxValid + 255) / 256, numSMs() * 16);
    photon_reorder_kernel<<<blocks, 256, 0, lc.st>>>(src, sortedSlots, g, dst);
    ++*lc.launches;
}

// ---- gather: IndirectRadianceEstimation::kernel (:69-129, :213-220), validPhoton :54-57, photonPower :59-67
//
// One 128-thread block gathers for an 8x4 pixel tile; lane l of every warp owns pixel l of the tile and the four
// warps split the tile's cell rows between them.  The reference lets every pixel walk its own cube of cells and
// read each photon through the cache (13 of 32 lanes active, SMs idle half of the kernel: profiles/r1_*).  Here a
// warp walks rows of the UNION of its lanes' cubes in the reference's order (z, then y, then the contiguous x-run
// of sorted photons), stages 32 photons at a time in shared memory with coalesced 16-byte loads, and every lane
// that needs the row tests the staged photons against its own hit point (shared-memory broadcast reads).  A lane
// accumulates exactly the photons the reference accumulates; the four per-warp partial sums are added in a fixed
// order.  Tiles are handed out through an atomic counter so that photon-dense tiles do not leave SMs idle.
OPR_DEV float exp2f_approx(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

#ifndef OPR_GATHER_UNROLL
#define OPR_GATHER_UNROLL 8
#endif
#define OPR_PRAGMA_(x) _Pragma(#x)
#define OPR_PRAGMA(x) OPR_PRAGMA_(x)
#ifndef OPR_GATHER_BRANCHFREE
#define OPR_GATHER_BRANCHFREE 1
#endif
constexpr int GATHER_THREADS = 128;
constexpr int GATHER_WARPS = GATHER_THREADS / 32;

// Packed FP32 (sm_100: FADD2 / FMUL2 / FFMA2, PTX *.f32x2).  Measured on B200 (scripts/micro/ffma2.cu): FFMA2 runs at HALF the
// instruction rate of FFMA, i.e. the same 70 Tflop/s — the FP32 pipe gains nothing, the ISSUE slots do: one slot feeds two lanes'
// worth of FMAs.  The inner loop of the gather was bound by issue (83 % issue-active: 14 FP32-pipe instructions + 7 others per
// candidate photon); with two photons per packed instruction it needs ~13 slots per photon for the same 14 pipe cycles, so the
// FP32 pipe becomes the limit.  Results are those of the scalar IEEE operations, lane by lane (rn, no contraction beyond the
// fmaf chains the scalar loop already used).
#ifndef OPR_GATHER_PACKED
#define OPR_GATHER_PACKED 1
#endif
// w if (d2 <= r2 and nd <= 0) else 0, as a chained predicate and ONE select (the compiler emits a select per condition)
OPR_DEV float keep_if_both(float w, float d2, float r2, float nd)
{
    float r;
    asm("{ .reg .pred p;\n\t setp.le.f32 p, %2, %3;\n\t setp.le.and.f32 p, %4, 0f00000000, p;\n\t selp.f32 %0, %1, 0f00000000, p; }" : "=f"(r) : "f"(w), "f"(d2), "f"(r2), "f"(nd));
    return r;
}
typedef unsigned long long f32x2;
OPR_DEV f32x2 pk2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
OPR_DEV void upk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
OPR_DEV f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
OPR_DEV f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
OPR_DEV f32x2 sub2(f32x2 a, f32x2 b) { f32x2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

// 61 registers: 8 CTAs per SM stay resident, which lets the block scheduler interleave the gather with the path passes of the next
// iteration.  A 96-register build (5 CTAs per SM) is 2 % faster alone on the photon-dense bench scene and 10 % slower on the sparse one.
__global__ void __launch_bounds__(GATHER_THREADS, 8) ppm_gather_kernel(HitpointBuf hp, unsigned numPixels, unsigned width, unsigned height, PhotonBuf sorted,
                                                                    const unsigned* __restrict__ offsets, const GridParams* __restrict__ gp, float ppmRadius, float ppmRadiusSquared,
                                                                    float emittedPhotons, float* __restrict__ indirect, unsigned* __restrict__ tileCounter, unsigned tilesPerBlock)
{
#if OPR_GATHER_PACKED
    // staged photons as nine component arrays per warp (x y z | dir x y z | power r g b): a 128-bit shared load fetches one component
    // of FOUR photons as two ready-made packed pairs
    __shared__ __align__(16) float sP[GATHER_WARPS][9][32];
#else
    __shared__ float4 sA[GATHER_WARPS][32];
    __shared__ float4 sB[GATHER_WARPS][32];
    __shared__ float sC[GATHER_WARPS][32];
#endif
    __shared__ float sAcc[GATHER_WARPS][32][3];
    __shared__ unsigned sTile;
    const GridParams g = *gp;
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned tilesX = width ? (width + 7) / 8 : 0;
    const unsigned numTiles = width ? tilesX * ((height + 3) / 4) : (numPixels + 31) / 32;
    const float radius2 = ppmRadiusSquared, radius = ppmRadius;
    const float invCellSize = 1.f / g.cellSize;
    const unsigned gx = g.gridSize.x, gxy = g.gridSize.x * g.gridSize.y;
    // photonPower (:59-67): alpha*(1 - (1 - exp(-beta d2 / 2r2)) / (1 - e^-beta)) = w0 + w1*exp2(kexp*d2).  The two IEEE
    // divisions and expf per photon saturated the XU pipe in the first version; one ex2.approx remains — the reference
    // itself was built with -use_fast_math (RenderEngine.vcxproj:129).
    const float alpha = 1.818f, beta = 1.953f, expNegativeBeta = 0.141847f;
    const float w1 = alpha / (1 - expNegativeBeta), w0 = alpha - w1;
    const float kexp = -beta / (2 * radius2) * 1.44269504088896341f;
    // cull radius: a relative margin (plus the fp32 error of cell coordinates ~ extent * 2^-22) keeps the cull conservative
    const float cullRadius = radius * 1.0001f + 1e-6f * g.cellSize * (float)max(max(g.gridSize.x, g.gridSize.y), g.gridSize.z);
    const float cullRadius2 = cullRadius * cullRadius;
    for (unsigned done = 0; tilesPerBlock == 0 || done < tilesPerBlock; done++) {
        __syncthreads();
        if (threadIdx.x == 0) sTile = atomicAdd(tileCounter, 1u);
        __syncthreads();
        const unsigned tile = sTile;
        if (tile >= numTiles) break;
        unsigned p; bool inImage;
        if (width) {
            const unsigned x = (tile % tilesX) * 8 + (lane & 7), y = (tile / tilesX) * 4 + (lane >> 3);
            inImage = x < width && y < height; p = y * width + x;
        } else { p = tile * 32 + lane; inImage = p < numPixels; }
        float3 pos = f3(0.f), nrm = f3(0.f);
        bool active = false;
        if (inImage) {
            const float4 pos4 = hp.pos[p];
            if (__float_as_uint(pos4.w) & PRD_HIT_NON_SPECULAR) { active = true; pos = f3(pos4); nrm = f3(hp.nrm[p]); }
        }
        unsigned x_lo = 0xffffffffu, y_lo = 0xffffffffu, z_lo = 0xffffffffu, x_hi = 0, y_hi = 0, z_hi = 0;
        float3 npos = f3(0.f);
        if (active) {
            const float3 np = pos - g.origo;
            npos = np;
            x_lo = (unsigned)max(0, (int)((np.x - radius) * invCellSize));
            y_lo = (unsigned)max(0, (int)((np.y - radius) * invCellSize));
            z_lo = (unsigned)max(0, (int)((np.z - radius) * invCellSize));
            x_hi = min(g.gridSize.x - 1, (unsigned)((np.x + radius) * invCellSize));
            y_hi = min(g.gridSize.y - 1, (unsigned)((np.y + radius) * invCellSize));
            z_hi = min(g.gridSize.z - 1, (unsigned)((np.z + radius) * invCellSize));
            if (!(x_lo <= x_hi) || !(y_lo <= y_hi) || !(z_lo <= z_hi)) active = false;
        }
        float3 acc = f3(0.f);
#if OPR_GATHER_PACKED
        const f32x2 posX2 = pk2(pos.x, pos.x), posY2 = pk2(pos.y, pos.y), posZ2 = pk2(pos.z, pos.z);
        const f32x2 nrmX2 = pk2(nrm.x, nrm.x), nrmY2 = pk2(nrm.y, nrm.y), nrmZ2 = pk2(nrm.z, nrm.z);
        const f32x2 kexp2 = pk2(kexp, kexp), w12 = pk2(w1, w1), w02 = pk2(w0, w0);
        f32x2 accX2 = pk2(0.f, 0.f), accY2 = accX2, accZ2 = accX2;      // even / odd photons of a row chunk
#endif
        if (__any_sync(0xffffffffu, active)) {
            const unsigned uy_lo = __reduce_min_sync(0xffffffffu, active ? y_lo : 0xffffffffu), uy_hi = __reduce_max_sync(0xffffffffu, active ? y_hi : 0u);
            const unsigned uz_lo = __reduce_min_sync(0xffffffffu, active ? z_lo : 0xffffffffu), uz_hi = __reduce_max_sync(0xffffffffu, active ? z_hi : 0u);
            const unsigned ny = uy_hi - uy_lo + 1;
            // Rows in (z, y) order, dealt to the four warps by (z + y) mod 4.  Photons lie on surfaces, so the rows that hold them
            // share one y (floor, ceiling) or one z (walls); row-index interleaving put all of them on the same warp whenever the
            // cube was a multiple of four rows high (20 % of all stall samples were at the block barrier: profiles/r1k).
            const unsigned nz = uz_hi - uz_lo + 1;
            for (unsigned rz = 0; rz < nz; rz++)
            for (unsigned ry = (w + GATHER_WARPS - (rz & (GATHER_WARPS - 1))) & (GATHER_WARPS - 1); ry < ny; ry += GATHER_WARPS) {
                const unsigned z = uz_lo + rz, y = uy_lo + ry;
                bool need = active && z >= z_lo && z <= z_hi && y >= y_lo && y <= y_hi;   // the rows the reference visits for this lane
                // Sphere cull (results unchanged): the reference walks the whole cube of cells and rejects photons one by one with
                // distance2 <= radius2; cells of this row that lie entirely outside the (slightly inflated) sphere cannot hold an
                // accepted photon, so the lane's x-run shrinks to the chord of the sphere at this (y, z) slab.
                unsigned lx_lo = x_lo, lx_hi = x_hi;
                if (need) {
                    const float cy0 = (float)y * g.cellSize, cz0 = (float)z * g.cellSize;
                    const float dy = fmaxf(fmaxf(cy0 - npos.y, npos.y - (cy0 + g.cellSize)), 0.f);
                    const float dz = fmaxf(fmaxf(cz0 - npos.z, npos.z - (cz0 + g.cellSize)), 0.f);
                    const float rem = cullRadius2 - dy * dy - dz * dz;
                    if (rem < 0.f) need = false;
                    else {
                        const float hw = sqrtf(rem);
                        lx_lo = max(x_lo, (unsigned)max(0, (int)((npos.x - hw) * invCellSize)));
                        lx_hi = min(x_hi, (unsigned)fmaxf((npos.x + hw) * invCellSize, 0.f));
                        if (lx_lo > lx_hi) need = false;
                    }
                }
                if (!__any_sync(0xffffffffu, need)) continue;
                const unsigned rx_lo = __reduce_min_sync(0xffffffffu, need ? lx_lo : 0xffffffffu), rx_hi = __reduce_max_sync(0xffffffffu, need ? lx_hi : 0u);
                const unsigned from = rx_lo + y * gx + z * gxy;
                const unsigned begin = __ldg(offsets + from), end = __ldg(offsets + from + (rx_hi - rx_lo) + 1);
                // software pipeline: the next 32 photons of the row are in flight (registers) while the current 32 are tested from
                // shared memory
                float4 ra = make_float4(0.f, 0.f, 0.f, 0.f), rb = ra; float rc = 0.f;
                if (begin + lane < end) { ra = __ldg(sorted.a + begin + lane); rb = __ldg(sorted.b + begin + lane); rc = __ldg(sorted.c + begin + lane); }
                for (unsigned base = begin; base < end; base += 32) {
#if OPR_GATHER_PACKED
                    {   // lanes beyond the row's end stage a photon that no pixel accepts (distance2 = inf, zero power)
                        const bool real = base + lane < end;
                        sP[w][0][lane] = real ? ra.x : 1e30f; sP[w][1][lane] = ra.y; sP[w][2][lane] = ra.z;
                        sP[w][3][lane] = ra.w; sP[w][4][lane] = rb.w; sP[w][5][lane] = rc;
                        sP[w][6][lane] = real ? rb.x : 0.f; sP[w][7][lane] = real ? rb.y : 0.f; sP[w][8][lane] = real ? rb.z : 0.f;
                    }
#else
                    sA[w][lane] = ra; sB[w][lane] = rb; sC[w][lane] = rc;
#endif
                    __syncwarp();
                    const unsigned nxt = base + 32 + lane;
                    if (nxt < end) { ra = __ldg(sorted.a + nxt); rb = __ldg(sorted.b + nxt); rc = __ldg(sorted.c + nxt); }
                    else { ra = make_float4(0.f, 0.f, 0.f, 0.f); rb = ra; rc = 0.f; }
                    const unsigned cnt = min(32u, end - base);
                    if (need) {
#if OPR_GATHER_PACKED
                        // four photons per trip, two per packed instruction; same operations per photon as the scalar loop below
                        OPR_PRAGMA(unroll 2)
                        for (unsigned j = 0; j < cnt; j += 4) {
                            const ulonglong2 X = *(const ulonglong2*)&sP[w][0][j], Y = *(const ulonglong2*)&sP[w][1][j], Z = *(const ulonglong2*)&sP[w][2][j];
                            const ulonglong2 DX = *(const ulonglong2*)&sP[w][3][j], DY = *(const ulonglong2*)&sP[w][4][j], DZ = *(const ulonglong2*)&sP[w][5][j];
                            const ulonglong2 PR = *(const ulonglong2*)&sP[w][6][j], PG = *(const ulonglong2*)&sP[w][7][j], PB = *(const ulonglong2*)&sP[w][8][j];
                            #pragma unroll
                            for (int h = 0; h < 2; h++) {
                                const f32x2 x = h ? X.y : X.x, y = h ? Y.y : Y.x, z = h ? Z.y : Z.x;
                                const f32x2 dx = sub2(posX2, x), dy = sub2(posY2, y), dz = sub2(posZ2, z);
                                // dot(a, b) as nvcc compiles the reference's a.x b.x + a.y b.y + a.z b.z for the GPU: one product and two FMAs
                                // (IndirectRadianceEstimation.cu:54-57,117-118).  The CPU oracle rounds every product, so on a radius sphere or
                                // tangent plane the last bit of distance2 / the dot decides differently: 2 of 2 073 600 pixels at the full
                                // configuration hold such a photon, and the stage check names them (tests/parity_rules.py gather_stage_exact).
                                // (Rounded products are not available at this price: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2
                                // whatever -fmad says; packed products with scalar adds cost four more instructions per photon.)
                                const f32x2 d2 = fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));
                                const f32x2 nd = fma2(h ? DZ.y : DZ.x, nrmZ2, fma2(h ? DY.y : DY.x, nrmY2, mul2(h ? DX.y : DX.x, nrmX2)));
                                float e0, e1, d20, d21, nd0, nd1;
                                upk2(mul2(kexp2, d2), e0, e1); upk2(d2, d20, d21); upk2(nd, nd0, nd1);
                                const f32x2 wgt = fma2(w12, pk2(exp2f_approx(e0), exp2f_approx(e1)), w02);
                                float wg0, wg1;
                                upk2(wgt, wg0, wg1);
                                wg0 = keep_if_both(wg0, d20, radius2, nd0);
                                wg1 = keep_if_both(wg1, d21, radius2, nd1);
                                const f32x2 wsel = pk2(wg0, wg1);
                                accX2 = fma2(h ? PR.y : PR.x, wsel, accX2); accY2 = fma2(h ? PG.y : PG.x, wsel, accY2); accZ2 = fma2(h ? PB.y : PB.x, wsel, accZ2);
                            }
                        }
#elif OPR_GATHER_BRANCHFREE
                        // No branch per photon: both tests are selects, a rejected photon adds an exact zero.  21 instructions per candidate
                        // for every lane instead of 9 + 12 for the ~55 % that pass the distance test — the same issue count once any lane of
                        // the warp accepts, without the divergence bookkeeping and with four candidates in flight (gather 1.94 -> 1.84 ms).
                        OPR_PRAGMA(unroll OPR_GATHER_UNROLL)
                        for (unsigned j = 0; j < cnt; j++) {
                            const float4 a = sA[w][j], b = sB[w][j];
                            const float c = sC[w][j];
                            const float dx = pos.x - a.x, dy = pos.y - a.y, dz = pos.z - a.z;
                            const float distance2 = (dx * dx + dy * dy) + dz * dz;            // the reference's dot(): see the packed loop
                            const float ndot = (a.w * nrm.x + b.w * nrm.y) + c * nrm.z;
                            float weight = fmaf(w1, exp2f_approx(kexp * distance2), w0);
                            weight = (distance2 <= radius2 && ndot <= 0.f) ? weight : 0.f;
                            acc.x = fmaf(b.x, weight, acc.x); acc.y = fmaf(b.y, weight, acc.y); acc.z = fmaf(b.z, weight, acc.z);
                        }
#else
                        // One branch per photon (the distance test, which rejects ~45 %); the direction test is a select, so an accepted
                        // photon is a straight run of loads and FMAs with no nested divergence region.
                        #pragma unroll 4
                        for (unsigned j = 0; j < cnt; j++) {
                            const float4 a = sA[w][j];
                            const float dx = pos.x - a.x, dy = pos.y - a.y, dz = pos.z - a.z;
                            const float distance2 = (dx * dx + dy * dy) + dz * dz;
                            if (distance2 <= radius2) {
                                const float4 b = sB[w][j];
                                const float c = sC[w][j];
                                const float ndot = (a.w * nrm.x + b.w * nrm.y) + c * nrm.z;           // dot(-rayDirection, normal) >= 0  <=>  ndot <= 0
                                float weight = fmaf(w1, exp2f_approx(kexp * distance2), w0);
                                weight = ndot <= 0.f ? weight : 0.f;                                   // a rejected photon adds an exact zero
                                acc.x = fmaf(b.x, weight, acc.x); acc.y = fmaf(b.y, weight, acc.y); acc.z = fmaf(b.z, weight, acc.z);
                            }
                        }
#endif
                    }
                    __syncwarp();
                }
            }
        }
#if OPR_GATHER_PACKED
        { float lo, hi; upk2(accX2, lo, hi); acc.x = lo + hi; upk2(accY2, lo, hi); acc.y = lo + hi; upk2(accZ2, lo, hi); acc.z = lo + hi; }
#endif
        sAcc[w][lane][0] = acc.x; sAcc[w][lane][1] = acc.y; sAcc[w][lane][2] = acc.z;
        __syncthreads();
        if (w == 0 && inImage) {
            float3 sum = f3(0.f);
            #pragma unroll
            for (int ww = 0; ww < GATHER_WARPS; ww++) { sum.x += sAcc[ww][lane][0]; sum.y += sAcc[ww][lane][1]; sum.z += sAcc[ww][lane][2]; }
            const float3 attenuation = f3(hp.att[p]);
            const float3 ind = sum * attenuation * (1.0f / (PI_F * ppmRadiusSquared)) * (1.0f / emittedPhotons);
            indirect[3 * (size_t)p] = ind.x; indirect[3 * (size_t)p + 1] = ind.y; indirect[3 * (size_t)p + 2] = ind.z;
        }
    }
}

void launch_gather(const HitpointBuf& hp, unsigned numPixels, unsigned width, unsigned height, const PhotonBuf& sorted, const unsigned* offsets, const GridParams* g,
                   float ppmRadius, float ppmRadiusSquared, float emittedPhotons, float* indirect, unsigned* tileCounter, LaunchCtx lc, unsigned blocksPerSM,
                   unsigned tilesPerBlock)
{
    const unsigned numTiles = width ? ((width + 7) / 8) * ((height + 3) / 4) : (numPixels + 31) / 32;
    int blocks = (int)std::min<unsigned>(numTiles, numSMs() * std::max(1u, blocksPerSM));
    if (tilesPerBlock) blocks = (int)((numTiles + tilesPerBlock - 1) / tilesPerBlock);   // short-lived CTAs: as many resident as fit
    OPR_CUDA_CHECK(cudaMemsetAsync(tileCounter, 0, sizeof(unsigned), lc.st));
    ppm_gather_kernel<<<blocks, GATHER_THREADS, 0, lc.st>>>(hp, numPixels, width, height, sorted, offsets, g, ppmRadius, ppmRadiusSquared, emittedPhotons, indirect,
                                                            tileCounter, tilesPerBlock);
    ++*lc.launches;
}

} // namespace opr
