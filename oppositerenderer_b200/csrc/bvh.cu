// bvh.cu — GPU build of the acceleration structure that replaces the OptiX "Trbvh"/"Bvh" builders the reference
// requests (scene/Scene.cpp:353,513,547; scene/Cornell.cpp:183,188) and the layout the traversal in scene.cuh walks.
//
//   1. Morton codes of the primitive-box centroids (63 bit) + stable radix sort            (bvh_morton_kernel)
//   2. PLOC — parallel locally-ordered clustering: every cluster looks R neighbours left and right in Morton order
//      for the partner that minimises the surface area of the merged box; mutual nearest neighbours merge.  This is
//      agglomerative SAH clustering, so the binary tree is of treelet-optimised (Trbvh-class) quality, not LBVH's.
//                                                                     (ploc_nn_kernel, ploc_merge_kernel, ploc_compact_kernel)
//   3. SAH leaf decision, bottom-up inside the merge: a subtree of <= maxLeaf primitives becomes one leaf when
//      Ci*count*A <= Ct*A + cost(children).
//   4. Collapse to 4-wide nodes, top-down, one kernel launch per level: a wide node adopts the two children of its
//      binary node and keeps replacing the child with the largest area by its own two children until it has four.
//      Leaves get consecutive slots in a leaf-ordered primitive list.                              (bvh_collapse_kernel)
//   5. Triangle records are copied into leaf-slot order (padded to 64 B) so that a leaf is one contiguous run of records.
//                                                                                                   (bvh_leaf_tri_kernel)
//
// Wide node, 128 bytes = one cache line = eight float4:
//   [0] lo.x of children 0..3   [1] hi.x   [2] lo.y   [3] hi.y   [4] lo.z   [5] hi.z   [6] child refs (int4)   [7] spare
//   child ref >= 0: wide node index;  < 0: leaf, ~ref = (firstSlot << 4) | ((count-1) << 1) | mixed;  empty: box = +inf.
#include "device_scene.h"
#include "radix_sort.cuh"
#include <vector>
#include <cstdlib>
#include <cstring>
#include <cstdio>

namespace opr {

__device__ __forceinline__ unsigned long long expandBits21(unsigned v)
{
    unsigned long long x = v & 0x1fffffu;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

__global__ void bvh_morton_kernel(const float4* __restrict__ lo, const float4* __restrict__ hi, unsigned n, float3 cmin, float3 cinv,
                                  unsigned long long* __restrict__ keys)
{
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 a = lo[i], b = hi[i];
    float cx = ((a.x + b.x) * 0.5f - cmin.x) * cinv.x;
    float cy = ((a.y + b.y) * 0.5f - cmin.y) * cinv.y;
    float cz = ((a.z + b.z) * 0.5f - cmin.z) * cinv.z;
    unsigned ix = (unsigned)fminf(fmaxf(cx * 2097152.0f, 0.0f), 2097151.0f);
    unsigned iy = (unsigned)fminf(fmaxf(cy * 2097152.0f, 0.0f), 2097151.0f);
    unsigned iz = (unsigned)fminf(fmaxf(cz * 2097152.0f, 0.0f), 2097151.0f);
    keys[i] = (expandBits21(ix) << 2) | (expandBits21(iy) << 1) | expandBits21(iz);
}

// Binary-tree arrays.  Node ids: [0, n) leaves in Morton order, [n, 2n-1) internal nodes in creation order.
struct Bvh2 {
    float4* lo; float4* hi;       // boxes; lo.w of a leaf = bits of ~((primIndex << 2) | primType)
    int* left; int* right;
    unsigned* count;              // primitives below
    float* cost;                  // SAH cost of the subtree with the leaf decision applied
    unsigned char* isLeaf;        // internal node collapsed into one leaf
};

__device__ __forceinline__ float half_area(const float4& lo, const float4& hi)
{
    const float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z;
    return dx * dy + dy * dz + dz * dx;
}

__global__ void bvh_init_leaves_kernel(const unsigned* __restrict__ sortedPrim, const float4* __restrict__ primLo, const float4* __restrict__ primHi, int n,
                                       Bvh2 T, int* __restrict__ cid, float ci)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned p = sortedPrim[i];
    const float4 lo = primLo[p], hi = primHi[p];
    T.lo[i] = lo; T.hi[i] = hi;
    T.count[i] = 1u;
    T.cost[i] = ci * half_area(lo, hi);
    T.isLeaf[i] = 1;
    cid[i] = i;
}

constexpr int PLOC_T = 256;
constexpr int PLOC_RMAX = 32;

// Nearest neighbour of every cluster among the R clusters before and after it.  The candidate order is the strict
// total order (merged area, min index, max index), so the globally best pair is always mutual and every round merges.
__global__ void __launch_bounds__(PLOC_T) ploc_nn_kernel(const int* __restrict__ cid, int m, int R, Bvh2 T, int* __restrict__ neighbor)
{
    __shared__ float4 sLo[PLOC_T + 2 * PLOC_RMAX], sHi[PLOC_T + 2 * PLOC_RMAX];
    const int base = blockIdx.x * PLOC_T - R;
    for (int k = threadIdx.x; k < PLOC_T + 2 * R; k += PLOC_T) {
        const int g = base + k;
        if (g >= 0 && g < m) { const int id = cid[g]; sLo[k] = T.lo[id]; sHi[k] = T.hi[id]; }
    }
    __syncthreads();
    const int i = blockIdx.x * PLOC_T + threadIdx.x;
    if (i >= m) return;
    const int li = threadIdx.x + R;
    const float4 alo = sLo[li], ahi = sHi[li];
    float best = FLT_MAX; int bj = -1;
    for (int k = -R; k <= R; k++) {
        const int g = i + k;
        if (k == 0 || g < 0 || g >= m) continue;
        const float4 blo = sLo[li + k], bhi = sHi[li + k];
        float4 ulo, uhi;
        ulo.x = fminf(alo.x, blo.x); ulo.y = fminf(alo.y, blo.y); ulo.z = fminf(alo.z, blo.z);
        uhi.x = fmaxf(ahi.x, bhi.x); uhi.y = fmaxf(ahi.y, bhi.y); uhi.z = fmaxf(ahi.z, bhi.z);
        const float a = half_area(ulo, uhi);
        bool better = a < best;
        if (a == best && bj >= 0) {
            const int p0 = min(i, g), p1 = max(i, g), q0 = min(i, bj), q1 = max(i, bj);
            better = p0 < q0 || (p0 == q0 && p1 < q1);
        }
        if (better) { best = a; bj = g; }
    }
    neighbor[i] = bj;
}

__global__ void __launch_bounds__(PLOC_T) ploc_merge_kernel(const int* __restrict__ cidIn, int m, const int* __restrict__ neighbor, Bvh2 T, int* __restrict__ nodeCounter,
                                                            int* __restrict__ cidTmp, unsigned* __restrict__ valid, float ci, float ct, unsigned maxLeaf)
{
    const int i = blockIdx.x * PLOC_T + threadIdx.x;
    if (i == 0) valid[m] = 0u;
    if (i >= m) return;
    const int j = neighbor[i];
    if (j >= 0 && neighbor[j] == i) {
        if (i < j) {
            const int id = atomicAdd(nodeCounter, 1);
            const int a = cidIn[i], b = cidIn[j];
            const float4 alo = T.lo[a], ahi = T.hi[a], blo = T.lo[b], bhi = T.hi[b];
            float4 lo = make_float4(fminf(alo.x, blo.x), fminf(alo.y, blo.y), fminf(alo.z, blo.z), 0.f);
            float4 hi = make_float4(fmaxf(ahi.x, bhi.x), fmaxf(ahi.y, bhi.y), fmaxf(ahi.z, bhi.z), 0.f);
            T.lo[id] = lo; T.hi[id] = hi; T.left[id] = a; T.right[id] = b;
            const unsigned cnt = T.count[a] + T.count[b];
            T.count[id] = cnt;
            const float A = half_area(lo, hi);
            const float cInt = ct * A + T.cost[a] + T.cost[b];
            const float cLeaf = cnt <= maxLeaf ? ci * (float)cnt * A : FLT_MAX;
            T.cost[id] = fminf(cInt, cLeaf);
            T.isLeaf[id] = cLeaf <= cInt ? 1 : 0;
            cidTmp[i] = id; valid[i] = 1u;
        } else {
            valid[i] = 0u;
        }
    } else {
        cidTmp[i] = cidIn[i]; valid[i] = 1u;
    }
}

__global__ void __launch_bounds__(PLOC_T) ploc_compact_kernel(const int* __restrict__ cidTmp, const unsigned* __restrict__ valid, const unsigned* __restrict__ pos, int m,
                                                              int* __restrict__ cidOut)
{
    const int i = blockIdx.x * PLOC_T + threadIdx.x;
    if (i < m && valid[i]) cidOut[pos[i]] = cidTmp[i];
}

// ---- collapse to 4-wide nodes ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) bvh_collapse_kernel(const int2* __restrict__ inQueue, int numIn, Bvh2 T, int n, float4* __restrict__ wide,
                                                           int* __restrict__ counters /* [0] wide nodes, [1] next queue, [2] leaf slots */,
                                                           int2* __restrict__ outQueue, int* __restrict__ leafPrim)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= numIn) return;
    const int2 item = inQueue[q];
    int ch[4]; int nch;
    auto leafLike = [&](int c) { return c < n || T.isLeaf[c]; };
    if (leafLike(item.x)) { ch[0] = item.x; nch = 1; }
    else {
        ch[0] = T.left[item.x]; ch[1] = T.right[item.x]; nch = 2;
        while (nch < 4) {
            int best = -1; float bestA = -1.f;
            for (int c = 0; c < nch; c++) {
                if (leafLike(ch[c])) continue;
                const float a = half_area(T.lo[ch[c]], T.hi[ch[c]]);
                if (a > bestA) { bestA = a; best = c; }
            }
            if (best < 0) break;
            const int node = ch[best];
            ch[best] = T.left[node]; ch[nch++] = T.right[node];
        }
    }
    const float INF = __int_as_float(0x7f800000);
    float lox[4], hix[4], loy[4], hiy[4], loz[4], hiz[4]; int ref[4];
    for (int c = 0; c < 4; c++) { lox[c] = hix[c] = loy[c] = hiy[c] = loz[c] = hiz[c] = INF; ref[c] = 0x7fffffff; }
    for (int c = 0; c < nch; c++) {
        const int node = ch[c];
        const float4 lo = T.lo[node], hi = T.hi[node];
        lox[c] = lo.x; hix[c] = hi.x; loy[c] = lo.y; hiy[c] = hi.y; loz[c] = lo.z; hiz[c] = hi.z;
        if (leafLike(node)) {
            const int cnt = (int)T.count[node];
            const int first = atomicAdd(&counters[2], cnt);
            int stack[16]; int sp = 0; int cur = node; int k = 0; int mixed = 0;
            for (;;) {   // walk the <= maxLeaf-primitive subtree, left to right
                if (cur < n) {
                    const int prim = ~__float_as_int(T.lo[cur].w);
                    if ((prim & 3) != PRIM_TRIANGLE) mixed = 1;
                    leafPrim[first + k++] = prim;
                    if (sp == 0) break;
                    cur = stack[--sp];
                } else { stack[sp++] = T.right[cur]; cur = T.left[cur]; }
            }
            ref[c] = ~((first << 4) | ((cnt - 1) << 1) | mixed);
        } else {
            const int w = atomicAdd(&counters[0], 1);
            const int slot = atomicAdd(&counters[1], 1);
            outQueue[slot] = make_int2(node, w);
            ref[c] = w;
        }
    }
    float4* N = wide + 8 * (size_t)item.y;
    N[0] = make_float4(lox[0], lox[1], lox[2], lox[3]); N[1] = make_float4(hix[0], hix[1], hix[2], hix[3]);
    N[2] = make_float4(loy[0], loy[1], loy[2], loy[3]); N[3] = make_float4(hiy[0], hiy[1], hiy[2], hiy[3]);
    N[4] = make_float4(loz[0], loz[1], loz[2], loz[3]); N[5] = make_float4(hiz[0], hiz[1], hiz[2], hiz[3]);
    N[6] = make_float4(__int_as_float(ref[0]), __int_as_float(ref[1]), __int_as_float(ref[2]), __int_as_float(ref[3]));
    N[7] = make_float4(0.f, 0.f, 0.f, 0.f);
}

__global__ void bvh_leaf_tri_kernel(const int* __restrict__ leafPrim, int n, const float4* __restrict__ triRec, float4* __restrict__ leafTri)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int prim = leafPrim[s];
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a, c = a;
    if ((prim & 3) == PRIM_TRIANGLE) { const int idx = prim >> 2; a = triRec[3 * idx]; b = triRec[3 * idx + 1]; c = triRec[3 * idx + 2]; }
    float4* dst = leafTri + LEAF_TRI_STRIDE * (size_t)s;
    dst[0] = a; dst[1] = b; dst[2] = c; dst[3] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// ---- final node formats ---------------------------------------------------------------------------------------------------------
// Both traversal formats store a child box as CENTRE and HALF-EXTENT per axis, not as lo / hi planes: the slab test then is
//   tc = c * (1/d) - o/d,  t_near = tc - h * |1/d|,  t_far = tc + h * |1/d|
// — three FFMAs per axis and child on the FP32 pipe (|x| and -x are operand modifiers) instead of two FFMAs and a min + a max on the ALU
// pipe, which runs at half rate and is what bounds the traversal kernels (ncu: ALU pipe 65 % active, FMA pipe 19 %; profiles/).
// The boxes stay conservative: h is rounded up so that [c - h, c + h] contains [lo, hi].
//
// Quantised copy: 64 bytes = 16 words: [0,1] centre.x of children 0..3 as 16-bit values, [2,3] half.x, [4,5] centre.y, [6,7] half.y,
// [8,9] centre.z, [10,11] half.z, [12..15] child refs.  A value is 0x8000 | q with q in [0, 32767] in units of extent / 32768 on the
// scene grid.  lo is rounded down and hi up, each by one further cell (covers the fp32 error of the decode and of the per-ray
// constants, 2^-22 of the extent), then widened by at most one more cell so that hi - lo is even and the centre lies on the grid.
__device__ __forceinline__ int quant_lo(float v, float origin, float invCell)
{
    return (int)fminf(fmaxf(floorf((v - origin) * invCell) - 1.0f, 0.0f), 32767.0f);
}
__device__ __forceinline__ int quant_hi(float v, float origin, float invCell)
{
    return (int)fminf(fmaxf(ceilf((v - origin) * invCell) + 1.0f, 0.0f), 32767.0f);
}
__device__ __forceinline__ void quant_box(float lo, float hi, float origin, float invCell, unsigned& c, unsigned& h)
{
    const float INF = __int_as_float(0x7f800000);
    if (!(lo < INF)) { c = 0x8000u; h = 0x8000u; return; }       // absent child: recognised by its ref in the traversal
    int ql = quant_lo(lo, origin, invCell), qh = quant_hi(hi, origin, invCell);
    if ((qh - ql) & 1) { if (qh < 32767) qh++; else ql--; }      // the grid frames the root box with a margin of 8 cells: ql >= 1 here
    c = 0x8000u | (unsigned)((ql + qh) >> 1);
    h = 0x8000u | (unsigned)((qh - ql) >> 1);
}
__global__ void bvh_quantize_kernel(const float4* __restrict__ wide, int numWide, float3 origin, float3 invCell, uint4* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= numWide) return;
    const float4* N = wide + 8 * (size_t)i;
    const float4 lox = N[0], hix = N[1], loy = N[2], hiy = N[3], loz = N[4], hiz = N[5], rf = N[6];
    auto axis = [&](const float4& lo, const float4& hi, float o, float ic) {
        unsigned c[4], h[4];
        quant_box(lo.x, hi.x, o, ic, c[0], h[0]); quant_box(lo.y, hi.y, o, ic, c[1], h[1]);
        quant_box(lo.z, hi.z, o, ic, c[2], h[2]); quant_box(lo.w, hi.w, o, ic, c[3], h[3]);
        return make_uint4(c[0] | (c[1] << 16), c[2] | (c[3] << 16), h[0] | (h[1] << 16), h[2] | (h[3] << 16));
    };
    uint4* Q = out + 4 * (size_t)i;
    Q[0] = axis(lox, hix, origin.x, invCell.x);
    Q[1] = axis(loy, hiy, origin.y, invCell.y);
    Q[2] = axis(loz, hiz, origin.z, invCell.z);
    Q[3] = make_uint4(__float_as_uint(rf.x), __float_as_uint(rf.y), __float_as_uint(rf.z), __float_as_uint(rf.w));
}

// fp32 nodes, in place: (lo4, hi4) per axis -> (centre4, half4).  An absent child (lo = +inf) becomes c = 0, h = -inf: t_near = +inf,
// t_far = -inf for every ray (|1/d| is finite and non-zero: safe_inverse_direction).
__device__ __forceinline__ void centre_half(float lo, float hi, float& c, float& h)
{
    const float INF = __int_as_float(0x7f800000);
    if (!(lo < INF)) { c = 0.f; h = -INF; return; }
    c = 0.5f * lo + 0.5f * hi;
    const float e = fmaxf(hi - c, c - lo);                       // each difference is within half an ulp of the true distance to c
    h = e + (fabsf(c) + fabsf(e)) * 2.4e-7f + 1e-37f;            // two ulps of the larger operand: covers the rounding of c, e and of c +- h
}
__global__ void bvh_centre_half_kernel(float4* __restrict__ wide, int numWide)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= numWide) return;
    float4* N = wide + 8 * (size_t)i;
    for (int a = 0; a < 3; a++) {
        const float4 lo = N[2 * a], hi = N[2 * a + 1];
        float4 c, h;
        centre_half(lo.x, hi.x, c.x, h.x); centre_half(lo.y, hi.y, c.y, h.y); centre_half(lo.z, hi.z, c.z, h.z); centre_half(lo.w, hi.w, c.w, h.w);
        N[2 * a] = c; N[2 * a + 1] = h;
    }
}

// ---- host driver -----------------------------------------------------------------------------------------------------
static int env_int(const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; }
static float env_float(const char* name, float dflt) { const char* e = getenv(name); return e ? (float)atof(e) : dflt; }

void DeviceScene::buildBvh(const std::vector<float4>& primLo, const std::vector<float4>& primHi, cudaStream_t st)
{
    const int n = (int)primLo.size();
    const int R = std::min(std::max(env_int("OPR_PLOC_RADIUS", 16), 1), PLOC_RMAX);
    const unsigned maxLeaf = (unsigned)std::min(std::max(env_int("OPR_BVH_MAXLEAF", 4), 1), 8);
    const float ci = 1.0f, ct = env_float("OPR_BVH_CT", 0.8f);

    // centroid bounds on the host (n is known here and this is a one-off per scene)
    float3 cmin = f3(FLT_MAX), cmax = f3(-FLT_MAX);
    for (int i = 0; i < n; i++) {
        float3 c = f3((primLo[i].x + primHi[i].x) * 0.5f, (primLo[i].y + primHi[i].y) * 0.5f, (primLo[i].z + primHi[i].z) * 0.5f);
        cmin = fmin3(cmin, c); cmax = fmax3(cmax, c);
    }
    float3 ext = cmax - cmin;
    float3 cinv = f3(ext.x > 0 ? 1.0f / ext.x : 0.f, ext.y > 0 ? 1.0f / ext.y : 0.f, ext.z > 0 ? 1.0f / ext.z : 0.f);

    std::vector<void*> tmp;
    auto dalloc = [&](size_t bytes) { void* p; devAlloc(&p, bytes); tmp.push_back(p); return p; };
    struct Cleanup { std::vector<void*>& v; ~Cleanup() { for (void* p : v) devFree(p); } } cleanup{tmp};

    float4* dLo = (float4*)dalloc(sizeof(float4) * n);
    float4* dHi = (float4*)dalloc(sizeof(float4) * n);
    OPR_CUDA_CHECK(cudaMemcpyAsync(dLo, primLo.data(), sizeof(float4) * n, cudaMemcpyHostToDevice, st));
    OPR_CUDA_CHECK(cudaMemcpyAsync(dHi, primHi.data(), sizeof(float4) * n, cudaMemcpyHostToDevice, st));
    unsigned long long* keys = (unsigned long long*)dalloc(8 * (size_t)n);
    unsigned long long* keysAlt = (unsigned long long*)dalloc(8 * (size_t)n);
    unsigned* vals = (unsigned*)dalloc(4 * (size_t)n);
    unsigned* valsAlt = (unsigned*)dalloc(4 * (size_t)n);
    unsigned* hist = (unsigned*)dalloc(4 * rs_hist_words(n));
    const int B256 = (n + 255) / 256;
    bvh_morton_kernel<<<B256, 256, 0, st>>>(dLo, dHi, n, cmin, cinv, keys);
    const bool flipped = radix_sort_pairs<unsigned long long>(keys, vals, keysAlt, valsAlt, n, 64, hist, true, st);
    const unsigned* sortedPrim = flipped ? valsAlt : vals;

    // ---- PLOC
    const size_t nn = 2 * (size_t)n;
    Bvh2 T;
    T.lo = (float4*)dalloc(sizeof(float4) * nn); T.hi = (float4*)dalloc(sizeof(float4) * nn);
    T.left = (int*)dalloc(4 * nn); T.right = (int*)dalloc(4 * nn);
    T.count = (unsigned*)dalloc(4 * nn); T.cost = (float*)dalloc(4 * nn); T.isLeaf = (unsigned char*)dalloc(nn);
    int* cid[2] = {(int*)dalloc(4 * (size_t)n), (int*)dalloc(4 * (size_t)n)};
    int* cidTmp = (int*)dalloc(4 * (size_t)n);
    int* neighbor = (int*)dalloc(4 * (size_t)n);
    unsigned* valid = (unsigned*)dalloc(4 * ((size_t)n + 1));
    unsigned* pos = (unsigned*)dalloc(4 * ((size_t)n + 1));
    unsigned* scanTmp = (unsigned*)dalloc(4 * scan_tmp_words(n + 1));
    int* counters = (int*)dalloc(16);
    {
        int h[4] = {n, 0, 0, 0};   // next internal node id
        OPR_CUDA_CHECK(cudaMemcpyAsync(counters, h, sizeof h, cudaMemcpyHostToDevice, st));
    }
    bvh_init_leaves_kernel<<<B256, 256, 0, st>>>(sortedPrim, dLo, dHi, n, T, cid[0], ci);
    int m = n, cur = 0, rounds = 0;
    while (m > 1) {
        const int blocks = (m + PLOC_T - 1) / PLOC_T;
        ploc_nn_kernel<<<blocks, PLOC_T, 0, st>>>(cid[cur], m, R, T, neighbor);
        ploc_merge_kernel<<<blocks, PLOC_T, 0, st>>>(cid[cur], m, neighbor, T, counters, cidTmp, valid, ci, ct, maxLeaf);
        exclusive_scan_u32(valid, pos, (unsigned)m + 1, scanTmp, st);
        ploc_compact_kernel<<<blocks, PLOC_T, 0, st>>>(cidTmp, valid, pos, m, cid[cur ^ 1]);
        unsigned newM = 0;
        OPR_CUDA_CHECK(cudaMemcpyAsync(&newM, pos + m, 4, cudaMemcpyDeviceToHost, st));
        OPR_CUDA_CHECK(cudaStreamSynchronize(st));
        if ((int)newM >= m || newM == 0) throw std::runtime_error("BVH build: PLOC made no progress");
        m = (int)newM; cur ^= 1; rounds++;
    }
    int root2 = 0;
    OPR_CUDA_CHECK(cudaMemcpyAsync(&root2, cid[cur], 4, cudaMemcpyDeviceToHost, st));
    OPR_CUDA_CHECK(cudaStreamSynchronize(st));

    // ---- collapse
    float4* wideTmp = (float4*)dalloc(128 * (size_t)std::max(n, 1));
    int2* queue[2] = {(int2*)dalloc(8 * (size_t)std::max(n, 1)), (int2*)dalloc(8 * (size_t)std::max(n, 1))};
    int* leafPrim = nullptr;
    devAlloc((void**)&leafPrim, 4 * (size_t)n); owned.push_back(leafPrim);
    {
        int h[4] = {1, 0, 0, 0};
        OPR_CUDA_CHECK(cudaMemcpyAsync(counters, h, sizeof h, cudaMemcpyHostToDevice, st));
        int2 first = make_int2(root2, 0);
        OPR_CUDA_CHECK(cudaMemcpyAsync(queue[0], &first, 8, cudaMemcpyHostToDevice, st));
    }
    int numIn = 1, qi = 0, levels = 0;
    while (numIn > 0) {
        OPR_CUDA_CHECK(cudaMemsetAsync(counters + 1, 0, 4, st));
        bvh_collapse_kernel<<<(numIn + 127) / 128, 128, 0, st>>>(queue[qi], numIn, T, n, wideTmp, counters, queue[qi ^ 1], leafPrim);
        OPR_CUDA_CHECK(cudaMemcpyAsync(&numIn, counters + 1, 4, cudaMemcpyDeviceToHost, st));
        OPR_CUDA_CHECK(cudaStreamSynchronize(st));
        qi ^= 1; levels++;
    }
    int h[4];
    OPR_CUDA_CHECK(cudaMemcpy(h, counters, sizeof h, cudaMemcpyDeviceToHost));
    const int numWide = h[0];
    if (h[2] != n) throw std::runtime_error("BVH build: leaf slots do not cover the primitives");
    if (3 * levels + 4 > TRAVERSAL_STACK) throw std::runtime_error("BVH build: tree deeper than the traversal stack");
    // one allocation: quantised nodes | fp32 nodes | leaf triangle records — what every traversal touches, as ONE address range, so that an
    // L2 access-policy window can cover it (Renderer::initScene, OPR_L2_PERSIST)
    const size_t qBytes = 64 * (size_t)numWide, nBytes = 128 * (size_t)numWide, tBytes = sizeof(float4) * LEAF_TRI_STRIDE * (size_t)n;
    char* block = nullptr;
    devAlloc((void**)&block, qBytes + nBytes + tBytes); owned.push_back(block);
    traversalBlock = block; traversalNodeBytes = qBytes + nBytes; traversalBytes = qBytes + nBytes + tBytes;
    float4* nodes = (float4*)(block + qBytes);
    OPR_CUDA_CHECK(cudaMemcpyAsync(nodes, wideTmp, 128 * (size_t)numWide, cudaMemcpyDeviceToDevice, st));
    float4* leafTri = (float4*)(block + qBytes + nBytes);
    bvh_leaf_tri_kernel<<<B256, 256, 0, st>>>(leafPrim, n, view.triRec, leafTri);
    OPR_CUDA_CHECK(cudaStreamSynchronize(st));
    OPR_CUDA_CHECK(cudaGetLastError());

    float rootCost = 0.f; float4 rlo, rhi;
    OPR_CUDA_CHECK(cudaMemcpy(&rootCost, T.cost + root2, 4, cudaMemcpyDeviceToHost));
    OPR_CUDA_CHECK(cudaMemcpy(&rlo, T.lo + root2, 16, cudaMemcpyDeviceToHost));
    OPR_CUDA_CHECK(cudaMemcpy(&rhi, T.hi + root2, 16, cudaMemcpyDeviceToHost));
    const float rootA = (rhi.x - rlo.x) * (rhi.y - rlo.y) + (rhi.y - rlo.y) * (rhi.z - rlo.z) + (rhi.z - rlo.z) * (rhi.x - rlo.x);
    sahCost = rootA > 0.f ? rootCost / rootA : 0.f;
    view.nodes = nodes; view.rootRef = 0; view.leafPrim = leafPrim; view.leafTri = leafTri;
    {   // quantised copy on a grid that frames the root box with a margin of a few cells
        const float3 ext = f3(rhi.x - rlo.x, rhi.y - rlo.y, rhi.z - rlo.z);
        const float big = std::max(std::max(ext.x, ext.y), std::max(ext.z, 1e-20f));
        auto axis = [&](float lo, float e, float& origin, float& extent) {
            const float e2 = std::max(e, big * 1e-4f);           // flat scenes: keep the cell size finite
            extent = e2 * (1.0f + 16.0f / 32768.0f);
            origin = lo - 8.0f * extent / 32768.0f;
        };
        float3 qo, qe;
        axis(rlo.x, ext.x, qo.x, qe.x); axis(rlo.y, ext.y, qo.y, qe.y); axis(rlo.z, ext.z, qo.z, qe.z);
        uint4* qnodes = (uint4*)block;
        bvh_quantize_kernel<<<(numWide + 255) / 256, 256, 0, st>>>(nodes, numWide, qo, f3(32768.0f / qe.x, 32768.0f / qe.y, 32768.0f / qe.z), qnodes);
        bvh_centre_half_kernel<<<(numWide + 255) / 256, 256, 0, st>>>(nodes, numWide);      // after the quantiser, which reads lo / hi
        OPR_CUDA_CHECK(cudaStreamSynchronize(st));
        OPR_CUDA_CHECK(cudaGetLastError());
        view.qnodes = (const float4*)qnodes; view.qOrigin = qo; view.qExtent = qe;
    }
    numNodes = numWide; bvhLevels = levels; plocRounds = rounds;
    if (getenv("OPR_BVH_VERBOSE"))
        fprintf(stderr, "[opr] BVH: %d prims, PLOC R=%d %d rounds, %d wide nodes (%.1f MB), %d levels, SAH %.2f\n", n, R, rounds, numWide, numWide * 128.0 / 1e6,
                levels, sahCost);
}

} // namespace opr
