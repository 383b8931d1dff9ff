// radix_sort.cuh — stable LSD radix sort (8-bit digits, or 10-bit for keys of 11..20 bits: two passes instead of three) of (key, uint32 value) pairs and a 3-kernel
// exclusive scan, hand-written for sm_100a.  Replaces thrust::sort_by_key / thrust::exclusive_scan of
// RenderEngine/renderer/OptixRenderer_SpatialHash.cu:193-207; also sorts Morton codes for the LBVH.
//
// Stability (needed for bit-exact parity with thrust's stable radix sort): inside a tile the keys are
// laid out warp-striped (warp w owns a contiguous chunk, round j / lane l reads chunk[j*32+l]); ranks
// come from __match_any_sync peer masks, so equal digits keep their input order.
#pragma once
#include "common.cuh"

namespace opr {

constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;
constexpr int RS_WARPS = RS_THREADS / 32;

template <typename Key, int DB>
__global__ void __launch_bounds__(RS_THREADS) rs_hist_kernel(const Key* __restrict__ keys, unsigned n, int shift,
                                                             unsigned* __restrict__ hist, unsigned numBlocks, bool skip, Key skipKey)
{
    constexpr unsigned ND = 1u << DB;
    __shared__ unsigned sh[ND];
    for (unsigned d = threadIdx.x; d < ND; d += RS_THREADS) sh[d] = 0;
    __syncthreads();
    const unsigned base = blockIdx.x * RS_TILE;
    #pragma unroll
    for (int j = 0; j < RS_ITEMS; j++) {
        unsigned i = base + j * RS_THREADS + threadIdx.x;
        if (i < n) { const Key k = keys[i]; if (!skip || k != skipKey) atomicAdd(&sh[(unsigned)(k >> shift) & (ND - 1)], 1u); }
    }
    __syncthreads();
    for (unsigned d = threadIdx.x; d < ND; d += RS_THREADS) hist[d * numBlocks + blockIdx.x] = sh[d];   // digit-major
}

template <typename Key, int DB>
__global__ void __launch_bounds__(RS_THREADS) rs_scatter_kernel(const Key* __restrict__ keysIn, const unsigned* __restrict__ valsIn,
                                                                Key* __restrict__ keysOut, unsigned* __restrict__ valsOut,
                                                                unsigned n, int shift, const unsigned* __restrict__ histScanned,
                                                                unsigned numBlocks, bool skip, Key skipKey)
{
    constexpr unsigned ND = 1u << DB;
    __shared__ unsigned warpCnt[RS_WARPS][ND];
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < RS_WARPS * (int)ND; i += RS_THREADS) (&warpCnt[0][0])[i] = 0;
    __syncthreads();

    const unsigned chunkBase = blockIdx.x * RS_TILE + w * (32 * RS_ITEMS);
    Key key[RS_ITEMS];
    unsigned val[RS_ITEMS];
    unsigned rank[RS_ITEMS];
    unsigned keep = 0;          // bit j: item j takes part
    #pragma unroll
    for (int j = 0; j < RS_ITEMS; j++) {
        unsigned i = chunkBase + j * 32 + lane;
        bool valid = i < n;
        key[j] = valid ? keysIn[i] : (Key)0;
        if (skip && key[j] == skipKey) valid = false;      // dropped from the sort: this pass compacts
        val[j] = valid ? (valsIn ? valsIn[i] : i) : 0u;
        unsigned digit = (unsigned)(key[j] >> shift) & (ND - 1);
        unsigned mkey = valid ? digit : (0x80000000u | lane);   // invalid lanes match nobody
        unsigned peers = __match_any_sync(0xffffffffu, mkey);
        unsigned leader = __ffs(peers) - 1;
        unsigned old = 0;
        if (valid && lane == leader) { old = warpCnt[w][digit]; warpCnt[w][digit] = old + __popc(peers); }
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[j] = old + __popc(peers & ((1u << lane) - 1));
        if (valid) keep |= 1u << j;
        __syncwarp();
    }
    __syncthreads();
    for (unsigned d = threadIdx.x; d < ND; d += RS_THREADS) {   // per digit: exclusive scan over the warps, seeded with this block's global offset
        unsigned running = histScanned[d * numBlocks + blockIdx.x];
        #pragma unroll
        for (int ww = 0; ww < RS_WARPS; ww++) { unsigned c = warpCnt[ww][d]; warpCnt[ww][d] = running; running += c; }
    }
    __syncthreads();
    #pragma unroll
    for (int j = 0; j < RS_ITEMS; j++) {
        if (keep & (1u << j)) {
            unsigned digit = (unsigned)(key[j] >> shift) & (ND - 1);
            unsigned pos = warpCnt[w][digit] + rank[j];
            keysOut[pos] = key[j];
            valsOut[pos] = val[j];
        }
    }
}

// ---- exclusive scan (uint32), three kernels: tile reduce, single-block scan of tile sums, tile scan ----
constexpr int SC_THREADS = 256;
constexpr int SC_ITEMS = 8;
constexpr int SC_TILE = SC_THREADS * SC_ITEMS;

OPR_DEV unsigned block_exclusive_scan(unsigned v, unsigned* total, unsigned* shWarp /* >= 33 */)
{
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    unsigned x = v;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) { unsigned y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) shWarp[w] = x;
    __syncthreads();
    if (w == 0) {
        unsigned s = lane < nw ? shWarp[lane] : 0;
        unsigned t = s;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) { unsigned y = __shfl_up_sync(0xffffffffu, t, o); if (lane >= o) t += y; }
        if (lane < nw) shWarp[lane] = t - s;
        if (lane == 31) shWarp[32] = t;
    }
    __syncthreads();
    unsigned res = x - v + shWarp[w];
    if (total) *total = shWarp[32];
    __syncthreads();
    return res;
}

static __global__ void __launch_bounds__(SC_THREADS) scan_reduce_kernel(const unsigned* __restrict__ in, unsigned n, unsigned* __restrict__ tileSums)
{
    __shared__ unsigned sh[33];
    unsigned base = blockIdx.x * SC_TILE + threadIdx.x * SC_ITEMS;
    unsigned s = 0;
    #pragma unroll
    for (int j = 0; j < SC_ITEMS; j++) if (base + j < n) s += in[base + j];
    unsigned total;
    block_exclusive_scan(s, &total, sh);
    if (threadIdx.x == 0) tileSums[blockIdx.x] = total;
}

// single block; scans `n` tile sums in place (exclusive)
static __global__ void __launch_bounds__(1024) scan_sums_kernel(unsigned* __restrict__ sums, unsigned n)
{
    __shared__ unsigned sh[33];
    unsigned carry = 0;
    for (unsigned base = 0; base < n; base += 1024) {
        unsigned i = base + threadIdx.x;
        unsigned v = i < n ? sums[i] : 0;
        unsigned total;
        unsigned e = block_exclusive_scan(v, &total, sh);
        if (i < n) sums[i] = e + carry;
        carry += total;
    }
}

static __global__ void __launch_bounds__(SC_THREADS) scan_apply_kernel(const unsigned* __restrict__ in, unsigned* __restrict__ out, unsigned n,
                                                                const unsigned* __restrict__ tileSums)
{
    __shared__ unsigned sh[33];
    unsigned base = blockIdx.x * SC_TILE + threadIdx.x * SC_ITEMS;
    unsigned v[SC_ITEMS];
    unsigned s = 0;
    #pragma unroll
    for (int j = 0; j < SC_ITEMS; j++) { v[j] = (base + j < n) ? in[base + j] : 0; s += v[j]; }
    unsigned e = block_exclusive_scan(s, nullptr, sh) + tileSums[blockIdx.x];
    #pragma unroll
    for (int j = 0; j < SC_ITEMS; j++) { if (base + j < n) out[base + j] = e; e += v[j]; }
}

// Host drivers.  `tmp` must hold scan_tmp_words(n) uint32.
inline size_t scan_tmp_words(unsigned n) { return (n + SC_TILE - 1) / SC_TILE + 1; }
inline void exclusive_scan_u32(const unsigned* in, unsigned* out, unsigned n, unsigned* tmp, cudaStream_t st)
{
    if (n == 0) return;
    unsigned tiles = (n + SC_TILE - 1) / SC_TILE;
    scan_reduce_kernel<<<tiles, SC_THREADS, 0, st>>>(in, n, tmp);
    scan_sums_kernel<<<1, 1024, 0, st>>>(tmp, tiles);
    scan_apply_kernel<<<tiles, SC_THREADS, 0, st>>>(in, out, n, tmp);
}

// Sorts n (key,value) pairs by the low `bits` bits of key, stable.  keys/vals are ping-ponged with
// keysAlt/valsAlt; returns true if the result ended in the Alt buffers.  If valsIn == nullptr on the first
// pass the values are the input indices 0..n-1.  hist must hold rs_hist_words(n) uint32.
inline size_t rs_hist_words(unsigned n) { return (size_t)1024 * ((n + RS_TILE - 1) / RS_TILE) + scan_tmp_words(1024 * ((n + RS_TILE - 1) / RS_TILE)); }

template <typename Key, int DB>
inline void radix_sort_pass(const Key* kin, const unsigned* vin, Key* kout, unsigned* vout, unsigned n, int shift, unsigned* hist, cudaStream_t st,
                            bool skip = false, Key skipKey = 0)
{
    const unsigned blocks = (n + RS_TILE - 1) / RS_TILE;
    unsigned* scanTmp = hist + (size_t)(1u << DB) * blocks;
    rs_hist_kernel<Key, DB><<<blocks, RS_THREADS, 0, st>>>(kin, n, shift, hist, blocks, skip, skipKey);
    exclusive_scan_u32(hist, hist, (1u << DB) * blocks, scanTmp, st);
    rs_scatter_kernel<Key, DB><<<blocks, RS_THREADS, 0, st>>>(kin, vin, kout, vout, n, shift, hist, blocks, skip, skipKey);
}

// Digit width: 8 bits, except keys of 11..20 bits (the photon cell keys: <= 10^6 + 1 cells), which are sorted in two 10-bit passes
// instead of three 8-bit ones (a third less key / value traffic; 32 KB of shared counters per block).
// dropKey / numKept (optional): items whose key equals *dropKey take no part — the first pass compacts them away (stable), the later
// passes run over the numKept items that remain, and the tail [numKept, n) of the result is filled with *dropKey (what a sort of all n
// items would have left there when *dropKey is the largest key: the invalid photon slots, 58 % of the 4 M slots of an iteration).
template <typename Key>
static __global__ void rs_fill_kernel(Key* __restrict__ p, unsigned from, unsigned to, Key v)
{
    for (unsigned i = from + blockIdx.x * blockDim.x + threadIdx.x; i < to; i += gridDim.x * blockDim.x) p[i] = v;
}

template <typename Key>
inline bool radix_sort_pairs(Key* keys, unsigned* vals, Key* keysAlt, unsigned* valsAlt, unsigned n, int bits,
                             unsigned* hist, bool valsAreIndices, cudaStream_t st, int* launches = nullptr,
                             const Key* dropKey = nullptr, unsigned numKept = 0)
{
    if (n == 0) return false;
    const int db = (bits > 10 && bits <= 20) ? 10 : 8;
    bool flipped = false;
    unsigned m = n;
    for (int shift = 0; shift < bits; shift += db) {
        Key* kin = flipped ? keysAlt : keys; Key* kout = flipped ? keys : keysAlt;
        unsigned* vin = flipped ? valsAlt : vals; unsigned* vout = flipped ? vals : valsAlt;
        const unsigned* v = (shift == 0 && valsAreIndices) ? nullptr : vin;
        const bool skip = dropKey && shift == 0;
        if (m) {
            if (db == 10) radix_sort_pass<Key, 10>(kin, v, kout, vout, m, shift, hist, st, skip, skip ? *dropKey : (Key)0);
            else radix_sort_pass<Key, 8>(kin, v, kout, vout, m, shift, hist, st, skip, skip ? *dropKey : (Key)0);
        }
        if (launches) *launches += 5;
        if (skip) m = numKept;
        flipped = !flipped;
    }
    if (dropKey && numKept < n) {
        Key* out = flipped ? keysAlt : keys;
        rs_fill_kernel<Key><<<256, 256, 0, st>>>(out, numKept, n, *dropKey);
        if (launches) *launches += 1;
    }
    return flipped;
}

} // namespace opr
