// devmem.cu — device allocations of the path, with optional guard bands.
//
// compute-sanitizer is not available where this library is tested, so out-of-bounds WRITES are caught the old way: with
// opr_debug_set_guards(1) every allocation made afterwards carries a 4 KiB band of 0xA5 bytes on either side, and
// opr_debug_check_guards() reads all bands back and counts the allocations whose band was touched.  tests/test_parity_gpu.py
// renders every method with the guards on.  Off (the default) this is plain cudaMalloc / cudaFree.
#include "common.cuh"
#include <mutex>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

namespace opr {

static constexpr size_t GUARD = 4096;
static constexpr unsigned char PATTERN = 0xA5;
static bool g_guards = false;
static std::mutex g_mutex;
struct GuardedBlock { char* base; size_t bytes; };
static std::unordered_map<void*, GuardedBlock> g_blocks;
static int g_freedDamaged = 0, g_freedChecked = 0;

void devAlloc(void** out, size_t bytes)
{
    bytes = bytes < 16 ? 16 : bytes;
    std::lock_guard<std::mutex> lock(g_mutex);
    if (!g_guards) { OPR_CUDA_CHECK(cudaMalloc(out, bytes)); return; }
    char* base = nullptr;
    OPR_CUDA_CHECK(cudaMalloc((void**)&base, bytes + 2 * GUARD));
    OPR_CUDA_CHECK(cudaMemset(base, PATTERN, GUARD));
    OPR_CUDA_CHECK(cudaMemset(base + GUARD + bytes, PATTERN, GUARD));
    *out = base + GUARD;
    g_blocks[*out] = GuardedBlock{base, bytes};
}

static bool bandsDamaged(const GuardedBlock& b)
{
    std::vector<unsigned char> host(2 * GUARD);
    OPR_CUDA_CHECK(cudaMemcpy(host.data(), b.base, GUARD, cudaMemcpyDeviceToHost));
    OPR_CUDA_CHECK(cudaMemcpy(host.data() + GUARD, b.base + GUARD + b.bytes, GUARD, cudaMemcpyDeviceToHost));
    for (unsigned char c : host) if (c != PATTERN) return true;
    return false;
}

void devFree(void* p)
{
    if (!p) return;
    std::lock_guard<std::mutex> lock(g_mutex);
    auto it = g_blocks.find(p);
    if (it == g_blocks.end()) { cudaFree(p); return; }
    // a guarded block is checked once more when it dies (temporaries of the BVH build and of the stand-alone entry points)
    if (cudaDeviceSynchronize() == cudaSuccess) {
        try { g_freedDamaged += bandsDamaged(it->second); g_freedChecked++; } catch (...) { g_freedDamaged++; }
    }
    cudaFree(it->second.base);
    g_blocks.erase(it);
}

void devSetGuards(bool on)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    if (on && !g_guards) g_freedDamaged = g_freedChecked = 0;
    g_guards = on;
}

// number of guarded allocations (live now, or freed since the guards were switched on) with a damaged band;
// -1 when no guarded allocation has been checked at all
int devCheckGuards(int* checkedBlocks)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    if (checkedBlocks) *checkedBlocks = (int)g_blocks.size() + g_freedChecked;
    if (g_blocks.empty() && g_freedChecked == 0) return -1;
    OPR_CUDA_CHECK(cudaDeviceSynchronize());
    int bad = g_freedDamaged;
    for (auto& kv : g_blocks) bad += bandsDamaged(kv.second);
    return bad;
}

int numSMs()
{
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (!cached[dev]) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;   // B200
        cached[dev] = n;
    }
    return cached[dev];
}

} // namespace opr
