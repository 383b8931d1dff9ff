// capi.cu — the extern "C" boundary declared in include/opr.h.  No exceptions cross it.
#include "renderer.h"
#include "radix_sort.cuh"
#include "../../include/opr.h"
#include <cstring>
#include <memory>

using namespace opr;

static thread_local std::string g_lastError;
extern "C" void opr_set_last_error(const char* msg) { g_lastError = msg ? msg : ""; }   // group.cu reports through the same slot
#define OPR_TRY(...) try { __VA_ARGS__; return 0; } catch (const std::exception& e) { g_lastError = e.what(); return 1; } catch (...) { g_lastError = "unknown error"; return 1; }

namespace {
struct DevBuf {
    void* p = nullptr;
    explicit DevBuf(size_t bytes) { opr::devAlloc(&p, bytes); }
    ~DevBuf() { opr::devFree(p); }
    template <class T> T* as() { return (T*)p; }
};
void requireDevice()
{
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
        throw std::runtime_error("no CUDA device: the oppositerenderer_b200 kernels have no CPU path");
}
template <class T> std::vector<T> download(const T* d, size_t n, cudaStream_t st)
{
    std::vector<T> h(n);
    if (n) { OPR_CUDA_CHECK(cudaMemcpyAsync(h.data(), d, n * sizeof(T), cudaMemcpyDeviceToHost, st)); OPR_CUDA_CHECK(cudaStreamSynchronize(st)); }
    return h;
}
}  // namespace

extern "C" {

const char* opr_last_error(void) { return g_lastError.c_str(); }

int opr_create(opr_renderer** out, const opr_config* cfg)
{
    OPR_TRY({
        std::unique_ptr<opr_renderer> h(new opr_renderer());
        if (cfg) {
            if (cfg->photonLaunchWidth) h->r.photonLaunchWidth = cfg->photonLaunchWidth;
            if (cfg->photonLaunchHeight) h->r.photonLaunchHeight = cfg->photonLaunchHeight;
            if (cfg->vcmMaxPathLength) h->r.vcmMaxPathLength = cfg->vcmMaxPathLength;
            if (cfg->seedIsSet) h->r.seed = cfg->randomSeed;
        }
        *out = h.release();
    })
}

void opr_destroy(opr_renderer* r) { delete r; }
int opr_initialize(opr_renderer* r, int ordinal) { OPR_TRY(r->r.initialize(ordinal)) }
int opr_init_scene(opr_renderer* r, const opr_scene_desc* scene) { OPR_TRY(r->r.initScene(scene)) }
int opr_render_next_iteration(opr_renderer* r, uint64_t it, uint64_t lit, float radius, int createOutput, const opr_request_details* d)
{ OPR_TRY(r->r.renderNextIteration(it, lit, radius, createOutput != 0, *d)) }
int opr_get_output_buffer(opr_renderer* r, void* dst) { OPR_TRY(r->r.getOutputBuffer(dst)) }
uint32_t opr_get_width(opr_renderer* r) { return r->r.width; }
uint32_t opr_get_height(opr_renderer* r) { return r->r.height; }
uint32_t opr_get_screen_buffer_size_bytes(opr_renderer* r) { return r->r.width * r->r.height * 3u * (uint32_t)sizeof(float); }
void* opr_get_output_device_ptr(opr_renderer* r)
{   // the pointer is only handed out once the stream-B work is ordered before `stream`: a caller would otherwise read a half-written sum
    try { OPR_CUDA_CHECK(cudaSetDevice(r->r.device)); r->r.joinStreams(); return r->r.output; }
    catch (const std::exception& e) { g_lastError = e.what(); return nullptr; }
}
int opr_get_output_buffer_async(opr_renderer* r, void* dst) { OPR_TRY(r->r.getOutputBufferAsync(dst)) }
int opr_wait_output(opr_renderer* r) { OPR_TRY(r->r.waitOutput()) }
void* opr_get_stream(opr_renderer* r) { return (void*)r->r.stream; }
int opr_synchronize(opr_renderer* r) { OPR_TRY({ OPR_CUDA_CHECK(cudaSetDevice(r->r.device)); r->r.joinStreams(); OPR_CUDA_CHECK(cudaStreamSynchronize(r->r.stream)); }) }
int opr_join_streams(opr_renderer* r) { OPR_TRY({ OPR_CUDA_CHECK(cudaSetDevice(r->r.device)); r->r.joinStreams(); }) }

int opr_get_stats(opr_renderer* r, uint64_t* out, int n)
{
    OPR_TRY({
        Renderer& R = r->r;
        if (R.stats64) {   // totals live on the device (folded after every iteration); trace_closest / trace_any calls add to the 32-bit words
            auto t = download(R.stats64, STAT_NUM, R.stream);
            R.stats[STAT_CLOSEST_RAYS] = t[STAT_CLOSEST_RAYS]; R.stats[STAT_ANY_RAYS] = t[STAT_ANY_RAYS];
            if (R.buffersAllocated) {
                auto c = download(R.q.counters + CTR_STATS, STAT_NUM, R.stream);
                R.stats[STAT_CLOSEST_RAYS] += c[STAT_CLOSEST_RAYS]; R.stats[STAT_ANY_RAYS] += c[STAT_ANY_RAYS];
            }
        }
        uint64_t all[4] = {R.stats[STAT_CLOSEST_RAYS], R.stats[STAT_ANY_RAYS], R.stats[STAT_PHOTONS_STORED], R.launches};
        for (int i = 0; i < n && i < 4; i++) out[i] = all[i];
    })
}

int opr_set_timing(opr_renderer* r, int enabled)
{
    Renderer& R = r->r;
    R.timing = enabled == 1; R.timeline = enabled == 2;
    if (!R.timeline) { for (auto& e : R.timelineEvents) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); } R.timelineEvents.clear(); }
    return 0;
}
int opr_get_timeline(opr_renderer* r, float* startMs, float* endMs, char* streams, int capacity)
{
    Renderer& R = r->r;
    try {
        OPR_CUDA_CHECK(cudaSetDevice(R.device));
        OPR_CUDA_CHECK(cudaDeviceSynchronize());
        const int n = (int)R.timelineEvents.size();
        for (int i = 0; i < n && i < capacity; i++) {
            const auto& e = R.timelineEvents[i];
            OPR_CUDA_CHECK(cudaEventElapsedTime(&startMs[i], R.timelineEvents[0].a, e.a));
            OPR_CUDA_CHECK(cudaEventElapsedTime(&endMs[i], R.timelineEvents[0].a, e.b));
            streams[i] = e.stream;
        }
        return n;
    } catch (const std::exception& e) { g_lastError = e.what(); return -1; }
}
const char* opr_get_timeline_name(opr_renderer* r, int index)
{
    if (index < 0 || index >= (int)r->r.timelineEvents.size()) return "";
    return r->r.timelineEvents[index].name;
}
int opr_get_timings_ms(opr_renderer* r, float* ms, int capacity)
{
    int n = (int)r->r.lastTimings.size();
    for (int i = 0; i < n && i < capacity; i++) ms[i] = r->r.lastTimings[i].second;
    return n;
}
const char* opr_get_timing_name(opr_renderer* r, int index)
{
    if (index < 0 || index >= (int)r->r.lastTimings.size()) return "";
    return r->r.lastTimings[index].first.c_str();
}

int opr_debug_get_buffer(opr_renderer* r, const char* name, void* dst, uint64_t bytes)
{
    OPR_TRY({
        Renderer& R = r->r;
        OPR_CUDA_CHECK(cudaSetDevice(R.device));
        R.joinStreams();
        cudaStream_t st = R.stream;
        const std::string n(name);
        const size_t px = (size_t)R.width * R.height;
        const size_t P = (size_t)R.photonLaunchWidth * R.photonLaunchHeight;
        auto need = [&](size_t b) { if (bytes < b) throw std::runtime_error("opr_debug_get_buffer: destination too small for " + n); };
        if (n == "hitpoints") {
            need(px * 64);
            auto pos = download(R.hp.pos, px, st); auto nrm = download(R.hp.nrm, px, st);
            auto att = download(R.hp.att, px, st); auto rad = download(R.hp.rad, px, st);
            float* o = (float*)dst;
            for (size_t i = 0; i < px; i++, o += 16) {
                o[0] = pos[i].x; o[1] = pos[i].y; o[2] = pos[i].z; o[3] = nrm[i].x; o[4] = nrm[i].y; o[5] = nrm[i].z;
                o[6] = att[i].x; o[7] = att[i].y; o[8] = att[i].z; o[9] = rad[i].x; o[10] = rad[i].y; o[11] = rad[i].z;
                memcpy(o + 12, &pos[i].w, 4); o[13] = o[14] = o[15] = 0.f;
            }
        } else if (n == "photons" || n == "sorted_photons") {
            const bool srt = n == "sorted_photons";
            const size_t cnt = srt ? R.latestGrid().numValid : P * MAX_PHOTON_DEPOSITS;
            need(cnt * 36);
            const PhotonBuf& pb = srt ? R.sorted : R.photons;
            auto a = download(pb.a, cnt, st); auto b = download(pb.b, cnt, st); auto c = download(pb.c, cnt, st);
            std::vector<unsigned> state;
            if (!srt) state = download(R.phState, P, st);
            float* o = (float*)dst;
            for (size_t i = 0; i < cnt; i++, o += 9) {
                bool stored = srt || ((i % MAX_PHOTON_DEPOSITS) < ((state[i / MAX_PHOTON_DEPOSITS] >> 8) & 0xffu));
                if (stored) { o[0] = b[i].x; o[1] = b[i].y; o[2] = b[i].z; o[3] = a[i].x; o[4] = a[i].y; o[5] = a[i].z; o[6] = a[i].w; o[7] = b[i].w; o[8] = c[i]; }
                else { for (int k = 0; k < 9; k++) o[k] = 0.f; }   // the reference clears power+position of unused slots
            }
        } else if (n == "photon_keys") {
            need(P * MAX_PHOTON_DEPOSITS * 4);
            OPR_CUDA_CHECK(cudaMemcpyAsync(dst, R.sortedKeys, P * MAX_PHOTON_DEPOSITS * 4, cudaMemcpyDeviceToHost, st));
            OPR_CUDA_CHECK(cudaStreamSynchronize(st));
        } else if (n == "offsets") {
            const size_t nc = R.latestGrid().numCells;
            need((nc + 1) * 4);
            OPR_CUDA_CHECK(cudaMemcpyAsync(dst, R.gridHist, (nc + 1) * 4, cudaMemcpyDeviceToHost, st));
            OPR_CUDA_CHECK(cudaStreamSynchronize(st));
        } else if (n == "indirect" || n == "direct") {
            need(px * 12);
            OPR_CUDA_CHECK(cudaMemcpyAsync(dst, n == "indirect" ? R.indirect : R.direct, px * 12, cudaMemcpyDeviceToHost, st));
            OPR_CUDA_CHECK(cudaStreamSynchronize(st));
        } else if (n == "random_states") {
            const size_t ns = (size_t)R.rsW * R.rsH;
            need(ns * 24);
            auto a = download(R.rng.s, 2 * ns, st);
            unsigned* o = (unsigned*)dst;
            for (size_t i = 0; i < ns; i++, o += 6) { o[0] = a[2 * i].x; o[1] = a[2 * i].y; o[2] = a[2 * i].z; o[3] = a[2 * i].w; o[4] = a[2 * i + 1].x; o[5] = a[2 * i + 1].y; }
        } else if (n == "grid") {
            need(32);
            float* o = (float*)dst; unsigned* u = (unsigned*)dst;
            const GridParams G = R.latestGrid();
            o[0] = G.cellSize; u[1] = G.gridSize.x; u[2] = G.gridSize.y; u[3] = G.gridSize.z;
            o[4] = G.origo.x; o[5] = G.origo.y; o[6] = G.origo.z; u[7] = G.numValid;
        } else if (n == "photon_path_lengths") {
            need(P * 4);
            auto s = download(R.phState, P, st);
            unsigned* o = (unsigned*)dst;
            for (size_t i = 0; i < P; i++) o[i] = s[i] & 0xffu;
        } else if (n == "queue_counts") {
            // ray-queue occupancy of the last PPM iteration: [0] capacity, [1] shadow rays queued, [2..18) camera-pass ray counts per
            // bounce (snapshot taken when the photon pass reset the counters), [18..34) photon-pass ray counts per bounce
            need(34 * 4);
            unsigned* o = (unsigned*)dst;
            auto c = download(R.q.counters, CTR_ALLOC, st);
            o[0] = R.q.capacity; o[1] = c[CTR_SHADOW_COUNT] + c[CTR_SHADOW2_COUNT];
            for (int i = 0; i < 16; i++) { o[2 + i] = c[CTR_SNAPSHOT + i]; o[18 + i] = c[CTR_RAYCOUNT + i]; }
            for (int i = 1; i < 16; i++) if (o[2 + i] == 0) { o[2 + i] = c[CTR_TAIL_COUNT]; break; }   // what the camera tail kernel started from
        } else if (n == "light_vertex_counts") {
            need(px * 4);
            if (R.debugLightVertexCounts((unsigned*)dst)) throw std::runtime_error("no BDPT iteration has run yet");
        } else if (n == "connection_counts") {
            need(px * 4);
            if (R.debugConnectionCountsGet((unsigned*)dst)) throw std::runtime_error("connection counts are off (opr_debug_set_connection_counts) or no BDPT iteration has run");
        } else if (n == "path_signatures") {
            need(px * 8);
            if (R.debugPathSignaturesGet((unsigned*)dst)) throw std::runtime_error("path signatures are off (opr_debug_set_connection_counts) or no BDPT iteration has run");
        } else if (n == "pt_signature") {
            need(px * 4);
            if (!R.ptSignature) throw std::runtime_error("path signatures are off (opr_debug_set_connection_counts) or no PATH_TRACING iteration has run");
            OPR_CUDA_CHECK(cudaMemcpyAsync(dst, R.ptSignature, px * 4, cudaMemcpyDeviceToHost, st));
            OPR_CUDA_CHECK(cudaStreamSynchronize(st));
        } else if (n == "direct_sample_mask") {
            // bit s of byte p: shadow sample s of pixel p was cast, unoccluded and non-zero (last PPM iteration)
            need(px);
            auto f = download(R.sampleFactor, px, st);
            unsigned char* o = (unsigned char*)dst;
            for (size_t i = 0; i < px; i++) o[i] = (unsigned char)((f[i].x > 0.f) | ((f[i].y > 0.f) << 1) | ((f[i].z > 0.f) << 2) | ((f[i].w > 0.f) << 3));
        } else if (n == "sample_factors") {
            // the four light samples of every pixel (last PPM iteration): un-occluded factor (0 = not cast or occluded), 16 B per pixel
            need(px * 16);
            OPR_CUDA_CHECK(cudaMemcpyAsync(dst, R.sampleFactor, px * 16, cudaMemcpyDeviceToHost, st));
            OPR_CUDA_CHECK(cudaStreamSynchronize(st));
        } else if (n == "sample_lights") {
            need(px * 8);       // light index of the four samples, 16 bits each
            OPR_CUDA_CHECK(cudaMemcpyAsync(dst, R.sampleLight, px * 8, cudaMemcpyDeviceToHost, st));
            OPR_CUDA_CHECK(cudaStreamSynchronize(st));
        } else if (n == "light_vertex_positions") {
            // [0] = number of stored light vertices, then (path, x, y, z) 16-byte records, path-major
            need(16);
            const long long cnt = R.debugLightVertexPositions((float*)dst + 4, (size_t)(bytes / 16) - 1);
            if (cnt < 0) throw std::runtime_error("no BDPT iteration has run yet");
            memset(dst, 0, 16); const unsigned c32 = (unsigned)cnt; memcpy(dst, &c32, 4);
        } else {
            throw std::runtime_error("opr_debug_get_buffer: unknown buffer " + n);
        }
    })
}

int opr_debug_set_trace_nodes(opr_renderer* r, int quantized) { r->r.rawQuantized = quantized != 0; return 0; }
uint32_t opr_debug_slot_pixel(uint32_t slot, uint32_t width, uint32_t height) { return opr::tiled_pixel(slot, width, height); }
int opr_debug_set_connection_counts(opr_renderer* r, int enabled) { r->r.debugConnectionCounts = enabled != 0; return 0; }
void opr_debug_set_guards(int enabled) { opr::devSetGuards(enabled != 0); }
int opr_debug_check_guards(int* liveBlocks)
{
    try { return opr::devCheckGuards(liveBlocks); }
    catch (const std::exception& e) { g_lastError = e.what(); return -2; }
}

// ---- stand-alone kernels ------------------------------------------------------------------------------------
__global__ void rng_init_range_kernel(unsigned* out6, unsigned seed, unsigned first, unsigned count)
{
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    Rng s; rng_seed(s, seed, first + i);
    unsigned* o = out6 + 6 * (size_t)i;
    o[0] = s.d; o[1] = s.v0; o[2] = s.v1; o[3] = s.v2; o[4] = s.v3; o[5] = s.v4;
}
__global__ void rng_uniform_kernel(unsigned* state6, float* out, unsigned n)
{
    Rng s; s.d = state6[0]; s.v0 = state6[1]; s.v1 = state6[2]; s.v2 = state6[3]; s.v3 = state6[4]; s.v4 = state6[5];
    for (unsigned i = 0; i < n; i++) out[i] = rng_uniform(s);
    state6[0] = s.d; state6[1] = s.v0; state6[2] = s.v1; state6[3] = s.v2; state6[4] = s.v3; state6[5] = s.v4;
}

int opr_rng_init(uint32_t* states6, uint32_t seed, uint32_t first, uint32_t count)
{
    OPR_TRY({
        requireDevice();
        if (count == 0) return 0;
        DevBuf d((size_t)count * 24);
        rng_init_range_kernel<<<(count + 255) / 256, 256>>>(d.as<unsigned>(), seed, first, count);
        OPR_CUDA_CHECK(cudaMemcpy(states6, d.p, (size_t)count * 24, cudaMemcpyDeviceToHost));
    })
}

int opr_rng_uniform(uint32_t* state6, float* out, uint32_t n)
{
    OPR_TRY({
        requireDevice();
        DevBuf ds(24), dout((size_t)n * 4);
        OPR_CUDA_CHECK(cudaMemcpy(ds.p, state6, 24, cudaMemcpyHostToDevice));
        rng_uniform_kernel<<<1, 1>>>(ds.as<unsigned>(), dout.as<float>(), n);
        OPR_CUDA_CHECK(cudaMemcpy(out, dout.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
        OPR_CUDA_CHECK(cudaMemcpy(state6, ds.p, 24, cudaMemcpyDeviceToHost));
    })
}

// AoS (reference layout) -> the SoA slot layout the build kernels consume; every slot counts as "stored"
__global__ void photons_aos_to_soa_kernel(const float* __restrict__ aos, unsigned n, PhotonBuf pb, unsigned* __restrict__ phState)
{
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float* p = aos + 9 * (size_t)i;
        pb.a[i] = make_float4(p[3], p[4], p[5], p[6]);
        pb.b[i] = make_float4(p[0], p[1], p[2], p[7]);
        pb.c[i] = p[8];
        if ((i % MAX_PHOTON_DEPOSITS) == 0) phState[i / MAX_PHOTON_DEPOSITS] = (MAX_PHOTON_DEPOSITS << 8);
    }
}
__global__ void photons_soa_to_aos_kernel(PhotonBuf pb, unsigned n, float* __restrict__ aos)
{
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float* p = aos + 9 * (size_t)i;
        const float4 a = pb.a[i], b = pb.b[i];
        p[0] = b.x; p[1] = b.y; p[2] = b.z; p[3] = a.x; p[4] = a.y; p[5] = a.z; p[6] = a.w; p[7] = b.w; p[8] = pb.c[i];
    }
}

int opr_build_photon_grid(const float* photonsAoS, uint32_t n, float* sortedPhotonsAoS, uint32_t* keys, uint32_t* offsets,
                          float* cellSize, uint32_t* gridSize3, float* origo3, uint32_t* numValid)
{
    try {
        requireDevice();
        if (n == 0) throw std::runtime_error("opr_build_photon_grid: n == 0");
        cudaStream_t st = 0;
        unsigned long long launches = 0;
        LaunchCtx lc{st, &launches};
        const unsigned nPad = (n + MAX_PHOTON_DEPOSITS - 1) / MAX_PHOTON_DEPOSITS * MAX_PHOTON_DEPOSITS;
        DevBuf dAos((size_t)n * 36), a((size_t)nPad * 16), b((size_t)nPad * 16), c((size_t)nPad * 4), state((size_t)nPad);
        DevBuf sa((size_t)nPad * 16), sb((size_t)nPad * 16), sc_((size_t)nPad * 4);
        DevBuf k0((size_t)nPad * 4), k1((size_t)nPad * 4), v0((size_t)nPad * 4), v1((size_t)nPad * 4);
        DevBuf hist(((size_t)PHOTON_GRID_MAX_SIZE + 2) * 4), sortTmp(rs_hist_words(nPad) * 4), scanTmp(scan_tmp_words(PHOTON_GRID_MAX_SIZE + 2) * 4), bbox(32);
        PhotonBuf pb{a.as<float4>(), b.as<float4>(), c.as<float>()}, sorted{sa.as<float4>(), sb.as<float4>(), sc_.as<float>()};
        OPR_CUDA_CHECK(cudaMemcpy(dAos.p, photonsAoS, (size_t)n * 36, cudaMemcpyHostToDevice));
        OPR_CUDA_CHECK(cudaMemset(state.p, 0, nPad));
        OPR_CUDA_CHECK(cudaMemset(b.p, 0, (size_t)nPad * 16));
        photons_aos_to_soa_kernel<<<numSMs() * 4, 256>>>(dAos.as<float>(), n, pb, state.as<unsigned>());
        float init[8] = {FLT_MAX, FLT_MAX, FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX, 0.f, 0.f};
        OPR_CUDA_CHECK(cudaMemcpy(bbox.p, init, 28, cudaMemcpyHostToDevice));
        launch_photon_bbox(pb, state.as<unsigned>(), nPad / MAX_PHOTON_DEPOSITS, bbox.as<float>(), lc);
        float h[8];
        OPR_CUDA_CHECK(cudaMemcpy(h, bbox.p, 28, cudaMemcpyDeviceToHost));
        unsigned nv; memcpy(&nv, h + 6, 4);
        float lo[3] = {h[0], h[1], h[2]}, hi[3] = {h[3], h[4], h[5]};
        if (nv == 0) { for (int i = 0; i < 3; i++) lo[i] = hi[i] = photonsAoS[3 + i]; }   // init = photons[0] (:125-126)
        GridParams g;
        bool ok = compute_grid_params(lo, hi, g);
        g.numValid = nv;
        *cellSize = g.cellSize; gridSize3[0] = g.gridSize.x; gridSize3[1] = g.gridSize.y; gridSize3[2] = g.gridSize.z;
        origo3[0] = g.origo.x; origo3[1] = g.origo.y; origo3[2] = g.origo.z; *numValid = nv;
        if (!ok) { g_lastError = "Too many cells in SpatialHash.cu, over defined PHOTON_GRID_MAX_SIZE."; return 2; }
        OPR_CUDA_CHECK(cudaMemset(hist.p, 0, ((size_t)g.numCells + 2) * 4));
        DevBuf gDev(sizeof(GridParams));
        OPR_CUDA_CHECK(cudaMemcpy(gDev.p, &g, sizeof(GridParams), cudaMemcpyHostToDevice));
        launch_grid_keys(pb, state.as<unsigned>(), nPad, gDev.as<GridParams>(), k0.as<unsigned>(), hist.as<unsigned>(), lc);
        int bits = 1; while ((1ull << bits) <= (unsigned long long)g.numCells + 1) bits++;
        bool flipped = radix_sort_pairs<unsigned>(k0.as<unsigned>(), v0.as<unsigned>(), k1.as<unsigned>(), v1.as<unsigned>(), nPad, bits,
                                                  sortTmp.as<unsigned>(), true, st);
        exclusive_scan_u32(hist.as<unsigned>(), hist.as<unsigned>(), g.numCells + 1, scanTmp.as<unsigned>(), st);
        const unsigned* sk = flipped ? k1.as<unsigned>() : k0.as<unsigned>();
        const unsigned* sv = flipped ? v1.as<unsigned>() : v0.as<unsigned>();
        launch_photon_reorder(pb, sv, nv, gDev.as<GridParams>(), sorted, lc);
        photons_soa_to_aos_kernel<<<numSMs() * 4, 256>>>(sorted, nv, dAos.as<float>());
        OPR_CUDA_CHECK(cudaDeviceSynchronize());
        OPR_CUDA_CHECK(cudaGetLastError());
        OPR_CUDA_CHECK(cudaMemcpy(sortedPhotonsAoS, dAos.p, (size_t)nv * 36, cudaMemcpyDeviceToHost));
        OPR_CUDA_CHECK(cudaMemcpy(keys, sk, (size_t)n * 4, cudaMemcpyDeviceToHost));
        OPR_CUDA_CHECK(cudaMemcpy(offsets, hist.p, ((size_t)g.numCells + 1) * 4, cudaMemcpyDeviceToHost));
        return 0;
    } catch (const std::exception& e) { g_lastError = e.what(); return 1; }
}

__global__ void hitpoints_aos_to_soa_kernel(const float* __restrict__ aos, unsigned n, HitpointBuf hp)
{
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float* p = aos + 16 * (size_t)i;
        hp.pos[i] = make_float4(p[0], p[1], p[2], p[12]);
        hp.nrm[i] = make_float4(p[3], p[4], p[5], 0.f);
        hp.att[i] = make_float4(p[6], p[7], p[8], 0.f);
        hp.rad[i] = make_float4(p[9], p[10], p[11], 0.f);
    }
}

int opr_gather(uint32_t numPixels, const float* hitpointsAoS, uint32_t numPhotons, const float* photonsAoS, const uint32_t* offsets,
               float cellSize, const uint32_t* gridSize3, const float* origo3, float ppmRadius, float emittedPhotons, float* indirectOut)
{
    OPR_TRY({
        requireDevice();
        cudaStream_t st = 0;
        unsigned long long launches = 0;
        LaunchCtx lc{st, &launches};
        GridParams g; g.cellSize = cellSize; g.gridSize = make_uint3(gridSize3[0], gridSize3[1], gridSize3[2]);
        g.origo = make_float3(origo3[0], origo3[1], origo3[2]); g.numCells = gridSize3[0] * gridSize3[1] * gridSize3[2]; g.numValid = numPhotons;
        DevBuf dHp((size_t)numPixels * 64), p0((size_t)numPixels * 16), p1((size_t)numPixels * 16), p2((size_t)numPixels * 16), p3((size_t)numPixels * 16);
        DevBuf dPh((size_t)numPhotons * 36), a((size_t)numPhotons * 16 + 64), b((size_t)numPhotons * 16 + 64), c((size_t)numPhotons * 4 + 64), state((size_t)numPhotons + 64);
        DevBuf dOff(((size_t)g.numCells + 1) * 4), dOut((size_t)numPixels * 12), dCounter(16);
        HitpointBuf hp{p0.as<float4>(), p1.as<float4>(), p2.as<float4>(), p3.as<float4>()};
        PhotonBuf pb{a.as<float4>(), b.as<float4>(), c.as<float>()};
        OPR_CUDA_CHECK(cudaMemcpy(dHp.p, hitpointsAoS, (size_t)numPixels * 64, cudaMemcpyHostToDevice));
        OPR_CUDA_CHECK(cudaMemcpy(dPh.p, photonsAoS, (size_t)numPhotons * 36, cudaMemcpyHostToDevice));
        OPR_CUDA_CHECK(cudaMemcpy(dOff.p, offsets, ((size_t)g.numCells + 1) * 4, cudaMemcpyHostToDevice));
        hitpoints_aos_to_soa_kernel<<<numSMs() * 4, 256>>>(dHp.as<float>(), numPixels, hp);
        if (numPhotons) photons_aos_to_soa_kernel<<<numSMs() * 4, 256>>>(dPh.as<float>(), numPhotons, pb, state.as<unsigned>());
        DevBuf gDev(sizeof(GridParams));
        OPR_CUDA_CHECK(cudaMemcpy(gDev.p, &g, sizeof(GridParams), cudaMemcpyHostToDevice));
        launch_gather(hp, numPixels, 0, 0, pb, dOff.as<unsigned>(), gDev.as<GridParams>(), ppmRadius, ppmRadius * ppmRadius, emittedPhotons, dOut.as<float>(), dCounter.as<unsigned>(), lc);
        OPR_CUDA_CHECK(cudaDeviceSynchronize());
        OPR_CUDA_CHECK(cudaGetLastError());
        OPR_CUDA_CHECK(cudaMemcpy(indirectOut, dOut.p, (size_t)numPixels * 12, cudaMemcpyDeviceToHost));
    })
}

__global__ void rays8_to_soa_kernel(const float* __restrict__ rays8, unsigned n, float4* __restrict__ o, float4* __restrict__ d)
{
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float* p = rays8 + 8 * (size_t)i;
        o[i] = make_float4(p[0], p[1], p[2], p[3]);
        d[i] = make_float4(p[4], p[5], p[6], p[7]);
    }
}

__global__ void hits_to_host_layout_kernel(SceneView sc, const float4* __restrict__ rayO, const float4* __restrict__ rayD, const float4* __restrict__ hits,
                                           unsigned n, float* __restrict__ outT, int* __restrict__ outMat, float* __restrict__ outNg)
{
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 h4 = hits[i];
        HitRec hit; hit.t = h4.x; hit.prim = __float_as_int(h4.y); hit.u = h4.z; hit.v = h4.w;
        if (hit.prim < 0) { outT[i] = -1.f; outMat[i] = -1; outNg[3 * i] = outNg[3 * i + 1] = outNg[3 * i + 2] = 0.f; continue; }
        const Surface s = eval_surface(sc, hit, f3(rayO[i]), f3(rayD[i]));
        outT[i] = hit.t; outMat[i] = s.material;
        outNg[3 * i] = s.geometricNormal.x; outNg[3 * i + 1] = s.geometricNormal.y; outNg[3 * i + 2] = s.geometricNormal.z;
    }
}

int opr_trace_closest(opr_renderer* r, const float* rays8, uint32_t n, float* outT, int32_t* outMaterial, float* outNg3)
{
    OPR_TRY({
        Renderer& R = r->r;
        if (!R.sceneLoaded) throw std::runtime_error("opr_trace_closest: no scene loaded");
        OPR_CUDA_CHECK(cudaSetDevice(R.device));
        LaunchCtx lc{R.stream, &R.launches};
        DevBuf dr((size_t)n * 32), o((size_t)n * 16), d((size_t)n * 16), h((size_t)n * 16), t((size_t)n * 4), m((size_t)n * 4), ng((size_t)n * 12);
        OPR_CUDA_CHECK(cudaMemcpyAsync(dr.p, rays8, (size_t)n * 32, cudaMemcpyHostToDevice, R.stream));
        rays8_to_soa_kernel<<<numSMs() * 4, 256, 0, R.stream>>>(dr.as<float>(), n, o.as<float4>(), d.as<float4>());
        launch_trace_closest_raw(R.scene.view, o.as<float4>(), d.as<float4>(), n, h.as<float4>(), lc, R.rawQuantized);
        hits_to_host_layout_kernel<<<numSMs() * 4, 256, 0, R.stream>>>(R.scene.view, o.as<float4>(), d.as<float4>(), h.as<float4>(), n, t.as<float>(), m.as<int>(), ng.as<float>());
        OPR_CUDA_CHECK(cudaMemcpyAsync(outT, t.p, (size_t)n * 4, cudaMemcpyDeviceToHost, R.stream));
        OPR_CUDA_CHECK(cudaMemcpyAsync(outMaterial, m.p, (size_t)n * 4, cudaMemcpyDeviceToHost, R.stream));
        OPR_CUDA_CHECK(cudaMemcpyAsync(outNg3, ng.p, (size_t)n * 12, cudaMemcpyDeviceToHost, R.stream));
        OPR_CUDA_CHECK(cudaStreamSynchronize(R.stream));
        OPR_CUDA_CHECK(cudaGetLastError());
    })
}

int opr_trace_any(opr_renderer* r, const float* rays8, uint32_t n, int32_t* outOccluded)
{
    OPR_TRY({
        Renderer& R = r->r;
        if (!R.sceneLoaded) throw std::runtime_error("opr_trace_any: no scene loaded");
        OPR_CUDA_CHECK(cudaSetDevice(R.device));
        LaunchCtx lc{R.stream, &R.launches};
        DevBuf dr((size_t)n * 32), o((size_t)n * 16), d((size_t)n * 16), occ((size_t)n * 4);
        OPR_CUDA_CHECK(cudaMemcpyAsync(dr.p, rays8, (size_t)n * 32, cudaMemcpyHostToDevice, R.stream));
        rays8_to_soa_kernel<<<numSMs() * 4, 256, 0, R.stream>>>(dr.as<float>(), n, o.as<float4>(), d.as<float4>());
        launch_trace_any_raw(R.scene.view, o.as<float4>(), d.as<float4>(), n, occ.as<int>(), lc, R.rawQuantized);
        OPR_CUDA_CHECK(cudaMemcpyAsync(outOccluded, occ.p, (size_t)n * 4, cudaMemcpyDeviceToHost, R.stream));
        OPR_CUDA_CHECK(cudaStreamSynchronize(R.stream));
        OPR_CUDA_CHECK(cudaGetLastError());
    })
}

}  // extern "C"
