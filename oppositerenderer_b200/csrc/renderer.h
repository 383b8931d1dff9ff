// renderer.h — device buffers and host-side state of the B200 renderer behind the OptixRenderer surface
// (RenderEngine/renderer/OptixRenderer.h:21-115).  Buffer roles follow SURVEY.md Appendix C; layouts are
// SoA / 16-byte vectors instead of the reference's AoS user-format buffers.
#pragma once
#include "device_scene.h"
#include <string>
#include <vector>

namespace opr {

constexpr int NUM_QUEUES = 7;     // six material kinds + miss
constexpr int QUEUE_MISS = 6;
// RayQueues::counters layout (unsigned words).  The PPM passes use one counter set per bounce so that the bounce loop
// needs no memset: the generation kernel of a pass resets [CTR_RAYCOUNT, CTR_TOTAL) and seeds bounce 0.
constexpr int CTR_TILE = 12;          // gather tile queue cursor
constexpr int CTR_STATS = 16;         // StatSlot counters
constexpr int CTR_RAYCOUNT = 32;      // ray count of bounce b at CTR_RAYCOUNT + b (b < 32)
constexpr int CTR_MATCOUNT = 64;      // material-queue counts of bounce b at CTR_MATCOUNT + 8 b (b < 24)
constexpr int CTR_FETCH = 192;        // ray-fetch cursor of the persistent trace kernel, bounce b at CTR_FETCH + b (b < 32)
constexpr int CTR_FETCH_LEGACY = 10;
constexpr int CTR_SHADOW_COUNT = 13;  // shadow-ray queue length of the PPM direct pass
constexpr int CTR_FETCH_SHADOW = 14;  // its fetch cursor  // same for the ping-pong (VCM) launcher
constexpr int CTR_TOTAL = 256;        // [CTR_RAYCOUNT, CTR_TOTAL) is reset by the generation kernel of every pass
constexpr int CTR_SNAPSHOT = 256;     // 16 words: the ray counts of the pass before the current one (ray-queue occupancy report)
constexpr int CTR_TAIL_COUNT = 272;    // camera tail queue length (outside the per-pass reset range: the tail runs on its own stream)
constexpr int CTR_DEFERRED_COUNT = 273; // photon threads deferred behind the camera tail
constexpr int CTR_SHADOW2_COUNT = 277; // shadow-ray queue of the late direct samples (pixels deferred behind the tails): length, fetch cursor
constexpr int CTR_FETCH_SHADOW2 = 278;
constexpr int CTR_DIRECT_DEFERRED = 279; // pixels whose direct samples wait for a tail kernel
constexpr int CTR_CONN2_COUNT = 274; // second BDPT connection queue (vcm.cu): length, -, fetch cursor at 274..276
constexpr int CTR_CONN2_FETCH = 276;
constexpr int CTR_TAIL_FETCH = 280;   // path-fetch cursors of the regenerating tail kernels: camera tail, deferred photon threads
constexpr int CTR_LATE_FETCH = 281;   // (the photon tail's cursor is CTR_FETCH + its first bounce: that bounce has no trace kernel)
constexpr int CTR_ALLOC = 288;

// Wavefront ray queue state (all device pointers)
struct RayQueues {
    float4* rayO[2];        // origin.xyz, tmin
    float4* rayD[2];        // direction.xyz, tmax
    unsigned* rayPath[2];   // owning path (pixel / photon thread / subpath) of each ray slot
    float4* hits;           // HitRec per ray slot of the current queue
    unsigned* matQueue;     // NUM_QUEUES x capacity ray-slot indices, bucketed by material kind
    // camera tail (specular continuations after the last wavefront bounce) and the photon threads deferred behind it
    float4* tailO; float4* tailD; unsigned* tailPath; unsigned char* camPending; unsigned* deferred;
    // photon threads still walking after the last wavefront bounce (their RNG state is not final), and the pixels whose direct
    // light samples therefore wait for the tail kernels
    unsigned char* photonPending; unsigned* directDeferred;
    unsigned* counters;     // CTR_ALLOC words; VCM passes still use [0],[1] ray counts (ping-pong) + [2..2+NUM_QUEUES) material counts
    unsigned capacity;
};

struct HitpointBuf {        // raytracePassOutputBuffer (Hitpoint.h:10-18) as four float4 arrays, 64 B/pixel
    float4* pos;            // position.xyz, flags (uint bits)
    float4* nrm;            // normal.xyz, depth (uint bits)
    float4* att;            // attenuation.xyz
    float4* rad;            // radiance.xyz
};

struct PhotonBuf {          // `photons` (ppm/Photon.h:10-33, 36 B) as three arrays; slot = 4*thread + k
    float4* a;              // position.xyz, rayDirection.x
    float4* b;              // power.xyz, rayDirection.y
    float* c;               // rayDirection.z
};

struct GridParams { float cellSize; uint3 gridSize; float3 origo; unsigned numCells; unsigned numValid; };

enum StatSlot { STAT_CLOSEST_RAYS = 0, STAT_ANY_RAYS = 1, STAT_PHOTONS_STORED = 2, STAT_NUM = 8 };

struct Renderer {
    // configuration (runtime versions of OptixRenderer.cpp:44-53 / config.h:50)
    unsigned photonLaunchWidth = 1024, photonLaunchHeight = 1024;
    unsigned vcmMaxPathLength = 10;
    // wavefront bounces before the per-path tail kernels take over (camera: specular chains only; photons: after RR starts)
    unsigned cameraWaveBounces = 1, photonWaveBounces = 4, vcmLightWaveBounces = 4, vcmCameraWaveBounces = 6;
    unsigned seed = 1645301512u;

    bool initialized = false, sceneLoaded = false;
    int device = -1;
    cudaStream_t stream = nullptr;
    DeviceScene scene;

    unsigned width = 10, height = 10;     // OptixRenderer.cpp:84-85: the first request always resizes
    bool buffersAllocated = false;
    unsigned rsW = 0, rsH = 0;
    RngBuf rng{};
    HitpointBuf hp{};                     // hit points of the latest camera pass (= hpBuf[parity])
    HitpointBuf hpBuf[2]{};               // double-buffered: the gather of iteration i overlaps the camera/photon passes of i+1
    RayQueues q{};
    // second stream: photon-grid build, gather and output of iteration i run here while `stream` traces iteration i+1
    cudaStream_t streamB = nullptr;
    bool overlap = true;                  // OPR_OVERLAP=0 puts everything on `stream`
    unsigned gatherBlocksPerSM = 4;       // OPR_GATHER_BLOCKS_PER_SM: gather CTAs per SM while it shares the SMs with the next iteration's path passes
    unsigned gatherBlocksPerSMAlone = 12; // OPR_GATHER_BLOCKS_PER_SM_ALONE: persistent gather CTAs per SM when nothing overlaps (read-back pattern)
    unsigned gatherTilesPerBlock = 5;     // OPR_GATHER_TILES_PER_BLOCK: > 0 = short-lived gather CTAs (that many tiles each) that the block scheduler can interleave with higher-priority work
    int earlyDirect = -1;                 // OPR_EARLY_DIRECT: 1 = light samples of the pixels not behind a tail are drawn while the tails run (shorter critical path,
                                          // two more launches); -1 = only for a caller that reads every iteration back (nothing else fills the GPU then); 0 = never
    bool streamPriorities = true;         // OPR_STREAM_PRIORITY=0: all streams at the default priority
    unsigned vcmConnectBlocksPerSM = 3;   // OPR_VCM_CONNECT_BLOCKS_PER_SM: CTAs per SM of the connection kernel while it shares the SMs with the next bounce (0 = as the other trace kernels)
    bool joinedSinceRender = true;        // the caller consumed the last iteration before asking for the next one (no overlap possible)
    unsigned parity = 0;
    cudaEvent_t evHpFree[2] = {nullptr, nullptr};   // stream B finished with hpBuf[b] / sample buffers b
    cudaEvent_t evGridDone = nullptr;               // the grid build of the latest iteration is complete (sorted photons + offsets ready)
    cudaEvent_t evPhotonsFree[2] = {nullptr, nullptr};   // the grid build finished reading photonsBuf[b] / phStateBuf[b]
    cudaEvent_t evGatherDone = nullptr;             // stream B finished reading the sorted photons / offset table (the next build overwrites them)
    bool photonsFreePending[2] = {false, false}, gatherDonePending = false;
    cudaEvent_t evDirectDone = nullptr;             // stream A finished the light samples of this iteration
    cudaEvent_t evJoin = nullptr;
    // third stream: the camera tail and the photon threads deferred behind it run while `stream` traces the photon wavefronts
    cudaStream_t streamC = nullptr;
    cudaEvent_t evCameraWaves = nullptr, evPhotonGen = nullptr, evTailDone = nullptr, evPhotonWaves = nullptr;
    cudaStream_t sC() const { return overlap ? streamC : stream; }
    // Which launches traverse the quantised nodes (OPR_QNODES, bit mask): 1 photon wavefront bounces, 2 photon tail + deferred photons,
    // 4 camera tail, 8 primary rays, 16 shadow rays, 32 BDPT bounces >= 1, 64 BDPT connections, 128 BDPT tails.  Default: the incoherent ones.
    enum { QN_PHOTON = 1, QN_PHOTON_TAILS = 2, QN_CAMERA_TAIL = 4, QN_PRIMARY = 8, QN_SHADOW = 16, QN_VCM_BOUNCES = 32, QN_VCM_CONNECT = 64, QN_VCM_TAILS = 128 };
    unsigned qnodeMask = QN_PHOTON | QN_PHOTON_TAILS | QN_VCM_BOUNCES | QN_VCM_TAILS;
    int l2Persist = 0;                    // OPR_L2_PERSIST: L2 access-policy window over the BVH (0 off, 1 node arrays, 2 nodes + leaf triangles)
    bool rawQuantized = false;            // opr_debug_set_trace_nodes: opr_trace_closest / opr_trace_any traverse the quantised nodes
    bool gridOnC = false;                 // OPR_GRID_STREAM=C: photon-grid build on stream C (greatest priority) instead of the low-priority stream B (measured: no gain,
                                          // the build is slowed by the persistent trace kernels holding every register of the SMs, not by its priority)
    cudaStream_t gridStream() const { return overlap ? (gridOnC ? streamC : streamB) : stream; }
    // BDPT: the occlusion tests of bounce b's connections (stream B) run while `stream` traces and shades bounce b+1
    cudaEvent_t evVcmShade = nullptr, evVcmConn[2] = {nullptr, nullptr};
    bool hpFreePending[2] = {false, false}, gridDonePending = false, workOnB = false;
    cudaStream_t sB() const { return overlap ? streamB : stream; }
    void joinStreams();                   // later work on `stream` is ordered after everything queued on stream B
    // photon pass
    float4* phPower = nullptr;            // power.xyz, weight   (PhotonPRD.h:11-12)
    unsigned* phState = nullptr;          // depth | numStored << 8   (= phStateBuf[parity])
    PhotonBuf photons{};                  // 4*P slots                 (= photonsBuf[parity])
    // double-buffered: the photon pass of iteration i+1 fills one set while the grid build of iteration i still reads the other.
    // With one set stream A waited 0.15-0.2 ms per iteration for the build, which runs slowly next to the persistent trace kernels
    // (timeline in profiles/); 155 MB more of 180 GB.
    unsigned* phStateBuf[2] = {nullptr, nullptr};
    PhotonBuf photonsBuf[2]{};
    // photon grid
    unsigned* gridKeys[2] = {nullptr, nullptr};
    unsigned* gridVals[2] = {nullptr, nullptr};
    unsigned* gridHist = nullptr;         // PHOTON_GRID_MAX_SIZE + 2 entries: histogram, then offsets in place
    unsigned* sortTmp = nullptr;
    unsigned* scanTmp = nullptr;
    float* gridBBox = nullptr;            // device: 6 ordered-int floats + valid count
    float* gridBBoxHost = nullptr;        // pinned
    float* gridBBoxHostDev = nullptr;     // the same memory as the device addresses it (photon_bbox_publish_kernel stores into it)
    bool bboxZeroCopy = true;             // OPR_BBOX_ZEROCOPY=0: the seven scalars travel by cudaMemcpyAsync instead (queues behind an image read-back)
    PhotonBuf sorted{};                   // compacted + sorted photons
    unsigned* sortedKeys = nullptr;       // alias of gridKeys[x] after the sort
    GridParams grid{};                    // host copy of the latest iteration's scalars: valid after the stream that built the grid was joined and synchronised
    // Grid scalars without a host wait: the photon bounding box lands in pinned memory, a host function enqueued in the grid stream
    // (cudaLaunchHostFunc) computes cell size / grid size / origin in fp32 exactly like the reference (compute_grid_params: powf, ceilf —
    // which is what keeps the cell keys bit-exact), and a 40-byte copy hands them to the kernels.  Double-buffered by iteration parity.
    struct GridJob { Renderer* self; const float* bboxHost; GridParams* paramsHost; unsigned parity; };
    GridJob gridJob[2]{};
    GridParams* gridParamsHost = nullptr;  // pinned, 2 entries
    GridParams* gridParamsDev = nullptr;   // device, 2 entries
    bool asyncGrid = false;               // OPR_ASYNC_GRID=1: no host wait (below).  Measured 2.7 % SLOWER than the blocking wait (4.39 vs 4.28 ms per iteration; the
                                          // first iterations 5.42 vs 5.23): the wait is hidden behind the direct-light kernels queued before it, while the host function
                                          // adds the driver thread's wake-up latency to stream B and the cell count unknown on the host costs a full-size memset, scan
                                          // and 20-bit sort.  Off by default; the host then waits for the bounding box and computes the scalars itself.
    bool gridTooLarge = false;            // set by the host function: "Too many cells" is then thrown by the next call
    const GridParams& latestGrid();       // joins + synchronises, then returns the scalars of the latest grid build
    // radiance
    // direct lighting: per-pixel light samples + compacted shadow-ray queue (4 slots per pixel)
    float4* sampleFactor = nullptr;       // un-occluded light factor of the four samples; occluded ones are cleared by the trace
    uint2* sampleLight = nullptr;         // light index of the four samples, 16 bits each
    float4* sampleFactorBuf[2] = {nullptr, nullptr};
    uint2* sampleLightBuf[2] = {nullptr, nullptr};
    float4* shadowD = nullptr;            // direction.xyz, tmax
    unsigned* shadowIdx = nullptr;        // 4 * pixel + sample
    float4* shadowD2 = nullptr; unsigned* shadowIdx2 = nullptr;   // the same for the late direct samples
    cudaEvent_t bboxEvent = nullptr;      // photon bbox + count landed in pinned host memory
    float* indirect = nullptr;            // float3 per pixel
    float* direct = nullptr;
    float* output = nullptr;              // running sum, float3 per pixel == reference outputBuffer
    bool keepDebugBuffers = true;

    // multi-GPU (group.cu): this rank's NCCL communicator and the scratch buffer a reduce lands in on the root
    void* ncclComm = nullptr; int commRank = 0, commWorld = 1;
    float* reduced = nullptr; size_t reducedCount = 0;
    // asynchronous read-back (opr_get_output_buffer_async): device snapshots of the running sum, copied out on `streamD`
    cudaStream_t streamD = nullptr;
    float* outputSnap[2] = {nullptr, nullptr};
    cudaEvent_t evOutputReady = nullptr;            // the last iteration's final write to `output` (on outputStream)
    cudaEvent_t evSnapTaken[2] = {nullptr, nullptr}, evSnapCopied[2] = {nullptr, nullptr};
    bool snapTakenPending = false, snapInFlight[2] = {false, false};
    unsigned snapNext = 0;
    bool outputOnB = false;                         // the last iteration finished `output` on stream B (PPM) rather than `stream`
    void getOutputBufferAsync(void* dstHost);
    void waitOutput();
    void waitSnapshotBeforeOutputWrite(cudaStream_t st);   // WAR: the next writer of `output` waits for the device snapshot

    // VCM state lives in vcm.cu
    struct VcmState* vcm = nullptr;
    int debugLightVertexCounts(unsigned* dstHost);
    bool debugConnectionCounts = false;   // opr_debug_set_connection_counts (BDPT connection counts, PATH_TRACING path signatures)
    unsigned* ptSignature = nullptr;      // per pixel: hash of the discrete decisions of the last PATH_TRACING iteration's path
    int debugConnectionCountsGet(unsigned* dstHost);
    int debugPathSignaturesGet(unsigned* dstHost);
    long long debugLightVertexPositions(float* dstHost, size_t capacityVerts);
    bool lightVertexCountEstimated = false;

    // statistics / timing
    unsigned long long stats[STAT_NUM] = {0};
    unsigned long long* stats64 = nullptr;   // device: the per-iteration 32-bit counters are folded in here after every iteration
    unsigned long long launches = 0;
    bool timing = false, timeline = false;
    std::vector<std::pair<std::string, float>> lastTimings;
    struct TimelineEntry { const char* name; char stream; cudaEvent_t a, b; };
    std::vector<TimelineEntry> timelineEvents;      // timeline mode: every pass since the mode was switched on
    std::string lastError;

    ~Renderer();
    void initialize(int cudaOrdinal);
    void initScene(const opr_scene_desc* d);
    void resizeBuffers(unsigned w, unsigned h);
    void freeBuffers();
    void renderNextIteration(unsigned long long iterationNumber, unsigned long long localIterationNumber, float PPMRadius,
                             bool createOutput, const opr_request_details& details);
    void renderPPM(unsigned long long iterationNumber, unsigned localIterationNumber, float PPMRadius, const opr_request_details& details);
    void renderVCM(unsigned long long iterationNumber, unsigned localIterationNumber, float PPMRadius, const opr_request_details& details);
    void photonBBoxAsync(cudaStream_t st);
    void buildPhotonGrid();
    void getOutputBuffer(void* dst);
};

// NVTX range with the reference's pass names (RenderEngine/renderer/helpers/nsight.h:119-127; call sites OptixRenderer.cpp:520,
// 562,574,611,624,656,663,669,782,791).  The ranges bracket the host-side ENQUEUE of a pass — the kernels run asynchronously on the
// renderer's streams — so a timeline shows them next to the kernels they queued.  Header-only NVTX v3: a no-op without a tool.
struct NvtxRange {
    explicit NvtxRange(const char* message);
    ~NvtxRange();
};

// CUDA-event timer around one pass (enabled by opr_set_timing).  Two modes:
//   timing (opr_set_timing 1): every pass is measured ALONE — the device is synchronised before each one, so the stream pipeline is off;
//   timeline (opr_set_timing 2): events are recorded on the pass's own stream without any synchronisation and read after the
//     iterations (opr_get_timeline): start / end of every pass on the device's clock while the three streams overlap as in production.
//     The start event of a pass is reached when the stream's earlier work is done, so end - start includes waits on other streams
//     and the slowdown from the kernels that share the SMs.
struct ScopedTimer {
    Renderer& r; const char* name; cudaStream_t st; cudaEvent_t a = nullptr, b = nullptr;
    ScopedTimer(Renderer& r_, const char* n, cudaStream_t st_ = nullptr) : r(r_), name(n), st(st_ ? st_ : r_.stream)
    {
        if (!r.timing && !r.timeline) return;
        if (r.timeline && r.timelineEvents.size() >= 65536) return;     // a forgotten timeline mode must not grow without bound
        if (r.timing) cudaDeviceSynchronize();    // timing mode serialises the streams: each pass is measured alone
        cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, st);
    }
    ~ScopedTimer()
    {
        if (!a) return;
        cudaEventRecord(b, st);
        if (r.timeline) { r.timelineEvents.push_back(Renderer::TimelineEntry{name, st == r.stream ? 'A' : (st == r.streamB ? 'B' : 'C'), a, b}); return; }
        cudaEventSynchronize(b);
        float ms = 0; cudaEventElapsedTime(&ms, a, b);
        r.lastTimings.emplace_back(name, ms);
        cudaEventDestroy(a); cudaEventDestroy(b);
    }
};

// ---- kernel launchers (ppm.cu, trace.cu, grid.cu) -------------------------------------------------------
struct LaunchCtx { cudaStream_t st; unsigned long long* launches; };

void launch_stats_fold(unsigned* counters32, unsigned long long* totals64, LaunchCtx lc);
void launch_rng_init(const RngBuf& rng, unsigned seed, size_t n, LaunchCtx lc);
// `quantized`: traverse the 64-byte quantised nodes (incoherent rays) instead of the 128-byte fp32 nodes (coherent rays): scene.cuh load_node
void launch_trace_closest(const SceneView& sc, const RayQueues& q, int cur, LaunchCtx lc, bool quantized = false);
void launch_trace_closest_bounce(const SceneView& sc, const RayQueues& q, int bounce, LaunchCtx lc, bool quantized = false, bool wholeTiles = false);
void launch_trace_closest_raw(const SceneView& sc, const float4* rayO, const float4* rayD, unsigned n, float4* hits, LaunchCtx lc, bool quantized = false);
void launch_trace_any_raw(const SceneView& sc, const float4* rayO, const float4* rayD, unsigned n, int* occluded, LaunchCtx lc, bool quantized = false);

void launch_ppm_camera_gen(const opr_camera& cam, unsigned w, unsigned h, unsigned rsW, const RngBuf& rng, const HitpointBuf& hp,
                           const RayQueues& q, LaunchCtx lc);
void launch_ppm_camera_shade(const SceneView& sc, unsigned w, unsigned rsW, const RngBuf& rng, const HitpointBuf& hp,
                             const RayQueues& q, int bounce, bool lastWave, LaunchCtx lc);
void launch_ppm_camera_tail(const SceneView& sc, unsigned w, unsigned rsW, const RngBuf& rng, const HitpointBuf& hp, const RayQueues& q, LaunchCtx lc, bool quantized = false);
void launch_ppm_photon_late(const SceneView& sc, unsigned pw, unsigned rsW, const RngBuf& rng, float4* phPower, unsigned* phState, const PhotonBuf& photons,
                            const RayQueues& q, LaunchCtx lc, bool quantized = false);
void launch_ppm_photon_gen(const SceneView& sc, unsigned pw, unsigned ph, unsigned rsW, const RngBuf& rng, float4* phPower,
                           unsigned* phState, const RayQueues& q, unsigned w, unsigned h, LaunchCtx lc, bool feedsTail = false);
void launch_ppm_photon_shade(const SceneView& sc, unsigned pw, unsigned rsW, const RngBuf& rng, float4* phPower, unsigned* phState,
                             const PhotonBuf& photons, const RayQueues& q, int bounce, LaunchCtx lc, bool lastWave = false);
void launch_ppm_photon_tail(const SceneView& sc, unsigned pw, unsigned rsW, const RngBuf& rng, float4* phPower, unsigned* phState,
                            const PhotonBuf& photons, const RayQueues& q, int bounce, LaunchCtx lc, bool quantized = false);
void launch_ppm_direct_gen(const SceneView& sc, unsigned w, unsigned h, unsigned rsW, const RngBuf& rng, const HitpointBuf& hp, float4* sampleFactor,
                           uint2* sampleLight, float4* shadowD, unsigned* shadowIdx, const RayQueues& q, unsigned pw, unsigned ph, int part, LaunchCtx lc);
void launch_trace_shadow_queue(const SceneView& sc, const float4* hitPos, const float4* shadowD, const unsigned* shadowIdx, const unsigned* count,
                               float* sampleFactor, unsigned* fetchCursor, unsigned* stat, LaunchCtx lc, bool quantized = false);
void launch_pt_generate_ray(const SceneView& sc, const opr_camera& cam, unsigned w, unsigned h, unsigned rsW, const RngBuf& rng, const HitpointBuf& hp,
                            float* output, unsigned localIteration, unsigned* stats, LaunchCtx lc, unsigned* signature = nullptr);
void launch_trace_connect_queue(const SceneView& sc, const float4* connO, const float4* connD, const float4* connC, const unsigned* count, float* target,
                                unsigned stride, unsigned* fetchCursor, unsigned* stat, LaunchCtx lc, unsigned blocksPerSM = 0, unsigned* dbgCount = nullptr,
                                bool quantized = false);
void launch_ppm_resolve_output(const SceneView& sc, unsigned n, const HitpointBuf& hp, const float4* sampleFactor, const uint2* sampleLight,
                               const float* indirect, float* direct, float* output, unsigned localIteration, LaunchCtx lc);

void launch_photon_bbox(const PhotonBuf& photons, const unsigned* phState, unsigned numThreads, float* bbox, LaunchCtx lc, float* hostOutMapped = nullptr);
// the grid scalars reach these kernels through a device pointer: they are computed by a host function IN the stream (renderer.cu)
void launch_grid_keys(const PhotonBuf& photons, const unsigned* phState, unsigned numSlots, const GridParams* g, unsigned* keys, unsigned* hist, LaunchCtx lc);
void launch_photon_reorder(const PhotonBuf& src, const unsigned* sortedSlots, unsigned maxValid, const GridParams* g, const PhotonBuf& dst, LaunchCtx lc);
void launch_gather(const HitpointBuf& hp, unsigned numPixels, unsigned width, unsigned height, const PhotonBuf& sorted, const unsigned* offsets, const GridParams* g,
                   float ppmRadius, float ppmRadiusSquared, float emittedPhotons, float* indirect, unsigned* tileCounter, LaunchCtx lc, unsigned blocksPerSM = 12, unsigned tilesPerBlock = 0);

// host-side restatement of createUniformGridPhotonMap's scalar part (OptixRenderer_SpatialHash.cu:209-250)
bool compute_grid_params(const float lo[3], const float hi[3], GridParams& g);

} // namespace opr

// the C-ABI handle (capi.cu, group.cu)
struct opr_renderer { opr::Renderer r; };
