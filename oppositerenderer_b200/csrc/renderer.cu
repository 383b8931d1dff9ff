// renderer.cu — host orchestration of one iteration (the body of OptixRenderer::renderNextIteration,
// RenderEngine/renderer/OptixRenderer.cpp:507-821) over the wavefront kernels, buffer management
// (resizeBuffers :828-848) and RNG (re)seeding (OptixRenderer_SpatialHash.cu:319-347).
#include "renderer.h"
#include "radix_sort.cuh"
#include <algorithm>
#include <cstring>
#include <cstdlib>
#include <cstdio>
#include <nvtx3/nvToolsExt.h>

namespace opr {

NvtxRange::NvtxRange(const char* message) { nvtxRangePushA(message); }
NvtxRange::~NvtxRange() { nvtxRangePop(); }

void vcm_free(Renderer& r);   // vcm.cu

template <class T> static void dalloc(T*& p, size_t count)
{
    devAlloc((void**)&p, std::max<size_t>(count, 1) * sizeof(T));
}
template <class T> static void dfree(T*& p) { devFree(p); p = nullptr; }

Renderer::~Renderer()
{
    if (device >= 0) cudaSetDevice(device);
    vcm_free(*this);
    freeBuffers();
    dfree(phPower);
    for (int b = 0; b < 2; b++) { dfree(phStateBuf[b]); dfree(photonsBuf[b].a); dfree(photonsBuf[b].b); dfree(photonsBuf[b].c); }
    dfree(sorted.a); dfree(sorted.b); dfree(sorted.c);
    dfree(gridKeys[0]); dfree(gridKeys[1]); dfree(gridVals[0]); dfree(gridVals[1]); dfree(gridHist); dfree(sortTmp); dfree(scanTmp);
    dfree(gridBBox); dfree(stats64); dfree(reduced);
    if (gridBBoxHost) cudaFreeHost(gridBBoxHost);
    if (gridParamsHost) cudaFreeHost(gridParamsHost);
    dfree(gridParamsDev);
    scene.release();
    if (bboxEvent) cudaEventDestroy(bboxEvent);
    for (cudaEvent_t e : {evPhotonsFree[0], evPhotonsFree[1], evGatherDone, evHpFree[0], evHpFree[1], evGridDone, evDirectDone, evJoin, evCameraWaves, evPhotonGen, evTailDone, evPhotonWaves, evVcmShade, evVcmConn[0], evVcmConn[1]}) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : {evOutputReady, evSnapTaken[0], evSnapTaken[1], evSnapCopied[0], evSnapCopied[1]}) if (e) cudaEventDestroy(e);
    if (streamD) cudaStreamDestroy(streamD);
    if (streamB) cudaStreamDestroy(streamB);
    if (streamC) cudaStreamDestroy(streamC);
    if (stream) cudaStreamDestroy(stream);
}

void Renderer::initialize(int cudaOrdinal)
{
    if (initialized) throw std::runtime_error("ERROR: Multiple OptixRenderer::initialize!\n");   // OptixRenderer.cpp:115-118
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        throw std::runtime_error("Did not find a CUDA device: this renderer has no CPU path");
    if (cudaOrdinal < 0 || cudaOrdinal >= count)
        throw std::runtime_error("Did not find OptiX device Number for given device. OptiX may not support this device.");   // :427-431 (message kept)
    device = cudaOrdinal;
    // tuning knobs (measurement only; results do not depend on them)
    if (const char* e = getenv("OPR_CAMERA_WAVES")) cameraWaveBounces = (unsigned)std::max(0, atoi(e));
    if (const char* e = getenv("OPR_PHOTON_WAVES")) photonWaveBounces = (unsigned)std::max(0, atoi(e));
    if (const char* e = getenv("OPR_VCM_LIGHT_WAVES")) vcmLightWaveBounces = (unsigned)std::max(1, atoi(e));
    if (const char* e = getenv("OPR_VCM_CAMERA_WAVES")) vcmCameraWaveBounces = (unsigned)std::max(1, atoi(e));
    OPR_CUDA_CHECK(cudaSetDevice(device));
    // Stream priorities: the RNG-ordered path passes (`stream`, and their tails on stream C) are the critical path of an
    // iteration; the gather / connection work on stream B only has to finish before it is needed again.
    int prLeast = 0, prGreatest = 0;
    OPR_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&prLeast, &prGreatest));
    if (const char* e = getenv("OPR_STREAM_PRIORITY")) streamPriorities = atoi(e) != 0;
    if (!streamPriorities) prLeast = prGreatest = 0;
    OPR_CUDA_CHECK(cudaStreamCreateWithPriority(&stream, cudaStreamNonBlocking, prGreatest));
    OPR_CUDA_CHECK(cudaEventCreateWithFlags(&bboxEvent, cudaEventDisableTiming));
    OPR_CUDA_CHECK(cudaStreamCreateWithPriority(&streamB, cudaStreamNonBlocking, prLeast));
    int prC = prGreatest;
    if (const char* e = getenv("OPR_STREAM_PRIORITY_C")) prC = std::min(prLeast, std::max(prGreatest, prGreatest + atoi(e)));
    OPR_CUDA_CHECK(cudaStreamCreateWithPriority(&streamC, cudaStreamNonBlocking, prC));
    for (cudaEvent_t* e : {&evPhotonsFree[0], &evPhotonsFree[1], &evGatherDone, &evHpFree[0], &evHpFree[1], &evGridDone, &evDirectDone, &evJoin, &evCameraWaves, &evPhotonGen, &evTailDone, &evPhotonWaves, &evVcmShade, &evVcmConn[0], &evVcmConn[1]})
        OPR_CUDA_CHECK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    OPR_CUDA_CHECK(cudaStreamCreateWithPriority(&streamD, cudaStreamNonBlocking, prLeast));
    for (cudaEvent_t* e : {&evOutputReady, &evSnapTaken[0], &evSnapTaken[1], &evSnapCopied[0], &evSnapCopied[1]})
        OPR_CUDA_CHECK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    if (const char* e = getenv("OPR_OVERLAP")) overlap = atoi(e) != 0;
    if (const char* e = getenv("OPR_QNODES")) qnodeMask = (unsigned)atoi(e);
    if (const char* e = getenv("OPR_L2_PERSIST")) l2Persist = atoi(e);
    if (const char* e = getenv("OPR_GRID_STREAM")) gridOnC = (e[0] == 'C' || e[0] == 'c');
    if (const char* e = getenv("OPR_GATHER_BLOCKS_PER_SM")) gatherBlocksPerSM = (unsigned)std::max(1, atoi(e));
    if (const char* e = getenv("OPR_VCM_CONNECT_BLOCKS_PER_SM")) vcmConnectBlocksPerSM = (unsigned)std::max(0, atoi(e));
    if (const char* e = getenv("OPR_GATHER_BLOCKS_PER_SM_ALONE")) gatherBlocksPerSMAlone = (unsigned)std::max(1, atoi(e));
    if (const char* e = getenv("OPR_EARLY_DIRECT")) earlyDirect = atoi(e);
    if (const char* e = getenv("OPR_GATHER_TILES_PER_BLOCK")) gatherTilesPerBlock = (unsigned)std::max(0, atoi(e));

    const size_t P = (size_t)photonLaunchWidth * photonLaunchHeight;
    const size_t N = P * MAX_PHOTON_DEPOSITS;
    dalloc(phPower, P);
    for (int b = 0; b < 2; b++) {
        dalloc(phStateBuf[b], P); dalloc(photonsBuf[b].a, N); dalloc(photonsBuf[b].b, N); dalloc(photonsBuf[b].c, N);
        OPR_CUDA_CHECK(cudaMemsetAsync(phStateBuf[b], 0, P * sizeof(unsigned), stream));
    }
    phState = phStateBuf[0]; photons = photonsBuf[0];
    dalloc(sorted.a, N); dalloc(sorted.b, N); dalloc(sorted.c, N);
    dalloc(gridKeys[0], N); dalloc(gridKeys[1], N); dalloc(gridVals[0], N); dalloc(gridVals[1], N);
    dalloc(gridHist, (size_t)PHOTON_GRID_MAX_SIZE + 2);
    dalloc(sortTmp, rs_hist_words((unsigned)N));
    dalloc(scanTmp, scan_tmp_words(PHOTON_GRID_MAX_SIZE + 2));
    dalloc(gridBBox, 8);
    dalloc(stats64, STAT_NUM);
    OPR_CUDA_CHECK(cudaMemsetAsync(stats64, 0, STAT_NUM * sizeof(unsigned long long), stream));
    OPR_CUDA_CHECK(cudaMallocHost((void**)&gridBBoxHost, 2 * 16 * sizeof(float)));      // two sets of (7 results + 7 initial values): by iteration parity
    if (const char* e = getenv("OPR_BBOX_ZEROCOPY")) bboxZeroCopy = atoi(e) != 0;
    OPR_CUDA_CHECK(cudaHostGetDevicePointer((void**)&gridBBoxHostDev, gridBBoxHost, 0));
    {
        const float arm[8] = {FLT_MAX, FLT_MAX, FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX, 0.f, 0.f};
        OPR_CUDA_CHECK(cudaMemcpy(gridBBox, arm, sizeof arm, cudaMemcpyHostToDevice));
    }
    OPR_CUDA_CHECK(cudaMallocHost((void**)&gridParamsHost, 2 * sizeof(GridParams)));
    memset(gridParamsHost, 0, 2 * sizeof(GridParams));
    dalloc(gridParamsDev, 2);
    if (const char* e = getenv("OPR_ASYNC_GRID")) asyncGrid = atoi(e) != 0;
    initialized = true;
}

void Renderer::initScene(const opr_scene_desc* d)
{
    if (!initialized) throw std::runtime_error("Cannot initialize scene before OptixRenderer.");   // :438-441
    if (!d || d->numLights == 0) throw std::runtime_error("No lights exists in this scene.");     // :443-447
    OPR_CUDA_CHECK(cudaSetDevice(device));
    lightVertexCountEstimated = false;
    sceneLoaded = false;      // DeviceScene::load releases the previous scene first: a failed reload must not leave a half-freed scene renderable
    scene.load(d, stream);
    sceneLoaded = true;
    // L2 residency of the BVH (OPR_L2_PERSIST: 0 off, 1 the two node arrays, 2 nodes + leaf triangles): an access-policy window on the streams
    // that trace, over the traversal block, with as much of the L2 set aside as the device allows.  The iteration's streaming buffers (photon
    // slots, hit points, queues: ~ 400 MB per iteration) otherwise evict the tree between two bounces.
    if (l2Persist > 0 && scene.traversalBlock) {
        int maxPersist = 0, maxWindow = 0;
        OPR_CUDA_CHECK(cudaDeviceGetAttribute(&maxPersist, cudaDevAttrMaxPersistingL2CacheSize, device));
        OPR_CUDA_CHECK(cudaDeviceGetAttribute(&maxWindow, cudaDevAttrMaxAccessPolicyWindowSize, device));
        const size_t want = l2Persist >= 2 ? scene.traversalBytes : scene.traversalNodeBytes;
        const size_t bytes = std::min<size_t>(want, (size_t)maxWindow);
        const size_t setAside = std::min<size_t>(bytes, (size_t)maxPersist);
        if (setAside > 0) {
            OPR_CUDA_CHECK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, setAside));
            cudaStreamAttrValue v{};
            v.accessPolicyWindow.base_ptr = const_cast<void*>(scene.traversalBlock);
            v.accessPolicyWindow.num_bytes = bytes;
            v.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)setAside / (double)bytes);
            v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            for (cudaStream_t s : {stream, streamC, streamB}) OPR_CUDA_CHECK(cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &v));
            if (getenv("OPR_BVH_VERBOSE")) fprintf(stderr, "[opr] L2 window: %.1f MB of the traversal block, %.1f MB set aside (device max %.1f MB)\n", bytes / 1e6, setAside / 1e6, maxPersist / 1e6);
        }
    }
}

void Renderer::joinStreams()
{
    joinedSinceRender = true;
    if (!overlap || !workOnB) return;
    OPR_CUDA_CHECK(cudaEventRecord(evJoin, streamB));
    OPR_CUDA_CHECK(cudaStreamWaitEvent(stream, evJoin, 0));
    workOnB = false;
}

void Renderer::freeBuffers()
{
    dfree(rng.s);
    for (int b = 0; b < 2; b++) { dfree(hpBuf[b].pos); dfree(hpBuf[b].nrm); dfree(hpBuf[b].att); dfree(hpBuf[b].rad); dfree(sampleFactorBuf[b]); dfree(sampleLightBuf[b]); }
    hp = HitpointBuf{}; sampleFactor = nullptr; sampleLight = nullptr;
    hpFreePending[0] = hpFreePending[1] = gridDonePending = false;
    photonsFreePending[0] = photonsFreePending[1] = gatherDonePending = false;
    for (int i = 0; i < 2; i++) { dfree(q.rayO[i]); dfree(q.rayD[i]); dfree(q.rayPath[i]); }
    dfree(q.hits); dfree(q.matQueue); dfree(q.counters);
    dfree(q.tailO); dfree(q.tailD); dfree(q.tailPath); dfree(q.camPending); dfree(q.deferred); dfree(q.photonPending); dfree(q.directDeferred);
    dfree(indirect); dfree(direct); dfree(output); dfree(ptSignature); dfree(outputSnap[0]); dfree(outputSnap[1]);
    snapTakenPending = snapInFlight[0] = snapInFlight[1] = false;
    dfree(shadowD); dfree(shadowIdx); dfree(shadowD2); dfree(shadowIdx2);
    buffersAllocated = false;
}

void Renderer::resizeBuffers(unsigned w, unsigned h)
{   // OptixRenderer.cpp:828-848
    OPR_CUDA_CHECK(cudaStreamSynchronize(stream));
    if (streamB) OPR_CUDA_CHECK(cudaStreamSynchronize(streamB));
    if (streamC) OPR_CUDA_CHECK(cudaStreamSynchronize(streamC));
    if (streamD) OPR_CUDA_CHECK(cudaStreamSynchronize(streamD));
    freeBuffers();
    const size_t px = (size_t)w * h;
    rsW = std::max(photonLaunchWidth, w); rsH = std::max(photonLaunchHeight, h);
    const size_t ns = (size_t)rsW * rsH;
    dalloc(rng.s, 2 * ns);
    for (int b = 0; b < 2; b++) {
        dalloc(hpBuf[b].pos, px); dalloc(hpBuf[b].nrm, px); dalloc(hpBuf[b].att, px); dalloc(hpBuf[b].rad, px);
        dalloc(sampleFactorBuf[b], px); dalloc(sampleLightBuf[b], px);
    }
    parity = 0; hp = hpBuf[0]; sampleFactor = sampleFactorBuf[0]; sampleLight = sampleLightBuf[0];
    const size_t cap = std::max(px, (size_t)photonLaunchWidth * photonLaunchHeight);
    q.capacity = (unsigned)cap;
    for (int i = 0; i < 2; i++) { dalloc(q.rayO[i], cap); dalloc(q.rayD[i], cap); dalloc(q.rayPath[i], cap); }
    dalloc(q.hits, cap); dalloc(q.matQueue, cap * NUM_QUEUES); dalloc(q.counters, CTR_ALLOC);
    dalloc(q.tailO, px); dalloc(q.tailD, px); dalloc(q.tailPath, px); dalloc(q.camPending, px);
    dalloc(q.deferred, (size_t)photonLaunchWidth * photonLaunchHeight);
    dalloc(q.photonPending, (size_t)photonLaunchWidth * photonLaunchHeight); dalloc(q.directDeferred, px);
    dalloc(indirect, px * 3); dalloc(direct, px * 3); dalloc(output, px * 3);
    dalloc(shadowD, px * 4); dalloc(shadowIdx, px * 4); dalloc(shadowD2, px * 4); dalloc(shadowIdx2, px * 4);
    OPR_CUDA_CHECK(cudaMemsetAsync(q.counters, 0, CTR_ALLOC * sizeof(unsigned), stream));
    OPR_CUDA_CHECK(cudaMemsetAsync(output, 0, px * 3 * sizeof(float), stream));
    buffersAllocated = true;
    LaunchCtx lc{stream, &launches};
    launch_rng_init(rng, seed, ns, lc);    // initializeRandomStates(): every resize re-seeds (:835)
    width = w; height = h;
    lightVertexCountEstimated = false;      // :847
}

void Renderer::photonBBoxAsync(cudaStream_t st)
{   // getPhotonsBoundingBox, OptixRenderer_SpatialHash.cu:123-128.  The result travels to pinned host memory behind an event, so
    // the direct-lighting kernels queued after it keep the GPU busy while the host computes the grid scalars.
    LaunchCtx lc{st, &launches};
    const unsigned P = photonLaunchWidth * photonLaunchHeight;
    ScopedTimer t(*this, "grid.bbox", st);
    float* h = gridBBoxHost + 16 * (overlap ? parity : 0);
    h[8] = FLT_MAX; h[9] = FLT_MAX; h[10] = FLT_MAX; h[11] = -FLT_MAX; h[12] = -FLT_MAX; h[13] = -FLT_MAX; h[14] = 0.f; h[15] = 0.f;
    if (bboxZeroCopy) {   // armed at allocation and re-armed by the publish kernel; the result is stored into the pinned block by the device
        launch_photon_bbox(photons, phState, P, gridBBox, lc, gridBBoxHostDev + (h - gridBBoxHost));
    } else {
        OPR_CUDA_CHECK(cudaMemcpyAsync(gridBBox, h + 8, 7 * sizeof(float), cudaMemcpyHostToDevice, st));
        launch_photon_bbox(photons, phState, P, gridBBox, lc);
        OPR_CUDA_CHECK(cudaMemcpyAsync(h, gridBBox, 7 * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    OPR_CUDA_CHECK(cudaEventRecord(bboxEvent, st));
}

// Host function in the grid stream: the scalar part of createUniformGridPhotonMap (OptixRenderer_SpatialHash.cu:209-250) on the bounding
// box that the preceding copy left in pinned memory.  Runs on a driver thread; touches only pinned memory and plain host fields.
static void CUDART_CB grid_scalars_host(void* p)
{
    Renderer::GridJob* job = (Renderer::GridJob*)p;
    const float* h = job->bboxHost;
    unsigned numValid; memcpy(&numValid, h + 6, 4);
    float lo[3] = {h[0], h[1], h[2]}, hi[3] = {h[3], h[4], h[5]};
    if (numValid == 0) { lo[0] = lo[1] = lo[2] = hi[0] = hi[1] = hi[2] = 0.f; }   // photons[0] is a cleared slot
    GridParams g;
    if (!compute_grid_params(lo, hi, g)) { job->self->gridTooLarge = true; g.gridSize = make_uint3(1, 1, 1); g.numCells = 1; }   // :244-248, reported by the next call
    g.numValid = numValid;
    *job->paramsHost = g;
    job->self->stats[STAT_PHOTONS_STORED] += numValid;
}

const GridParams& Renderer::latestGrid()
{
    joinStreams();
    OPR_CUDA_CHECK(cudaStreamSynchronize(stream));
    if (asyncGrid) grid = gridParamsHost[overlap ? parity : 0];
    return grid;
}

void Renderer::buildPhotonGrid()
{   // createUniformGridPhotonMap, OptixRenderer_SpatialHash.cu:209-282
    // Stream B by default.  (OPR_GRID_STREAM=C moves it to the greatest-priority stream C: measured, no gain.)
    const cudaStream_t st = gridStream();
    LaunchCtx lc{st, &launches};
    const unsigned P = photonLaunchWidth * photonLaunchHeight;
    const unsigned N = P * MAX_PHOTON_DEPOSITS;
    const unsigned pb = overlap ? parity : 0;
    GridParams* gDev = gridParamsDev + pb;
    OPR_CUDA_CHECK(cudaStreamWaitEvent(st, bboxEvent, 0));   // the build starts once the photon pass of this iteration is complete
    // the previous iteration's gather still reads the sorted photons and the offset table this build overwrites
    if (st != streamB && gatherDonePending) { OPR_CUDA_CHECK(cudaStreamWaitEvent(st, evGatherDone, 0)); gatherDonePending = false; }
    workOnB = true;
    unsigned numCellsBound = PHOTON_GRID_MAX_SIZE;
    if (asyncGrid) {
        // no host wait: the scalars are computed by a host function that the stream runs when the bounding box has landed
        gridJob[pb] = GridJob{this, gridBBoxHost + 16 * pb, gridParamsHost + pb, pb};
        OPR_CUDA_CHECK(cudaLaunchHostFunc(st, grid_scalars_host, &gridJob[pb]));
        OPR_CUDA_CHECK(cudaMemcpyAsync(gDev, gridParamsHost + pb, sizeof(GridParams), cudaMemcpyHostToDevice, st));
    } else {
        OPR_CUDA_CHECK(cudaEventSynchronize(bboxEvent));   // the reference syncs here too (:222)
        const float* h = gridBBoxHost + 16 * pb;
        unsigned numValid; memcpy(&numValid, h + 6, 4);
        float lo[3] = {h[0], h[1], h[2]}, hi[3] = {h[3], h[4], h[5]};
        if (numValid == 0) { lo[0] = lo[1] = lo[2] = hi[0] = hi[1] = hi[2] = 0.f; }   // photons[0] is a cleared slot
        if (!compute_grid_params(lo, hi, grid))
            throw std::runtime_error("Too many cells in SpatialHash.cu, over defined PHOTON_GRID_MAX_SIZE.");   // :246
        grid.numValid = numValid;
        stats[STAT_PHOTONS_STORED] += numValid;
        gridParamsHost[pb] = grid;
        OPR_CUDA_CHECK(cudaMemcpyAsync(gDev, gridParamsHost + pb, sizeof(GridParams), cudaMemcpyHostToDevice, st));
        numCellsBound = grid.numCells;
    }
    {
        ScopedTimer t(*this, "grid.keys", st);
        OPR_CUDA_CHECK(cudaMemsetAsync(gridHist, 0, ((size_t)numCellsBound + 2) * sizeof(unsigned), st));
        launch_grid_keys(photons, phState, N, gDev, gridKeys[0], gridHist, lc);
    }
    bool flipped;
    {
        ScopedTimer t(*this, "grid.sort", st);
        int bits = 1; while ((1ull << bits) <= (unsigned long long)numCellsBound + 1) bits++;      // 20 bits when the cell count is not known on the host
        int nl = 0;
        // the invalid slots (key numCells + 1, 58 % of the 4 M) are dropped by the first pass when their number is known on the host
        const unsigned invalidKey = numCellsBound + 1;
        if (!asyncGrid) flipped = radix_sort_pairs<unsigned>(gridKeys[0], gridVals[0], gridKeys[1], gridVals[1], N, bits, sortTmp, true, st, &nl, &invalidKey, grid.numValid);
        else flipped = radix_sort_pairs<unsigned>(gridKeys[0], gridVals[0], gridKeys[1], gridVals[1], N, bits, sortTmp, true, st, &nl);
        launches += nl;
    }
    sortedKeys = flipped ? gridKeys[1] : gridKeys[0];
    const unsigned* sortedSlots = flipped ? gridVals[1] : gridVals[0];
    {
        ScopedTimer t(*this, "grid.offsets", st);
        exclusive_scan_u32(gridHist, gridHist, numCellsBound + 1, scanTmp, st);
        launches += 3;
    }
    {
        ScopedTimer t(*this, "grid.reorder", st);
        launch_photon_reorder(photons, sortedSlots, N, gDev, sorted, lc);
    }
}

void Renderer::renderPPM(unsigned long long iterationNumber, unsigned localIterationNumber, float PPMRadius, const opr_request_details& details)
{
    LaunchCtx lc{stream, &launches};
    const SceneView& sc = scene.view;
    const cudaStream_t stB = sB();
    // Streams.  `stream` (A, greatest priority) runs what consumes the RNG states, in the reference's order: camera pass, photon pass, direct
    // light samples.  Stream B runs what does not: photon-grid build, gather, sample resolve + output.  B's work for iteration i
    // overlaps A's work for iteration i+1 (hit points and light samples are double-buffered), so the compute-bound gather fills
    // the SMs that the latency-bound tail kernels and launch gaps of the path passes leave idle.
    // A caller that reads the image back after every iteration leaves nothing to overlap with: give the gather the whole GPU then.
    const unsigned gatherBlocks = (overlap && !joinedSinceRender) ? gatherBlocksPerSM : gatherBlocksPerSMAlone;
    const bool early = overlap && (earlyDirect > 0 || (earlyDirect < 0 && joinedSinceRender));
    joinedSinceRender = false;
    parity ^= 1u;
    const unsigned pb = parity;
    hp = hpBuf[pb]; sampleFactor = sampleFactorBuf[pb]; sampleLight = sampleLightBuf[pb];
    photons = photonsBuf[overlap ? pb : 0]; phState = phStateBuf[overlap ? pb : 0];
    if (hpFreePending[pb]) { OPR_CUDA_CHECK(cudaStreamWaitEvent(stream, evHpFree[pb], 0)); hpFreePending[pb] = false; }
    // ---- PPM_RAYTRACE_PASS (:576): wavefront bounces on `stream`; the specular continuations go to the tail queue, which a
    // per-path tail kernel finishes on stream C while the photon wavefronts run (the photon threads whose RNG state those
    // pixels still hold are deferred behind it: ppm_photon_gen_kernel)
    const cudaStream_t stC = sC();
    LaunchCtx lcC{stC, &launches};
    {
        NvtxRange nv("OptixEntryPoint::RAYTRACE_PASS");
        ScopedTimer t(*this, "ppm.camera");
        { ScopedTimer tg(*this, "camera.gen"); launch_ppm_camera_gen(details.camera, width, height, rsW, rng, hp, q, lc); }
        const unsigned waves = std::max(1u, std::min(cameraWaveBounces, MAX_RADIANCE_TRACE_DEPTH + 1));
        for (unsigned bounce = 0; bounce < waves; bounce++) {
            { ScopedTimer t0(*this, bounce == 0 ? "trace.camera0" : "trace.camera1+"); launch_trace_closest_bounce(sc, q, bounce, lc, (qnodeMask & QN_PRIMARY) != 0, bounce == 0); }
            { ScopedTimer t0(*this, bounce == 0 ? "shade.camera0" : "shade.camera1+"); launch_ppm_camera_shade(sc, width, rsW, rng, hp, q, bounce, bounce + 1 == waves, lc); }
        }
        OPR_CUDA_CHECK(cudaEventRecord(evCameraWaves, stream));
    }
    const float ppmRadiusSquared = PPMRadius * PPMRadius;   // :585
    // ---- PPM_PHOTON_PASS (:612)
    {
        NvtxRange nv("OptixEntryPoint::PHOTON_PASS");
        ScopedTimer t(*this, "ppm.photon");
        // the previous iteration's grid build (stream B) must be done reading the photon slots before they are overwritten
        if (photonsFreePending[pb]) { OPR_CUDA_CHECK(cudaStreamWaitEvent(stream, evPhotonsFree[pb], 0)); photonsFreePending[pb] = false; }
        const unsigned waves = std::min(photonWaveBounces, MAX_PHOTON_TRACE_DEPTH + 1);
        { ScopedTimer tg(*this, "photon.gen");
          launch_ppm_photon_gen(sc, photonLaunchWidth, photonLaunchHeight, rsW, rng, phPower, phState, q, width, height, lc, waves == 0); }
        OPR_CUDA_CHECK(cudaEventRecord(evPhotonGen, stream));
        // stream C: camera tail, then the deferred photon threads
        OPR_CUDA_CHECK(cudaStreamWaitEvent(stC, evCameraWaves, 0));
        { ScopedTimer t1(*this, "ppm.camera_tail", stC); launch_ppm_camera_tail(sc, width, rsW, rng, hp, q, lcC, (qnodeMask & QN_CAMERA_TAIL) != 0); }
        OPR_CUDA_CHECK(cudaStreamWaitEvent(stC, evPhotonGen, 0));
        { ScopedTimer t1(*this, "ppm.photon_late", stC); launch_ppm_photon_late(sc, photonLaunchWidth, rsW, rng, phPower, phState, photons, q, lcC, (qnodeMask & QN_PHOTON_TAILS) != 0); }
        static const char* tn[] = {"trace.photon0", "trace.photon1", "trace.photon2", "trace.photon3", "trace.photon4+"};
        static const char* sn[] = {"shade.photon0", "shade.photon1", "shade.photon2", "shade.photon3", "shade.photon4+"};
        for (unsigned bounce = 0; bounce < waves; bounce++) {
            { ScopedTimer t0(*this, tn[std::min(bounce, 4u)]); launch_trace_closest_bounce(sc, q, bounce, lc, (qnodeMask & QN_PHOTON) != 0); }
            { ScopedTimer t0(*this, sn[std::min(bounce, 4u)]);
              launch_ppm_photon_shade(sc, photonLaunchWidth, rsW, rng, phPower, phState, photons, q, bounce, lc, bounce + 1 == waves); }
        }
        if (early) {
            // stream C again: the photons still alive finish their walk there, then the photon bounding box (first step of
            // createUniformGridPhotonMap, :628) goes to the host, while `stream` already draws the light samples of every pixel
            // whose RNG state and hit point are final (all but the pixels behind the two tails)
            OPR_CUDA_CHECK(cudaEventRecord(evPhotonWaves, stream));
            OPR_CUDA_CHECK(cudaStreamWaitEvent(stC, evPhotonWaves, 0));
        }
        if (waves <= MAX_PHOTON_TRACE_DEPTH) {
            ScopedTimer t1(*this, "ppm.photon_tail", early ? stC : stream);
            launch_ppm_photon_tail(sc, photonLaunchWidth, rsW, rng, phPower, phState, photons, q, waves, early ? lcC : lc, (qnodeMask & QN_PHOTON_TAILS) != 0);
        }
        if (early) photonBBoxAsync(stC);
        OPR_CUDA_CHECK(cudaEventRecord(evTailDone, stC));
        if (!early) {
            OPR_CUDA_CHECK(cudaStreamWaitEvent(stream, evTailDone, 0));   // hit points, RNG states and photons of the deferred paths are complete
            photonBBoxAsync(stream);
        }
    }
    // ---- PPM_DIRECT_RADIANCE_ESTIMATION_PASS (:664) moved before the grid build: it needs hit points and RNG states, not the
    // photon map, and fills the GPU while the host waits for the bounding box.  The RNG order per state is unchanged (camera pass
    // -> photon pass -> direct pass): with `early`, a pixel's samples are drawn in the first launch only if nothing before them in
    // that order is still running, else in the late one.
    {
        NvtxRange nv("OptixEntryPoint::PPM_DIRECT_RADIANCE_ESTIMATION_PASS");
        ScopedTimer t(*this, "ppm.direct");
        { ScopedTimer t0(*this, "direct.gen");
          launch_ppm_direct_gen(sc, width, height, rsW, rng, hp, sampleFactor, sampleLight, shadowD, shadowIdx, q, photonLaunchWidth, photonLaunchHeight,
                                early ? 1 : 0, lc); }
        { ScopedTimer t0(*this, "trace.shadow");
          launch_trace_shadow_queue(sc, hp.pos, shadowD, shadowIdx, q.counters + CTR_SHADOW_COUNT, (float*)sampleFactor, q.counters + CTR_FETCH_SHADOW,
                                    q.counters + CTR_STATS + STAT_ANY_RAYS, lc, (qnodeMask & QN_SHADOW) != 0); }
        if (early) {
            OPR_CUDA_CHECK(cudaStreamWaitEvent(stream, evTailDone, 0));   // hit points, RNG states and photons of the tail paths are complete
            ScopedTimer t0(*this, "direct.late");
            launch_ppm_direct_gen(sc, width, height, rsW, rng, hp, sampleFactor, sampleLight, shadowD2, shadowIdx2, q, photonLaunchWidth, photonLaunchHeight, 2, lc);
            launch_trace_shadow_queue(sc, hp.pos, shadowD2, shadowIdx2, q.counters + CTR_SHADOW2_COUNT, (float*)sampleFactor, q.counters + CTR_FETCH_SHADOW2,
                                      q.counters + CTR_STATS + STAT_ANY_RAYS, lc, (qnodeMask & QN_SHADOW) != 0);
        }
        OPR_CUDA_CHECK(cudaEventRecord(evDirectDone, stream));
    }
    // ---- rest of createUniformGridPhotonMap (stream B; the host waits for the bounding box inside)
    { NvtxRange nv("Creating photon map"); buildPhotonGrid(); }
    OPR_CUDA_CHECK(cudaEventRecord(evPhotonsFree[pb], gridStream())); photonsFreePending[pb] = overlap;
    if (gridStream() != stB) {   // the gather (stream B) needs the sorted photons and the offset table
        OPR_CUDA_CHECK(cudaEventRecord(evGridDone, gridStream()));
        OPR_CUDA_CHECK(cudaStreamWaitEvent(stB, evGridDone, 0));
    }
    LaunchCtx lcB{stB, &launches};
    // ---- PPM_INDIRECT_RADIANCE_ESTIMATION_PASS (:657)
    {
        NvtxRange nv("OptixEntryPoint::INDIRECT_RADIANCE_ESTIMATION");
        ScopedTimer t(*this, "ppm.gather", stB);
        launch_gather(hp, width * height, width, height, sorted, gridHist, gridParamsDev + (overlap ? parity : 0), PPMRadius, ppmRadiusSquared,
                      float(photonLaunchWidth * photonLaunchHeight), indirect, q.counters + CTR_TILE, lcB, gatherBlocks,
                      (overlap && gatherBlocks == gatherBlocksPerSM) ? gatherTilesPerBlock : 0u);
    }
    if (gridStream() != stB) { OPR_CUDA_CHECK(cudaEventRecord(evGatherDone, stB)); gatherDonePending = true; }
    // ---- sum of the light samples + PPM_OUTPUT_PASS (:670)
    {
        OPR_CUDA_CHECK(cudaStreamWaitEvent(stB, evDirectDone, 0));
        waitSnapshotBeforeOutputWrite(stB);
        NvtxRange nv("OptixEntryPoint::PPM_OUTPUT_PASS");
        ScopedTimer t(*this, "ppm.resolve_output", stB);
        launch_ppm_resolve_output(sc, width * height, hp, sampleFactor, sampleLight, indirect, keepDebugBuffers ? direct : nullptr, output, localIterationNumber, lcB);
    }
    OPR_CUDA_CHECK(cudaEventRecord(evHpFree[pb], stB)); hpFreePending[pb] = true;
    outputOnB = overlap;
    (void)iterationNumber;
}

void Renderer::renderNextIteration(unsigned long long iterationNumber, unsigned long long localIterationNumber, float PPMRadius,
                                   bool /*createOutput — accepted and ignored, as in the reference*/, const opr_request_details& details)
{
    if (!initialized) throw std::runtime_error("Traced before OptixRenderer was initialized.");   // :513-516
    if (!sceneLoaded) throw std::runtime_error("Traced before a scene was initialized.");
    OPR_CUDA_CHECK(cudaSetDevice(device));
    if (details.width == 0 || details.height == 0) throw std::runtime_error("zero-sized render request");
    if (gridTooLarge) { gridTooLarge = false; throw std::runtime_error("Too many cells in SpatialHash.cu, over defined PHOTON_GRID_MAX_SIZE."); }   // :246, found by the previous iteration's host function
    char rangeName[48];
    snprintf(rangeName, sizeof rangeName, "OptixRenderer::Trace Iteration %llu", iterationNumber);   // OptixRenderer.cpp:518-520
    NvtxRange nvIteration(rangeName);
    lastTimings.clear();
    if (!buffersAllocated || details.width != width || details.height != height) resizeBuffers(details.width, details.height);   // :534-537
    outputOnB = false;
    if (localIterationNumber == 0) {                                                                                             // :540-545
        joinStreams();
        waitSnapshotBeforeOutputWrite(stream);
        OPR_CUDA_CHECK(cudaMemsetAsync(output, 0, (size_t)width * height * 3 * sizeof(float), stream));
    }
    if (details.renderMethod != OPR_PROGRESSIVE_PHOTON_MAPPING) { joinStreams(); waitSnapshotBeforeOutputWrite(stream); }   // BDPT / PT accumulate into `output` on `stream`

    if (details.renderMethod == OPR_PROGRESSIVE_PHOTON_MAPPING)
        renderPPM(iterationNumber, (unsigned)localIterationNumber, PPMRadius, details);
    else if (details.renderMethod == OPR_VCM_BIDIRECTIONAL_PATH_TRACING)
        renderVCM(iterationNumber, (unsigned)localIterationNumber, PPMRadius, details);
    else if (details.renderMethod == OPR_PATH_TRACING) {   // :559-567
        NvtxRange nv("OptixEntryPoint::PT_RAYTRACE_PASS");
        ScopedTimer t(*this, "pt.generateRay");
        LaunchCtx lc{stream, &launches};
        if (debugConnectionCounts && !ptSignature) dalloc(ptSignature, (size_t)width * height);
        launch_pt_generate_ray(scene.view, details.camera, width, height, rsW, rng, hp, output, (unsigned)localIterationNumber,
                               q.counters + CTR_STATS, lc, debugConnectionCounts ? ptSignature : nullptr);
    } else
        throw std::runtime_error("unknown RenderMethod");
    { LaunchCtx lc{stream, &launches}; launch_stats_fold(q.counters + CTR_STATS, stats64, lc); }   // every ray kernel of an iteration runs on `stream`
    OPR_CUDA_CHECK(cudaEventRecord(evOutputReady, outputOnB ? streamB : stream));   // the running sum of this iteration is complete behind this event
    OPR_CUDA_CHECK(cudaGetLastError());
}

void Renderer::getOutputBuffer(void* dst)
{   // OptixRenderer.cpp:860-865: copies W*H*12 bytes into a caller-owned buffer
    if (!buffersAllocated) throw std::runtime_error("getOutputBuffer before the first iteration");
    OPR_CUDA_CHECK(cudaSetDevice(device));
    joinStreams();
    OPR_CUDA_CHECK(cudaMemcpyAsync(dst, output, (size_t)width * height * 3 * sizeof(float), cudaMemcpyDeviceToHost, stream));
    OPR_CUDA_CHECK(cudaStreamSynchronize(stream));
}

// ---- asynchronous read-back --------------------------------------------------------------------------------------------
// The driver loop of the reference copies the output after every iteration (Standalone/StandaloneRenderManager.cpp:114-128) and
// blocks while it does.  Here the copy leaves through a third party: stream D waits for the event behind the iteration's last
// write to the running sum, snapshots the sum device-to-device (25 MB at 1080p: ~10 us) and sends the snapshot to the host, while
// streams A / B / C already render the next iterations.  Only the next WRITER of `output` waits — for the snapshot, not the copy.
void Renderer::waitSnapshotBeforeOutputWrite(cudaStream_t st)
{
    if (!snapTakenPending) return;
    OPR_CUDA_CHECK(cudaStreamWaitEvent(st, evSnapTaken[snapNext ^ 1u], 0));
    snapTakenPending = false;
}

void Renderer::getOutputBufferAsync(void* dstHost)
{
    if (!buffersAllocated) throw std::runtime_error("getOutputBuffer before the first iteration");
    OPR_CUDA_CHECK(cudaSetDevice(device));
    const unsigned s = snapNext;
    const size_t bytes = (size_t)width * height * 3 * sizeof(float);
    if (snapInFlight[s]) { OPR_CUDA_CHECK(cudaEventSynchronize(evSnapCopied[s])); snapInFlight[s] = false; }   // third read-back in flight: wait for the oldest
    if (!outputSnap[s]) dalloc(outputSnap[s], (size_t)width * height * 3);
    OPR_CUDA_CHECK(cudaStreamWaitEvent(streamD, evOutputReady, 0));
    OPR_CUDA_CHECK(cudaMemcpyAsync(outputSnap[s], output, bytes, cudaMemcpyDeviceToDevice, streamD));
    OPR_CUDA_CHECK(cudaEventRecord(evSnapTaken[s], streamD));
    OPR_CUDA_CHECK(cudaMemcpyAsync(dstHost, outputSnap[s], bytes, cudaMemcpyDeviceToHost, streamD));
    OPR_CUDA_CHECK(cudaEventRecord(evSnapCopied[s], streamD));
    snapInFlight[s] = true; snapTakenPending = true;
    snapNext = s ^ 1u;
}

void Renderer::waitOutput()
{
    OPR_CUDA_CHECK(cudaSetDevice(device));
    for (unsigned s = 0; s < 2; s++) if (snapInFlight[s]) { OPR_CUDA_CHECK(cudaEventSynchronize(evSnapCopied[s])); snapInFlight[s] = false; }
}

} // namespace opr
