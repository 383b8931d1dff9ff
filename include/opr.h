/*
 * opr.h — the C-ABI drop-in boundary of the B200 light-transport path.
 *
 * Every entry point replaces one member of the reference's renderer-facing C++ surface
 * (class OptixRenderer, RenderEngine/renderer/OptixRenderer.h:21-44, exported from RenderEngine.dll
 * through RenderEngine/render_engine_export_api.h:8-16).  Plain pointers and sizes only; no exceptions
 * cross this boundary: functions return 0 on success and non-zero on failure, with the message (the
 * text the reference would have thrown as std::exception) available from opr_last_error().
 * One host thread per renderer object, as in the reference (Standalone/StandaloneApplication.cpp:23-26).
 * There is no CPU fallback: without a CUDA device opr_initialize fails.
 */
#ifndef OPR_H
#define OPR_H

#include "opr_scene.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct opr_renderer opr_renderer;

/* message of the last failure on the calling thread */
const char* opr_last_error(void);

/* OptixRenderer::OptixRenderer() — OptixRenderer.cpp:79-104.  cfg may be NULL (reference constants). */
int opr_create(opr_renderer** out, const opr_config* cfg);
/* OptixRenderer::~OptixRenderer() — OptixRenderer.cpp:106-111 */
void opr_destroy(opr_renderer* r);
/* OptixRenderer::initialize(const ComputeDevice&) — OptixRenderer.cpp:113-400; only getDeviceId() is used (:416) */
int opr_initialize(opr_renderer* r, int cuda_device_ordinal);
/* OptixRenderer::initScene(IScene&) — OptixRenderer.cpp:436-485 (uploads geometry, builds the BVH) */
int opr_init_scene(opr_renderer* r, const opr_scene_desc* scene);
/* OptixRenderer::renderNextIteration(iterationNumber, localIterationNumber, PPMRadius, createOutput, details)
 * — OptixRenderer.cpp:507-821 */
int opr_render_next_iteration(opr_renderer* r, uint64_t iterationNumber, uint64_t localIterationNumber, float PPMRadius,
                              int createOutput, const opr_request_details* details);
/* OptixRenderer::getOutputBuffer(void*) — OptixRenderer.cpp:860-865: copies W*H*3 floats (the un-normalised
 * running sum over localIterationNumber = 0..k) into caller-owned host memory */
int opr_get_output_buffer(opr_renderer* r, void* dst);
/* OptixRenderer::getWidth/getHeight/getScreenBufferSizeBytes — OptixRenderer.cpp:850-870 */
uint32_t opr_get_width(opr_renderer* r);
uint32_t opr_get_height(opr_renderer* r);
uint32_t opr_get_screen_buffer_size_bytes(opr_renderer* r);

/* Multi-GPU (replaces Client/ + Server/ TCP distribution): the running-sum buffer stays in
 * HBM; the caller reduces it across ranks (NCCL) through this device pointer (W*H*3 floats). */
void* opr_get_output_device_ptr(opr_renderer* r);
int opr_synchronize(opr_renderer* r);
/* The renderer overlaps the photon-grid build / gather / output of iteration i (internal low-priority stream; BDPT: the connection occlusion tests of a bounce) with the path
 * passes of iteration i+1.  opr_join_streams orders everything queued so far before later work on opr_get_stream(r);
 * opr_get_output_buffer, opr_get_output_device_ptr, opr_synchronize and opr_debug_get_buffer join implicitly. */
int opr_join_streams(opr_renderer* r);
/* the CUDA stream (cudaStream_t) every kernel of r is launched on — for event timing and for ordering the NCCL reduce */
void* opr_get_stream(opr_renderer* r);
/* Closed form of the driver's radius schedule r_{i+1}^2 = r_i^2 (i+alpha)/(i+1)
 * (Standalone/StandaloneRenderManager.cpp:105-107, Client/DistributedApplication.cpp:105-113), so that rank g
 * can render global iterations g, g+G, ... without exchanging radii. */
double opr_ppm_radius_for_iteration(double initialRadius, uint64_t iteration, double alpha);

/* ---- asynchronous read-back (Standalone/StandaloneRenderManager.cpp:114-128 copies the output after EVERY iteration) --------
 * opr_get_output_buffer_async snapshots the running sum as of the last opr_render_next_iteration (device copy ordered behind that
 * iteration only) and streams it to dstHost on the renderer's copy stream while the next iterations render; it returns at once.
 * dstHost must stay valid until opr_wait_output(r) returns (pinned memory makes the copy truly asynchronous).  Up to two
 * read-backs may be in flight; a third call first waits for the oldest.  The synchronous opr_get_output_buffer is unchanged. */
int opr_get_output_buffer_async(opr_renderer* r, void* dstHost);
int opr_wait_output(opr_renderer* r);

/* ---- multi-GPU behind the C ABI: iteration partition + one NCCL reduce ------------------------------------------------------
 * Replaces the client's hand-out loop (Client/DistributedApplication.cpp:96-122: the next global iteration number and its
 * radius go to whichever server is free) and the servers' result packets (Server/server/RenderServerRenderer.cpp:154-175,
 * summed by the client): global iteration i is rendered by rank i % G as its localIterationNumber i / G with the closed-form
 * radius opr_ppm_radius_for_iteration(i); the W*H*3-float running sums are combined by ONE ncclReduce(sum) over NVLink into a
 * scratch buffer on the root (the running sums themselves are never a reduce target, so the call can be repeated).
 * NCCL is loaded at run time (libnccl.so.2); without it these functions fail with a message, nothing else is affected. */
/* seed of rank `rank`: disjoint ranges of the per-state seeds `seed + i` (i < 2^28 states) for up to 16 ranks */
uint32_t opr_group_rank_seed(uint32_t baseSeed, int rank);
/* (a) one process per GPU (torchrun, MPI): rank 0 makes the 128-byte NCCL id, the launcher transports it, every rank joins */
int opr_comm_get_unique_id(void* id128);
int opr_comm_init_rank(opr_renderer* r, const void* id128, int rank, int world);
int opr_comm_destroy(opr_renderer* r);
/* ncclReduce(sum) of every rank's running sum onto `root`, queued on opr_get_stream(r) behind everything rendered so far.
 * dstHost (root only, may be NULL) receives W*H*3 floats and makes the call synchronous; opr_get_reduced_device_ptr(r)
 * is the device copy of the total on the root. */
int opr_reduce_output(opr_renderer* r, int root, void* dstHost);
void* opr_get_reduced_device_ptr(opr_renderer* r);
/* (b) one process, G GPUs: a group owns one renderer per device and one host thread per renderer */
typedef struct opr_group opr_group;
int opr_group_create(opr_group** out, int numDevices, const int* cudaOrdinals, const opr_config* cfg);
void opr_group_destroy(opr_group* g);
int opr_group_size(opr_group* g);
opr_renderer* opr_group_renderer(opr_group* g, int rank);
int opr_group_init_scene(opr_group* g, const opr_scene_desc* scene);
/* renders the global iterations [first, first + count) — all ranks in parallel — with radii from (initialRadius, alpha) */
int opr_group_render_next_iterations(opr_group* g, uint64_t firstIteration, uint64_t count, double initialRadius, double alpha,
                                     const opr_request_details* details);
/* sum of all ranks' running sums -> dstHost (W*H*3 floats) through rank 0 */
int opr_group_reduce_output(opr_group* g, void* dstHost);

/* counters since the last resize: [0] closest-hit rays, [1] any-hit rays, [2] photons stored, [3] kernel launches */
int opr_get_stats(opr_renderer* r, uint64_t* out, int n);
/* per-pass CUDA-event timings of the last iteration (opr_set_timing(r,1) first); returns the number written.
 * opr_set_timing(r, 2) = timeline mode: no synchronisation, the passes of all following iterations are bracketed by events on their own
 * streams; opr_get_timeline synchronises and returns start / end (ms since the first pass) and the stream ('A' path passes, 'B'
 * grid + gather + output, 'C' tails) of every pass since — the overlapped schedule as the device ran it.  Returns the count (-1: error). */
int opr_get_timeline(opr_renderer* r, float* startMs, float* endMs, char* streams, int capacity);
const char* opr_get_timeline_name(opr_renderer* r, int index);
int opr_set_timing(opr_renderer* r, int enabled);
int opr_get_timings_ms(opr_renderer* r, float* ms, int capacity);
const char* opr_get_timing_name(opr_renderer* r, int index);

/* Intermediate buffers in the REFERENCE's layouts for stage-by-stage parity tests.  name is one of
 * "hitpoints" (Hitpoint.h, 64 B x W*H), "photons" (Photon.h, 36 B x 4P, as emitted), "sorted_photons"
 * (36 B x numValid), "photon_keys" (u32 x 4P, sorted), "offsets" (u32 x numCells+1), "indirect", "direct"
 * (float3 x W*H), "random_states" ({d,v0..v4} u32 x 6 per state), "grid" (cellSize, gridSize[3], origo[3],
 * numValid as 8 x 4 bytes), "photon_path_lengths" (u32 x P), "light_vertex_counts" (u32 x W*H), "queue_counts" (34 x u32: ray-queue
 * capacity, shadow rays queued, camera-pass and photon-pass ray counts per bounce of the last PPM iteration).
 * Parity diagnostics (they let a test tell WHICH discrete decision made a pixel differ from the oracle):
 * "direct_sample_mask" (u8 x W*H: bit s = shadow sample s of DirectRadianceEstimation.cu:54-70 was unoccluded and non-zero),
 * "light_vertex_positions" (u32 count + 12 B pad, then {u32 path, float3 hitPoint} per stored light vertex of the last light
 * pass), "connection_counts" (u32 x W*H: contributions that passed isOccluded, vcm.h:43-59, per target pixel, last BDPT
 * iteration; needs opr_debug_set_connection_counts(r, 1) before the iteration), "sample_factors" (float4 x W*H) and "sample_lights" (4 x u16 x W*H): factor and light index of the four direct-light samples of
 * DirectRadianceEstimation.cu:54-70 per pixel (0 = not cast or occluded), "path_signatures" (u32 x 2*W*H, same switch: per BDPT light subpath, then per camera subpath, a hash of the material kinds it met),
 * "pt_signature" (u32 x W*H, same switch: hash over
 * the segments of the last PATH_TRACING iteration's path of (hit flags, light sample contributed, ended by Russian roulette) —
 * RayGeneratorPT.cu:84-127 draws those decisions from a stream whose final state is discarded, so the RNG state cannot show them). */
int opr_debug_get_buffer(opr_renderer* r, const char* name, void* dst, uint64_t bytes);
int opr_debug_set_connection_counts(opr_renderer* r, int enabled);

/* Out-of-bounds write detection without a sanitizer: after opr_debug_set_guards(1) every device allocation of the library
 * carries a 4 KiB band of a known byte pattern on both sides; opr_debug_check_guards synchronises the device, reads every band
 * back and returns the number of allocations whose band was modified — live ones, plus those freed since the guards were switched
 * on, which are checked as they die (-1: nothing guarded has been checked, -2: CUDA error).  *checkedBlocks (may be NULL)
 * receives the number of allocations that count covers. */
void opr_debug_set_guards(int enabled);
int opr_debug_check_guards(int* checkedBlocks);

/* ---- stand-alone kernels of the path (host buffers in, host buffers out) ------------------------------ */
/* initRandomStateBuffer (OptixRenderer_SpatialHash.cu:310-317): states6[6*i..] = XORWOW state of seed+first+i */
int opr_rng_init(uint32_t* states6, uint32_t seed, uint32_t first, uint32_t count);
/* n draws of getRandomUniformFloat (helpers/random.h:65-69) from one state (updated in place) */
int opr_rng_uniform(uint32_t* state6, float* out, uint32_t n);
/* createUniformGridPhotonMap (OptixRenderer_SpatialHash.cu:209-282) on n photons in the reference's AoS layout
 * (36 B each).  Outputs: sortedPhotons (36 B x n, first numValid entries defined), keys (u32 x n, sorted),
 * offsets (u32 x numCells+1, capacity 10^6+2).  Returns 2 if the grid would exceed PHOTON_GRID_MAX_SIZE. */
int opr_build_photon_grid(const float* photonsAoS, uint32_t n, float* sortedPhotonsAoS, uint32_t* keys, uint32_t* offsets,
                          float* cellSize, uint32_t* gridSize3, float* origo3, uint32_t* numValid);
/* IndirectRadianceEstimation::kernel (ppm/IndirectRadianceEstimation.cu:69-227) on caller-supplied hitpoints
 * (64 B x numPixels) and a sorted photon list + offset table as produced by opr_build_photon_grid */
int opr_gather(uint32_t numPixels, const float* hitpointsAoS, uint32_t numPhotons, const float* sortedPhotonsAoS, const uint32_t* offsets,
               float cellSize, const uint32_t* gridSize3, const float* origo3, float ppmRadius, float emittedPhotons, float* indirectOut);
/* closest-hit / any-hit queries against the scene loaded in r.  rays8: n x (o.xyz, tmin, d.xyz, tmax) */
int opr_trace_closest(opr_renderer* r, const float* rays8, uint32_t n, float* outT, int32_t* outMaterial, float* outGeometricNormal3);
int opr_trace_any(opr_renderer* r, const float* rays8, uint32_t n, int32_t* outOccluded);
/* The BVH exists in two node formats (128-byte fp32 nodes for coherent launches, 64-byte quantised nodes for incoherent ones); this
 * switch makes the two queries above traverse the quantised one, so that both are checked ray by ray against the oracle. */
int opr_debug_set_trace_nodes(opr_renderer* r, int quantized);
/* Pixel index of queue slot `slot` of the primary-ray and direct-light passes of a width x height image: 32 consecutive slots cover an
 * 8 x 4-pixel tile when the image divides into tiles, else slot == pixel.  Host arithmetic, the function the kernels call (no reference
 * counterpart: OptiX chooses its own launch-index-to-thread mapping).  Exposed so that the mapping can be checked to be a permutation. */
uint32_t opr_debug_slot_pixel(uint32_t slot, uint32_t width, uint32_t height);

/* ---- host-side pieces of the surface -------------------------------------------------------------------- */
/* Camera::setup / setAspectRatio (renderer/Camera.cpp:333-345, :294-318) */
void opr_camera_setup(opr_camera* c);
void opr_camera_set_aspect_ratio(opr_camera* c, float ratio);
/* Light::Light area / point constructors (renderer/Light.cpp:14-40) */
void opr_make_area_light(opr_light* out, const float* power3, const float* position3, const float* v1, const float* v2);
void opr_make_point_light(opr_light* out, const float* power3, const float* position3);
/* IScene::getSceneInitialPPMRadiusEstimate (scene/IScene.cpp:23-31) */
float opr_initial_ppm_radius(const float* aabbMin3, const float* aabbMax3);
/* Checks every count, pointer and index of a scene description that the renderer will follow (material kinds, material /
 * texture / vertex indices, light types).  0 = sound; 1 = not, with the reason in msg.  opr_init_scene runs it first and fails
 * with the same text. */
int opr_scene_desc_check(const opr_scene_desc* scene, char* msg, int capacity);
/* Built-in scenes: "Cornell" (scene/Cornell.cpp) and "CornellSmall" (scene/CornellSmall.cpp; config = Config bits,
 * 0 = Default).  The returned handle owns the description. */
void* opr_builtin_scene_create(const char* name, uint32_t config);
void opr_builtin_scene_destroy(void* scene);
const opr_scene_desc* opr_builtin_scene_desc(void* scene);
void opr_builtin_scene_camera(void* scene, opr_camera* out);

/* ---- I/O either side of the path (host only) -------------------------------------------------------------- */
/* message of the last failure of the functions below on the calling thread */
const char* opr_io_last_error(void);
/* display normalisation of the running sum: pow(color / previewedIterations, 1/gamma) * 255 as 8-bit BGR
 * (fragment shader Gui/gui/RenderWidget.cpp:187-199; per-pixel loop of saveImageAsBMP :297-315) */
int opr_display_bgr8(const float* sumRGB, uint32_t width, uint32_t height, float previewedIterations, float gamma, uint8_t* outBGR);
/* RenderWidget::saveImageAsBMP (Gui/gui/RenderWidget.cpp:268-317): 24-bit BMP, rows in buffer order, no row padding */
int opr_save_bmp(const char* fileName, const float* sumRGB, uint32_t width, uint32_t height, float previewedIterations, float gamma);
/* Image::Image (RenderEngine/util/Image.cpp:17-140): texel array of a texture file as RGBA8 — uncompressed TGA exactly as the
 * reference reads it (BGR->RGB, no flip, alpha = 1), 24/32-bit BMP.  Two calls: size, then pixels into caller memory. */
int opr_image_info(const char* path, int32_t* width, int32_t* height);
int opr_image_read(const char* path, uint8_t* rgba, uint64_t capacityBytes);

/* Scene::createFromFile (RenderEngine/scene/Scene.cpp:73-176) for COLLADA 1.4 (.dae) files: geometry with node transforms
 * baked, material classification (Scene.cpp:178-265: emissive -> textured (+normal map) -> refractive index > 1 glass ->
 * reflective mirror -> diffuse), point / spot lights (:267-285), quad emitters (:287-311), scene AABB (:147-160), default
 * camera (:469-485).  The handle owns the description and its arrays.  flags bit 0: list every emitter light once (the
 * reference lists it twice: Scene.cpp:141 and :508).  Returns NULL on failure (message: opr_scene_file_last_error). */
void* opr_scene_file_load(const char* path, int32_t flags);
void opr_scene_file_destroy(void* scene);
const opr_scene_desc* opr_scene_file_desc(void* scene);
int opr_scene_file_camera(void* scene, opr_camera* out);      /* 0 when the file has a camera */
uint32_t opr_scene_file_num_triangles(void* scene);
uint32_t opr_scene_file_num_meshes(void* scene);
const char* opr_scene_file_warnings(void* scene);
const char* opr_scene_file_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
