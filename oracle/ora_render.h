// ORACLE — test infrastructure only (see ora_math.h).  Renderer state restated from
// RenderEngine/renderer/OptixRenderer.{h,cpp}; buffers follow SURVEY.md Appendix C.
#pragma once
#include "ora_core.h"

namespace ora {

// renderer/vcm/LightVertex.h:14-27 is restated in ora_vcm.h; forward-declared container here.
struct LightVertex;

struct Renderer {
    // configuration (runtime versions of OptixRenderer.cpp:44-53 constants)
    uint photonLaunchWidth = 1024, photonLaunchHeight = 1024;
    uint vcmMaxPathLength = 10;
    uint seed = 1645301512u;    // config.h:50

    unsigned long long uid = 0;   // unique per Renderer object (the oracle/_ref build keys its OptiX-style context on it)
    bool initialized = false;
    bool sceneLoaded = false;
    Scene scene;

    uint width = 10, height = 10;           // OptixRenderer.cpp:84-85
    unsigned rngEpoch = 0;                  // bumped by initializeRandomStates (oracle/refshim keeps curandState copies)
    uint rsW = 0, rsH = 0;                  // randomStates dims: max(photonLaunchW, W) x max(photonLaunchH, H)
    std::vector<RandomState> randomStates;
    std::vector<Hitpoint> hitpoints;        // raytracePassOutputBuffer
    std::vector<Photon> photons;            // NUM_PHOTONS = emitted * 4
    std::vector<Photon> photonsAsEmitted;   // copy taken before the grid build sorts `photons` in place (slot = 4 * thread + k): parity diagnostics
    std::vector<uint> photonKeys;           // m_photonsHashCells
    std::vector<uint> offsets;              // hashmapOffsetTable
    PhotonGrid grid{};
    std::vector<float3> indirect, direct, output;
    std::vector<uint> photonPathLength;     // debugPhotonPathLengthBuffer
    std::vector<uint> ptSignature;                 // parity diagnostics: per-pixel hash of the PATH_TRACING path's discrete decisions
    std::vector<unsigned char> directSampleMask;   // parity diagnostics: bit s = shadow sample s of the pixel was unoccluded and non-zero

    // per-iteration context scalars (OptixRenderer.cpp:555-557, 583-589, 616-617)
    Camera camera{};
    float iterationNumber = 0.f;
    uint localIterationNumber = 0;
    float ppmAlpha = 0.f, ppmRadius = 0.f, ppmRadiusSquared = 0.f;
    float emittedPhotonsPerIterationFloat = 0.f;

    // VCM (OptixRenderer.cpp:677-797)
    bool lightVertexCountEstimated = false;
    uint lightVertexCountEstimatePass = 1;
    uint maxPathLen = 10;
    float vmNormalizationFactor = 0.f, misVmWeightFactor = 0.f, misVcWeightFactor = 0.f;
    float2 pixelSizeFactor = {1.f, 1.f};
    uint lightSubpathCount = 0;
    std::vector<LightVertex>* lightVertexBuffer = nullptr;
    std::atomic<uint> lightVertexBufferIndex{0};
    std::vector<uint> lightSubpathVertexCount;
    std::vector<uint> lightSubpathVertexIndex;   // [path][vcmMaxPathLength-1]
    uint lightVertexBufferSize = 0;
    std::vector<uint> pathSignature;             // parity diagnostics: hash of the material kinds a BDPT subpath met, light paths [0, W*H), camera paths [W*H, 2 W*H)
    std::vector<uint> connectionCount;           // parity diagnostics: contributions that passed isOccluded, per target pixel, last iteration
    int vcmUseVC = 1, vcmUseVM = 0;
    int numThreads = 0;                     // 0 = OpenMP default

    RayStats stats;
    std::string lastError;

    Renderer();
    ~Renderer();
    void initialize();
    void initScene(const opr_scene_desc* d);
    void resizeBuffers(uint w, uint h);
    void initializeRandomStates();
    void renderNextIteration(unsigned long long iterationNumber, unsigned long long localIterationNumber,
                             float PPMRadius, bool createOutput, const opr_request_details& details);
    inline RandomState& rs(uint x, uint y) { return randomStates[(size_t)y * rsW + x]; }
};

// called once, serially, before every pass launch (no-op in the restatement; the _ref build refreshes its OptiX-style context)
void pass_begin(Renderer& r);

// PPM programs (ora_ppm.cpp)
void ppm_generateRay(Renderer& r, uint x, uint y);       // ppm/RayGeneratorPPM.cu:31-66
void ppm_photonGenerator(Renderer& r, uint x, uint y);   // ppm/PhotonGenerator.cu:81-128
void ppm_indirect(Renderer& r, uint x, uint y);          // ppm/IndirectRadianceEstimation.cu:69-227
void ppm_direct(Renderer& r, uint x, uint y);            // ppm/DirectRadianceEstimation.cu:29-77
void ppm_output(Renderer& r, uint x, uint y);            // ppm/Output.cu:32-37
void pt_generateRay(Renderer& r, uint x, uint y);        // pt/RayGeneratorPT.cu:46-131

// VCM programs (ora_vcm.cpp)
void vcm_lightPass(Renderer& r, uint x, uint y);         // vcm/VCMLightPass.cu:60-89
void vcm_cameraPass(Renderer& r, uint x, uint y);        // vcm/VCMCameraPass.cu:52-80
void vcm_alloc_vertices(Renderer& r, uint n);
void vcm_free(Renderer& r);

} // namespace ora
