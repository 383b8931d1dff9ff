// Exercises the C++ façade (include/OptixRenderer.h) the way the reference's StandaloneRenderManager drives OptixRenderer
// (Standalone/StandaloneRenderManager.cpp:60-140): initialize, initScene, the radius-schedule loop, getOutputBuffer.
//   facade_check rules            -> state rules / error strings that need no GPU
//   facade_check render W H N out -> N PPM iterations of the built-in Cornell scene, raw float3 sum buffer written to `out`
//   facade_check group W H N out G -> the same N global iterations handed out over G GPUs (OptixRendererGroup: the client loop of
//                                     Client/DistributedApplication.cpp:96-122 + one NCCL reduce), summed image written to `out`;
//                                     read back asynchronously after every hand-out round like the client's preview
#include "OptixRenderer.h"
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

static int expect_throw(const char* what, const char* needle, void (*fn)())
{
    try { fn(); } catch (const std::exception& e) {
        if (std::string(e.what()).find(needle) != std::string::npos) { printf("ok %s\n", what); return 0; }
        printf("FAIL %s: wrong message '%s'\n", what, e.what()); return 1;
    }
    printf("FAIL %s: no exception\n", what); return 1;
}

int main(int argc, char** argv)
{
    std::string mode = argc > 1 ? argv[1] : "rules";
    if (mode == "rules") {
        int bad = 0;
        bad += expect_throw("initScene before initialize", "Cannot initialize scene before OptixRenderer", [] {
            OptixRenderer r; Cornell scene; r.initScene(scene); });
        bad += expect_throw("render before initialize", "Traced before OptixRenderer was initialized", [] {
            OptixRenderer r; Cornell scene;
            RenderServerRenderRequestDetails d(scene.getDefaultCamera(), scene.getSceneName(), RenderMethod::PROGRESSIVE_PHOTON_MAPPING, 8, 8, 2.0 / 3.0);
            r.renderNextIteration(0, 0, 1.0f, true, d); });
        Cornell c; CornellSmall s(CornellSmall::LightPoint | CornellSmall::FloorGlossy);
        if (c.getSceneLights().size() != 1 || s.getSceneLights().size() != 1 || s.getSceneLights()[0].lightType != Light::POINT) { printf("FAIL lights\n"); bad++; }
        if (sizeof(Camera) != 96 || sizeof(Light) != 88) { printf("FAIL pod sizes\n"); bad++; }
        printf("r0 %.6f\n", c.getSceneInitialPPMRadiusEstimate());
        return bad;
    }
    const unsigned W = atoi(argv[2]), H = atoi(argv[3]); const int N = atoi(argv[4]);
    opr_config cfg = {64, 64, 0, 1645301512u, 1};
    if (OptixRenderer::PPM_INITIAL_RADIUS != 0.0f) return 4;
    if (mode == "group") {
        const int G = atoi(argv[6]);
        std::vector<int> devices;
        for (int i = 0; i < G; i++) devices.push_back(i);
        OptixRendererGroup group(devices, &cfg);
        if (group.size() != G) return 5;
        Cornell scene;
        group.initScene(scene);
        Camera cam = scene.getDefaultCamera();
        cam.setAspectRatio((float)W / (float)H);
        RenderServerRenderRequestDetails details(cam, scene.getSceneName(), RenderMethod::PROGRESSIVE_PHOTON_MAPPING, W, H, 2.0 / 3.0);
        std::vector<float> out((size_t)W * H * 3), preview((size_t)W * H * 3);
        for (int it = 0; it < N; it += G) {       // one hand-out round: the next G global iterations, one per GPU
            group.renderNextIterations(it, std::min(G, N - it), scene.getSceneInitialPPMRadiusEstimate(), details);
            group.getOutputBuffer(preview.data());   // a reduce per round must not disturb the running sums
        }
        group.getOutputBuffer(out.data());
        if (out != preview) return 6;
        FILE* f = fopen(argv[5], "wb"); fwrite(out.data(), sizeof(float), out.size(), f); fclose(f);
        return 0;
    }
    OptixRenderer renderer(cfg);
    renderer.initialize(ComputeDevice(0));
    Cornell scene;
    renderer.initScene(scene);
    Camera cam = scene.getDefaultCamera();
    cam.setAspectRatio((float)W / (float)H);
    RenderServerRenderRequestDetails details(cam, scene.getSceneName(), RenderMethod::PROGRESSIVE_PHOTON_MAPPING, W, H, 2.0 / 3.0);
    double PPMRadius = scene.getSceneInitialPPMRadiusEstimate();
    const double ppmAlpha = 2.0 / 3.0;
    for (int it = 0; it < N; it++) {                               // StandaloneRenderManager.cpp:104-107
        renderer.renderNextIteration(it, it, (float)PPMRadius, true, details);
        double r2 = PPMRadius * PPMRadius;
        r2 = r2 * (it + ppmAlpha) / (it + 1);
        PPMRadius = sqrt(r2);
    }
    std::vector<float> out((size_t)W * H * 3), async((size_t)W * H * 3);
    if (renderer.getScreenBufferSizeBytes() != out.size() * sizeof(float)) return 3;
    renderer.getOutputBufferAsync(async.data());
    renderer.waitOutput();
    renderer.getOutputBuffer(out.data());
    if (out != async) return 7;
    FILE* f = fopen(argv[5], "wb"); fwrite(out.data(), sizeof(float), out.size(), f); fclose(f);
    return 0;
}
