// opr_render — command-line stand-in for the reference's Standalone application: renders a built-in scene with the driver
// loop of Standalone/StandaloneRenderManager.cpp and writes the display image as RenderWidget::saveImageAsBMP does.
//   opr_render [--scene Cornell|CornellSmall] [--config BITS] [--method ppm|bdpt|pt] [--size WxH] [--iterations N] [--out file.bmp] [--raw file.f32]
//              [--gpus G]   G > 0: the Client/Server arrangement instead (DistributedRenderManager: iterations handed out over G GPUs, one NCCL reduce)
#include "StandaloneRenderManager.h"
#include "DistributedRenderManager.h"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>

int main(int argc, char** argv)
{
    std::string sceneName = "Cornell", method = "ppm", out = "opr_render.bmp", raw;
    unsigned config = 0, width = 512, height = 512, iterations = 20, gpus = 0;   // Gui/Application.cpp:36-40 default 512x512
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() -> const char* { if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", a.c_str()); exit(2); } return argv[++i]; };
        if (a == "--scene") sceneName = next();
        else if (a == "--config") config = (unsigned)strtoul(next(), nullptr, 0);
        else if (a == "--method") method = next();
        else if (a == "--size") { if (sscanf(next(), "%ux%u", &width, &height) != 2) { fprintf(stderr, "--size WxH\n"); return 2; } }
        else if (a == "--iterations") iterations = (unsigned)atoi(next());
        else if (a == "--gpus") gpus = (unsigned)atoi(next());
        else if (a == "--out") out = next();
        else if (a == "--raw") raw = next();
        else { fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
    }
    try {
        std::unique_ptr<IScene> scene;
        if (sceneName == "Cornell") scene.reset(new Cornell());
        else if (sceneName == "CornellSmall") scene.reset(config ? new CornellSmall(config) : new CornellSmall());
        else { fprintf(stderr, "unknown scene %s\n", sceneName.c_str()); return 2; }
        const RenderMethod::E m = method == "bdpt" ? RenderMethod::VCM_BIDIRECTIONAL_PATH_TRACING : (method == "pt" ? RenderMethod::PATH_TRACING : RenderMethod::PROGRESSIVE_PHOTON_MAPPING);
        if (gpus > 0) {
            std::vector<int> devices;
            for (unsigned g = 0; g < gpus; g++) devices.push_back((int)g);
            DistributedRenderManager manager(devices, *scene, m, width, height);
            const auto t0 = std::chrono::steady_clock::now();
            while (manager.getNextIterationNumber() < iterations) manager.renderNextRound(1);
            manager.fetchOutput();
            const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            manager.saveImageAsBMP(out.c_str());
            if (!raw.empty()) { FILE* f = fopen(raw.c_str(), "wb"); fwrite(manager.getOutputBuffer().data(), sizeof(float), manager.getOutputBuffer().size(), f); fclose(f); }
            printf("%s %s %ux%u on %u GPU(s): %llu iterations in %.3f s (%.1f it/s, includes scene upload + BVH build), next radius %.6g -> %s\n", sceneName.c_str(),
                   method.c_str(), width, height, gpus, manager.getNextIterationNumber(), s, manager.getNextIterationNumber() / s, manager.getPPMRadius(), out.c_str());
            return 0;
        }
        OptixRenderer renderer;
        renderer.initialize(ComputeDevice(0));
        StandaloneRenderManager manager(renderer, *scene, m, width, height);
        const auto t0 = std::chrono::steady_clock::now();
        for (unsigned i = 0; i < iterations; i++) manager.renderNextIteration();
        manager.fetchOutput();
        const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        manager.saveImageAsBMP(out.c_str());
        if (!raw.empty()) { FILE* f = fopen(raw.c_str(), "wb"); fwrite(manager.getOutputBuffer().data(), sizeof(float), manager.getOutputBuffer().size(), f); fclose(f); }
        printf("%s %s %ux%u: %u iterations in %.3f s (%.1f it/s, first iteration includes scene upload + BVH build), final radius %.6g -> %s\n", sceneName.c_str(),
               method.c_str(), width, height, iterations, s, iterations / s, manager.getPPMRadius(), out.c_str());
    } catch (const std::exception& e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
