// DistributedRenderManager.h — the client side of the reference's distributed rendering, for the GPUs of one box: the request
// loop of Client/DistributedApplication.cpp:94-122 (the next `numIterations` global iteration numbers go out together with the
// PPM radius of each, carried by the host-double recurrence of :105-113), the render servers that answer them
// (Server/server/RenderServerRenderer.cpp:154-175) and the client-side sum of the returned images
// (Client/client/RenderResultPacketReceiver + Gui/gui/RenderWidget.cpp:120,197) — all behind OptixRendererGroup (opr_group_*):
// iteration i runs on GPU i % G with the radius of that recurrence (opr_ppm_radius_for_iteration is its closed form), and one
// NCCL reduce replaces the result packets.
#pragma once
#include "OptixRenderer.h"
#include <vector>

class DistributedRenderManager {
public:
    DistributedRenderManager(const std::vector<int>& cudaOrdinals, IScene& scene, RenderMethod::E method, unsigned width, unsigned height,
                             const opr_config* cfg = nullptr)
        : m_group(cudaOrdinals, cfg), m_currentScene(&scene), m_renderMethod(method), m_width(width), m_height(height), m_camera(scene.getDefaultCamera()),
          m_nextRenderServerRenderRequestIteration(0), m_compileScene(true)
    {
        m_camera.setAspectRatio((float)width / (float)height);
    }
    int numServers() const { return m_group.size(); }
    // one round of requests: every GPU gets `iterationsPerServer` iterations (DistributedApplication::getNextRenderServerRenderRequest)
    void renderNextRound(unsigned iterationsPerServer = 1)
    {
        if (m_compileScene) { m_group.initScene(*m_currentScene); m_compileScene = false; }
        const double PPMAlpha = 2.0 / 3.0;
        RenderServerRenderRequestDetails details(m_camera, m_currentScene->getSceneName(), m_renderMethod, m_width, m_height, PPMAlpha);
        const unsigned long long count = (unsigned long long)iterationsPerServer * (unsigned long long)m_group.size();
        m_group.renderNextIterations(m_nextRenderServerRenderRequestIteration, count, m_currentScene->getSceneInitialPPMRadiusEstimate(), details);
        m_nextRenderServerRenderRequestIteration += count;
    }
    // sum of all servers' running sums (one NCCL reduce); previewedIterations = iterations rendered by all of them
    void fetchOutput()
    {
        m_outputBuffer.resize((size_t)m_width * m_height * 3);
        m_group.getOutputBuffer(m_outputBuffer.data());
        m_previewedIterations = (float)m_nextRenderServerRenderRequestIteration;
    }
    void saveImageAsBMP(const char* fileName, float gamma = 2.2f)
    {
        if (m_outputBuffer.empty()) fetchOutput();
        if (opr_save_bmp(fileName, m_outputBuffer.data(), m_width, m_height, m_previewedIterations, gamma)) throw std::runtime_error(opr_io_last_error());
    }
    const std::vector<float>& getOutputBuffer() const { return m_outputBuffer; }
    unsigned long long getNextIterationNumber() const { return m_nextRenderServerRenderRequestIteration; }
    double getPPMRadius() const
    { return opr_ppm_radius_for_iteration(m_currentScene->getSceneInitialPPMRadiusEstimate(), m_nextRenderServerRenderRequestIteration, 2.0 / 3.0); }

private:
    OptixRendererGroup m_group;
    IScene* m_currentScene;
    RenderMethod::E m_renderMethod;
    unsigned m_width, m_height;
    Camera m_camera;
    unsigned long long m_nextRenderServerRenderRequestIteration;
    bool m_compileScene;
    float m_previewedIterations = 1.f;
    std::vector<float> m_outputBuffer;
};
