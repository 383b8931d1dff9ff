// StandaloneRenderManager.h — the driver loop on the caller's side of the hot path, mirroring
// Standalone/StandaloneRenderManager.cpp:60-140 without Qt: one renderNextIteration() per call, the PPM radius recurrence in
// host double (:105-107), a read-back of the running sum every 5th iteration (:91, "to make fair comparison with distributed
// renderer"), and RenderWidget's display normalisation / BMP export (Gui/gui/RenderWidget.cpp:187-199, 268-317).
#pragma once
#include "OptixRenderer.h"
#include <vector>

class StandaloneRenderManager {
public:
    StandaloneRenderManager(OptixRenderer& renderer, IScene& scene, RenderMethod::E method, unsigned width, unsigned height)
        : m_renderer(renderer), m_currentScene(&scene), m_renderMethod(method), m_width(width), m_height(height), m_camera(scene.getDefaultCamera()),
          m_nextIterationNumber(0), m_lastRendererIterationNumber(0), m_compileScene(true), m_outputEvery(5)
    {
        m_camera.setAspectRatio((float)width / (float)height);
        m_PPMRadius = scene.getSceneInitialPPMRadiusEstimate();     // StandaloneRenderManager::start, :60-72
    }
    void setCamera(const Camera& c) { m_camera = c; restart(); }
    void setOutputEvery(unsigned n) { m_outputEvery = n ? n : 1; }
    void restart() { m_nextIterationNumber = 0; m_PPMRadius = m_currentScene->getSceneInitialPPMRadiusEstimate(); }   // onCameraUpdated, :170

    // one pass of the render thread's loop (:75-140); returns true when the sum buffer was read back this iteration
    bool renderNextIteration()
    {
        if (m_compileScene) { m_renderer.initScene(*m_currentScene); m_compileScene = false; }
        const bool shouldOutputIteration = m_nextIterationNumber % m_outputEvery == 0;
        const double PPMAlpha = 2.0 / 3.0;
        RenderServerRenderRequestDetails details(m_camera, m_currentScene->getSceneName(), m_renderMethod, m_width, m_height, PPMAlpha);
        m_renderer.renderNextIteration(m_nextIterationNumber, m_nextIterationNumber, (float)m_PPMRadius, shouldOutputIteration, details);
        const double ppmRadiusSquared = m_PPMRadius * m_PPMRadius;
        const double ppmRadiusSquaredNew = ppmRadiusSquared * (m_nextIterationNumber + PPMAlpha) / double(m_nextIterationNumber + 1);
        m_PPMRadius = sqrt(ppmRadiusSquaredNew);
        m_lastRendererIterationNumber = m_nextIterationNumber;
        if (shouldOutputIteration) fetchOutput();
        m_nextIterationNumber++;
        return shouldOutputIteration;
    }
    void fetchOutput()
    {
        m_outputBuffer.resize((size_t)m_width * m_height * 3);
        m_renderer.getOutputBuffer(m_outputBuffer.data());
        m_previewedIterations = (float)(m_lastRendererIterationNumber + 1);     // RenderWidget.cpp:120
    }
    // RenderWidget::saveImageAsBMP with the application's default gamma 2.2 (Gui/Application.cpp:36-40)
    void saveImageAsBMP(const char* fileName, float gamma = 2.2f)
    {
        if (m_outputBuffer.empty()) fetchOutput();
        if (opr_save_bmp(fileName, m_outputBuffer.data(), m_width, m_height, m_previewedIterations, gamma)) throw std::runtime_error(opr_io_last_error());
    }
    const std::vector<float>& getOutputBuffer() const { return m_outputBuffer; }
    unsigned long long getNextIterationNumber() const { return m_nextIterationNumber; }
    double getPPMRadius() const { return m_PPMRadius; }

private:
    OptixRenderer& m_renderer;
    IScene* m_currentScene;
    RenderMethod::E m_renderMethod;
    unsigned m_width, m_height;
    Camera m_camera;
    unsigned long long m_nextIterationNumber, m_lastRendererIterationNumber;
    bool m_compileScene;
    unsigned m_outputEvery;
    double m_PPMRadius;
    float m_previewedIterations = 1.f;
    std::vector<float> m_outputBuffer;
};
