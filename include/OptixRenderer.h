// OptixRenderer.h — C++ host surface of the B200 light-transport path, mirroring the reference's class,
// method and argument names so a caller of RenderEngine.dll can switch by changing an include:
//
//   reference                                                   here
//   --------------------------------------------------------    ----------------------------------------
//   class OptixRenderer   (renderer/OptixRenderer.h:21-44)      class OptixRenderer (same methods)
//   class ComputeDevice   (ComputeDevice.h:12-76)               class ComputeDevice (getDeviceId only, :416)
//   class IScene          (scene/IScene.h:16-29)                class IScene: getSceneRootGroup(optix::Context&)
//                                                                 is replaced by getSceneDescription()
//   class Cornell / CornellSmall (scene/*.cpp)                  same names, same constants
//   class Camera / Light / Vector3 / AAB / Sphere               same names, same fields (PODs are opr_camera/opr_light)
//   RenderMethod::E       (renderer/RenderMethod.h:13-19)       same enum
//   RenderServerRenderRequestDetails (clientserver/...h:15-33)  same getters; QByteArray -> std::string
//
// Header-only over the C ABI in opr.h; errors become std::runtime_error (the reference throws
// std::exception with the same messages).
#pragma once
#include "opr.h"
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace RenderMethod { enum E { PATH_TRACING, VCM_BIDIRECTIONAL_PATH_TRACING, PROGRESSIVE_PHOTON_MAPPING, NONE }; }

class Vector3 {
public:
    float x, y, z;
    Vector3() : x(0), y(0), z(0) {}
    Vector3(float a) : x(a), y(a), z(a) {}
    Vector3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
    explicit Vector3(const float* p) : x(p[0]), y(p[1]), z(p[2]) {}
    Vector3 operator+(const Vector3& o) const { return Vector3(x + o.x, y + o.y, z + o.z); }
    Vector3 operator-(const Vector3& o) const { return Vector3(x - o.x, y - o.y, z - o.z); }
    Vector3 operator*(float s) const { return Vector3(x * s, y * s, z * s); }
    // math/Vector3.cpp:30-33 multiplies a.z*b.x; kept because the scene bounding sphere depends on it
    static float dot(const Vector3& a, const Vector3& b) { return a.x * b.x + a.y * b.y + a.z * b.x; }
    float length2() const { return dot(*this, *this); }
    float length() const { return sqrtf(length2()); }
    float max() const { float m = x > y ? x : y; return m > z ? m : z; }
    void store(float* p) const { p[0] = x; p[1] = y; p[2] = z; }
};
inline Vector3 operator+(const Vector3& a, float b) { return Vector3(a.x + b, a.y + b, a.z + b); }

class Sphere { public: Vector3 center; float radius; Sphere() : radius(0) {} Sphere(const Vector3& c, float r) : center(c), radius(r) {} };

class AAB {
public:
    Vector3 min, max;
    AAB() {}
    AAB(const Vector3& mn, const Vector3& mx) : min(mn), max(mx) {}
    Vector3 getExtent() const { return max - min; }
    Vector3 getCenter() const { return (min + max) * 0.5f; }
    Sphere getBoundingSphere() const { Vector3 c = getCenter(); return Sphere(c, (max - c).length()); }   // math/AAB.cpp:31-38
};

// renderer/Camera.h:13-80 — byte layout of opr_camera
class Camera : public opr_camera {
public:
    enum AspectRatioMode { KeepVertical, KeepHorizontal, KeepNone };
    Camera() { init(Vector3(1, 1, 1), Vector3(0, 0, 0), Vector3(0, 1, 0), 60.f, 60.f, 0.f, KeepVertical); }
    Camera(Vector3 eye_, Vector3 lookat_) { init(eye_, lookat_, Vector3(0, 1, 0), 60.f, 60.f, 0.f, KeepVertical); }
    Camera(Vector3 eye_, Vector3 lookat_, Vector3 up_, float hfov_ = 60, float vfov_ = 60, float aperture_ = 0, AspectRatioMode arm = KeepVertical)
    { init(eye_, lookat_, up_, hfov_, vfov_, aperture_, arm); }
    explicit Camera(const opr_camera& c) { *static_cast<opr_camera*>(this) = c; }
    void setup() { opr_camera_setup(this); }
    void setAspectRatio(float ratio) { opr_camera_set_aspect_ratio(this, ratio); }
private:
    void init(Vector3 e, Vector3 l, Vector3 u, float hf, float vf, float ap, AspectRatioMode arm)
    {
        memset(static_cast<opr_camera*>(this), 0, sizeof(opr_camera));
        e.store(eye); l.store(lookat); u.store(up); hfov = hf; vfov = vf; aperture = ap; aspectRatioMode = arm;
        setup();
    }
};

// renderer/Light.h:14-54 — byte layout of opr_light
class Light : public opr_light {
public:
    enum LightType { AREA, POINT, SPOT };
    Light() { memset(static_cast<opr_light*>(this), 0, sizeof(opr_light)); }
    Light(Vector3 power_, Vector3 position_, Vector3 v1_, Vector3 v2_)
    { float p[3], o[3], a[3], b[3]; power_.store(p); position_.store(o); v1_.store(a); v2_.store(b); opr_make_area_light(this, p, o, a, b); }
    Light(Vector3 power_, Vector3 position_) { float p[3], o[3]; power_.store(p); position_.store(o); opr_make_point_light(this, p, o); }
};

class ComputeDevice {   // ComputeDevice.h:12-76; the renderer reads getDeviceId() only (OptixRenderer.cpp:416)
    int m_deviceId;
public:
    explicit ComputeDevice(int cudaOrdinal = 0) : m_deviceId(cudaOrdinal) {}
    int getDeviceId() const { return m_deviceId; }
};

// scene/IScene.h:16-29
class IScene {
public:
    virtual ~IScene() {}
    virtual const opr_scene_desc& getSceneDescription() = 0;     // replaces getSceneRootGroup(optix::Context&)
    virtual const std::vector<Light>& getSceneLights() const = 0;
    virtual Camera getDefaultCamera() const = 0;
    virtual const char* getSceneName() const = 0;
    virtual AAB getSceneAABB() const = 0;
    virtual float getSceneInitialPPMRadiusEstimate() const
    {
        AAB a = getSceneAABB(); float mn[3], mx[3]; a.min.store(mn); a.max.store(mx);
        return opr_initial_ppm_radius(mn, mx);
    }
    virtual unsigned int getNumTriangles() const = 0;
    virtual unsigned int getNumMeshes() const = 0;
};

// Built-in scenes (scene/Cornell.cpp, scene/CornellSmall.cpp); the tables live in the library.
class BuiltinSceneBase : public IScene {
protected:
    void* m_handle;
    std::vector<Light> m_sceneLights;
    BuiltinSceneBase(const char* name, unsigned config) : m_handle(opr_builtin_scene_create(name, config))
    {
        if (!m_handle) throw std::runtime_error("unknown built-in scene");
        const opr_scene_desc* d = opr_builtin_scene_desc(m_handle);
        for (int i = 0; i < d->numLights; i++) { Light l; *static_cast<opr_light*>(&l) = d->lights[i]; m_sceneLights.push_back(l); }
    }
public:
    ~BuiltinSceneBase() { opr_builtin_scene_destroy(m_handle); }
    const opr_scene_desc& getSceneDescription() { return *opr_builtin_scene_desc(m_handle); }
    const std::vector<Light>& getSceneLights() const { return m_sceneLights; }
    Camera getDefaultCamera() const { opr_camera c; opr_builtin_scene_camera(m_handle, &c); return Camera(c); }
    AAB getSceneAABB() const { const opr_scene_desc* d = opr_builtin_scene_desc(m_handle); return AAB(Vector3(d->aabbMin), Vector3(d->aabbMax)); }
    unsigned int getNumTriangles() const { return 16; }   // Cornell.cpp:224-232 returns the constant 16
    unsigned int getNumMeshes() const { return 16; }
};

class Cornell : public BuiltinSceneBase {
public:
    Cornell() : BuiltinSceneBase("Cornell", 0) {}
    const char* getSceneName() const { return getCornellSceneName(); }
    static const char* getCornellSceneName() { return "Cornell"; }
};

class CornellSmall : public BuiltinSceneBase {
public:
    enum Config {   // scene/CornellSmall.h:25-42
        SmallVCMColors = 1 << 0, LightArea = 1 << 1, LightAreaUpwards = 1 << 2, LightPoint = 1 << 3, LightPointStrong = 1 << 4,
        LightPointDistant = 1 << 5, BackwallBlue = 1 << 6, FloorMirror = 1 << 7, FloorGlossy = 1 << 8, Blocks = 1 << 9,
        LargeMirrorSphere = 1 << 10, LargeGlassSphere = 1 << 11, SmallMirrorSphere = 1 << 12, SmallGlassSphere = 1 << 13,
        Default = (LightArea | Blocks)
    };
    CornellSmall() : BuiltinSceneBase("CornellSmall", Default) {}
    explicit CornellSmall(unsigned config) : BuiltinSceneBase("CornellSmall", config) {}
    const char* getSceneName() const { return getCornellSceneName(); }
    static const char* getCornellSceneName() { return "CornellSmall"; }
};

// scene/Scene.h — a scene read from a Collada file (Scene::createFromFile, scene/Scene.cpp:73-176)
class Scene : public IScene {
    void* m_handle;
    std::vector<Light> m_sceneLights;
    explicit Scene(void* h) : m_handle(h)
    {
        const opr_scene_desc* d = opr_scene_file_desc(m_handle);
        for (int i = 0; i < d->numLights; i++) { Light l; *static_cast<opr_light*>(&l) = d->lights[i]; m_sceneLights.push_back(l); }
    }
public:
    static IScene* createFromFile(const char* filename)
    {
        void* h = opr_scene_file_load(filename, 0);
        if (!h) throw std::runtime_error(opr_scene_file_last_error());
        return new Scene(h);
    }
    ~Scene() { opr_scene_file_destroy(m_handle); }
    const opr_scene_desc& getSceneDescription() { return *opr_scene_file_desc(m_handle); }
    const std::vector<Light>& getSceneLights() const { return m_sceneLights; }
    Camera getDefaultCamera() const { opr_camera c; if (opr_scene_file_camera(m_handle, &c)) return Camera(); return Camera(c); }
    const char* getSceneName() const { return opr_scene_file_desc(m_handle)->name; }
    AAB getSceneAABB() const { const opr_scene_desc* d = opr_scene_file_desc(m_handle); return AAB(Vector3(d->aabbMin), Vector3(d->aabbMax)); }
    unsigned int getNumTriangles() const { return opr_scene_file_num_triangles(m_handle); }
    unsigned int getNumMeshes() const { return opr_scene_file_num_meshes(m_handle); }
};

// clientserver/RenderServerRenderRequestDetails.h:15-33
class RenderServerRenderRequestDetails {
    Camera m_camera; std::string m_sceneName; RenderMethod::E m_renderMethod; unsigned m_width, m_height; double m_ppmAlpha;
public:
    RenderServerRenderRequestDetails() : m_renderMethod(RenderMethod::NONE), m_width(0), m_height(0), m_ppmAlpha(0) {}
    RenderServerRenderRequestDetails(const Camera& camera, const std::string& sceneName, RenderMethod::E renderMethod,
                                     unsigned int width, unsigned int height, double ppmAlpha)
        : m_camera(camera), m_sceneName(sceneName), m_renderMethod(renderMethod), m_width(width), m_height(height), m_ppmAlpha(ppmAlpha) {}
    unsigned int getWidth() const { return m_width; }
    unsigned int getHeight() const { return m_height; }
    double getPPMAlpha() const { return m_ppmAlpha; }
    const Camera& getCamera() const { return m_camera; }
    const std::string& getSceneName() const { return m_sceneName; }
    const RenderMethod::E getRenderMethod() const { return m_renderMethod; }
};

// renderer/OptixRenderer.h:21-44
class OptixRenderer {
    opr_renderer* m_r;
    static void check(int rc) { if (rc != 0) throw std::runtime_error(opr_last_error()); }
public:
    // OptixRenderer.cpp:38-61
    static const unsigned int PHOTON_GRID_MAX_SIZE = 100 * 100 * 100;
    static const unsigned int MAX_PHOTON_COUNT = 4;
    static const unsigned int PHOTON_LAUNCH_WIDTH = 1024;
    static const unsigned int PHOTON_LAUNCH_HEIGHT = 1024;
    static const unsigned int EMITTED_PHOTONS_PER_ITERATION = PHOTON_LAUNCH_WIDTH * PHOTON_LAUNCH_HEIGHT;
    static const unsigned int NUM_PHOTONS = EMITTED_PHOTONS_PER_ITERATION * MAX_PHOTON_COUNT;
    static const unsigned int VCM_MAX_PATH_LENGTH = 10u;
    // OptixRenderer.h:41 declares this member and no translation unit of the reference defines or reads it (the initial radius
    // comes from IScene::getSceneInitialPPMRadiusEstimate, scene/IScene.cpp:23-31).  Declared for source compatibility; defined
    // below as 0 so that code naming it links.
    const static float PPM_INITIAL_RADIUS;

    OptixRenderer() : m_r(nullptr) { check(opr_create(&m_r, nullptr)); }
    explicit OptixRenderer(const opr_config& cfg) : m_r(nullptr) { check(opr_create(&m_r, &cfg)); }
    ~OptixRenderer() { opr_destroy(m_r); }
    OptixRenderer(const OptixRenderer&) = delete;
    OptixRenderer& operator=(const OptixRenderer&) = delete;

    void initialize(const ComputeDevice& device) { check(opr_initialize(m_r, device.getDeviceId())); }
    void initScene(IScene& scene) { check(opr_init_scene(m_r, &scene.getSceneDescription())); }
    void renderNextIteration(unsigned long long iterationNumber, unsigned long long localIterationNumber, float PPMRadius,
                             bool createOutput, const RenderServerRenderRequestDetails& details)
    {
        opr_request_details d; memset(&d, 0, sizeof d);
        d.camera = details.getCamera(); d.renderMethod = (int)details.getRenderMethod();
        d.width = details.getWidth(); d.height = details.getHeight(); d.ppmAlpha = (float)details.getPPMAlpha();
        d.sceneName = details.getSceneName().c_str();
        check(opr_render_next_iteration(m_r, iterationNumber, localIterationNumber, PPMRadius, createOutput ? 1 : 0, &d));
    }
    void getOutputBuffer(void* data) { check(opr_get_output_buffer(m_r, data)); }
    unsigned int getWidth() const { return opr_get_width(m_r); }
    unsigned int getHeight() const { return opr_get_height(m_r); }
    unsigned int getScreenBufferSizeBytes() const { return opr_get_screen_buffer_size_bytes(m_r); }
    // additions: read-back that overlaps the next iterations (data must stay valid until waitOutput returns)
    void getOutputBufferAsync(void* data) { check(opr_get_output_buffer_async(m_r, data)); }
    void waitOutput() { check(opr_wait_output(m_r)); }
    opr_renderer* handle() const { return m_r; }
};
inline const float OptixRenderer::PPM_INITIAL_RADIUS = 0.0f;

// The GPUs of one box behind one object: replaces the client's hand-out loop (Client/DistributedApplication.cpp:96-122) and the
// render servers' result packets (Server/server/RenderServerRenderer.cpp:154-175).  Global iteration i is rendered by GPU
// i % size() as its local iteration i / size(), with the radius of the driver's recurrence; getOutputBuffer sums the GPUs'
// running sums with one NCCL reduce.
class OptixRendererGroup {
    opr_group* m_g;
    static void check(int rc) { if (rc != 0) throw std::runtime_error(opr_last_error()); }
public:
    explicit OptixRendererGroup(const std::vector<int>& cudaOrdinals, const opr_config* cfg = nullptr) : m_g(nullptr)
    { check(opr_group_create(&m_g, (int)cudaOrdinals.size(), cudaOrdinals.data(), cfg)); }
    ~OptixRendererGroup() { opr_group_destroy(m_g); }
    OptixRendererGroup(const OptixRendererGroup&) = delete;
    OptixRendererGroup& operator=(const OptixRendererGroup&) = delete;
    int size() const { return opr_group_size(m_g); }
    void initScene(IScene& scene) { check(opr_group_init_scene(m_g, &scene.getSceneDescription())); }
    void renderNextIterations(unsigned long long firstIteration, unsigned long long count, double initialPPMRadius,
                              const RenderServerRenderRequestDetails& details)
    {
        opr_request_details d; memset(&d, 0, sizeof d);
        d.camera = details.getCamera(); d.renderMethod = (int)details.getRenderMethod();
        d.width = details.getWidth(); d.height = details.getHeight(); d.ppmAlpha = (float)details.getPPMAlpha();
        d.sceneName = details.getSceneName().c_str();
        check(opr_group_render_next_iterations(m_g, firstIteration, count, initialPPMRadius, details.getPPMAlpha(), &d));
    }
    void getOutputBuffer(void* data) { check(opr_group_reduce_output(m_g, data)); }
};
