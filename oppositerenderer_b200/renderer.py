"""Python mirror of the reference's renderer-facing classes over the C ABI (include/opr.h).

Names, argument order and error behaviour follow RenderEngine/renderer/OptixRenderer.h:21-44,
RenderEngine/scene/IScene.h:16-29, RenderEngine/clientserver/RenderServerRenderRequestDetails.h:15-33.
"""
import ctypes as C

import numpy as np

from . import capi


class RenderMethod:  # RenderEngine/renderer/RenderMethod.h:13-19
    PATH_TRACING = capi.PATH_TRACING
    VCM_BIDIRECTIONAL_PATH_TRACING = capi.VCM_BIDIRECTIONAL_PATH_TRACING
    PROGRESSIVE_PHOTON_MAPPING = capi.PROGRESSIVE_PHOTON_MAPPING
    NONE = capi.RENDER_METHOD_NONE


class ComputeDevice:  # RenderEngine/ComputeDevice.h:12-76 (only getDeviceId is used by the renderer)
    def __init__(self, device_id=0):
        self._id = int(device_id)

    def getDeviceId(self):
        return self._id


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


class Camera:
    """RenderEngine/renderer/Camera.h:13-80.  ``pod`` is the 96-byte struct uploaded to the device."""
    KeepVertical, KeepHorizontal, KeepNone = 0, 1, 2

    def __init__(self, eye=(1, 1, 1), lookat=(0, 0, 0), up=(0, 1, 0), hfov=60.0, vfov=60.0, aperture=0.0,
                 aspectRatioMode=0, pod=None):
        lib = capi.load_library()
        if pod is not None:
            self.pod = pod
            return
        self.pod = capi.opr_camera()
        self.pod.eye, self.pod.lookat, self.pod.up = _f3(eye), _f3(lookat), _f3(up)
        self.pod.hfov, self.pod.vfov, self.pod.aperture = hfov, vfov, aperture
        self.pod.aspectRatioMode = aspectRatioMode
        lib.opr_camera_setup(C.byref(self.pod))

    def setAspectRatio(self, ratio):
        capi.load_library().opr_camera_set_aspect_ratio(C.byref(self.pod), float(ratio))


class Light:
    """RenderEngine/renderer/Light.h:14-54 (88-byte POD) with the area / point constructors of Light.cpp:14-40."""

    def __init__(self, power, position, v1=None, v2=None):
        lib = capi.load_library()
        self.pod = capi.opr_light()
        if v1 is not None:
            lib.opr_make_area_light(C.byref(self.pod), _f3(power), _f3(position), _f3(v1), _f3(v2))
        else:
            lib.opr_make_point_light(C.byref(self.pod), _f3(power), _f3(position))


class RenderServerRenderRequestDetails:
    def __init__(self, camera, sceneName, renderMethod, width, height, ppmAlpha):
        self._camera, self._sceneName, self._renderMethod = camera, sceneName, renderMethod
        self._width, self._height, self._ppmAlpha = int(width), int(height), float(ppmAlpha)

    def getWidth(self):
        return self._width

    def getHeight(self):
        return self._height

    def getPPMAlpha(self):
        return self._ppmAlpha

    def getCamera(self):
        return self._camera

    def getSceneName(self):
        return self._sceneName

    def getRenderMethod(self):
        return self._renderMethod

    def pod(self):
        name = self._sceneName.encode() if isinstance(self._sceneName, str) else self._sceneName
        return capi.opr_request_details(self._camera.pod, self._renderMethod, self._width, self._height, self._ppmAlpha, name)


class _BuiltinScene:
    """IScene over the library's built-in scene tables (scene/Cornell.cpp, scene/CornellSmall.cpp)."""

    def __init__(self, name, config=0):
        self._lib = capi.load_library()
        self._h = self._lib.opr_builtin_scene_create(name.encode(), config)
        if not self._h:
            raise ValueError("unknown built-in scene %r" % name)
        self._name = name

    def __del__(self):
        try:
            self._lib.opr_builtin_scene_destroy(self._h)
        except Exception:
            pass

    def getSceneDescription(self):
        return self._lib.opr_builtin_scene_desc(self._h)

    def getSceneLights(self):
        d = self.getSceneDescription().contents
        return [d.lights[i] for i in range(d.numLights)]

    def getDefaultCamera(self):
        pod = capi.opr_camera()
        self._lib.opr_builtin_scene_camera(self._h, C.byref(pod))
        return Camera(pod=pod)

    def getSceneName(self):
        return self._name

    def getSceneAABB(self):
        d = self.getSceneDescription().contents
        return np.array(list(d.aabbMin), np.float32), np.array(list(d.aabbMax), np.float32)

    def getSceneInitialPPMRadiusEstimate(self):
        d = self.getSceneDescription().contents
        return self._lib.opr_initial_ppm_radius(d.aabbMin, d.aabbMax)

    def getNumTriangles(self):
        return 16

    def getNumMeshes(self):
        return 16


class Cornell(_BuiltinScene):
    def __init__(self):
        super().__init__("Cornell", 0)


class CornellSmall(_BuiltinScene):
    # scene/CornellSmall.h:25-42
    SmallVCMColors, LightArea, LightAreaUpwards, LightPoint, LightPointStrong = 1 << 0, 1 << 1, 1 << 2, 1 << 3, 1 << 4
    LightPointDistant, BackwallBlue, FloorMirror, FloorGlossy, Blocks = 1 << 5, 1 << 6, 1 << 7, 1 << 8, 1 << 9
    LargeMirrorSphere, LargeGlassSphere, SmallMirrorSphere, SmallGlassSphere = 1 << 10, 1 << 11, 1 << 12, 1 << 13
    Default = (1 << 1) | (1 << 9)

    def __init__(self, config=None):
        super().__init__("CornellSmall", self.Default if config is None else config)


class Scene:
    """RenderEngine/scene/Scene.h — a scene read from a Collada file (Scene::createFromFile, scene/Scene.cpp:73-176)."""

    def __init__(self, handle, lib):
        self._h, self._lib = handle, lib

    @staticmethod
    def createFromFile(filename, single_emitter_lights=False):
        lib = capi.load_library()
        h = lib.opr_scene_file_load(str(filename).encode(), 1 if single_emitter_lights else 0)
        if not h:
            raise RuntimeError(lib.opr_scene_file_last_error().decode(errors="replace"))
        return Scene(h, lib)

    def __del__(self):
        try:
            self._lib.opr_scene_file_destroy(self._h)
        except Exception:
            pass

    def getSceneDescription(self):
        return self._lib.opr_scene_file_desc(self._h)

    def getSceneLights(self):
        d = self.getSceneDescription().contents
        return [d.lights[i] for i in range(d.numLights)]

    def getDefaultCamera(self):
        pod = capi.opr_camera()
        if self._lib.opr_scene_file_camera(self._h, C.byref(pod)) != 0:
            return Camera()                     # Scene::m_defaultCamera default-constructed (Scene.h)
        return Camera(pod=pod)

    def getSceneName(self):
        return self.getSceneDescription().contents.name.decode(errors="replace")

    def getSceneAABB(self):
        d = self.getSceneDescription().contents
        return np.array(list(d.aabbMin), np.float32), np.array(list(d.aabbMax), np.float32)

    def getSceneInitialPPMRadiusEstimate(self):
        d = self.getSceneDescription().contents
        return self._lib.opr_initial_ppm_radius(d.aabbMin, d.aabbMax)

    def getNumTriangles(self):
        return self._lib.opr_scene_file_num_triangles(self._h)

    def getNumMeshes(self):
        return self._lib.opr_scene_file_num_meshes(self._h)

    def warnings(self):
        return self._lib.opr_scene_file_warnings(self._h).decode(errors="replace")


class SceneDescription:
    """IScene built from numpy arrays (triangle meshes, parallelograms, spheres, materials, textures, lights):
    the neutral replacement of IScene::getSceneRootGroup.  Keeps the arrays alive."""

    def __init__(self, name, aabb_min, aabb_max, materials, lights, meshes=(), parallelograms=(), spheres=(), textures=(),
                 default_camera=None):
        self._keep = []
        self._name = name
        self._camera = default_camera
        d = capi.opr_scene_desc()
        self._name_b = name.encode()
        d.name = self._name_b
        ms = (capi.opr_mesh * max(len(meshes), 1))()
        for i, m in enumerate(meshes):
            v = np.ascontiguousarray(m["vertices"], np.float32)
            idx = np.ascontiguousarray(m["indices"], np.int32)
            self._keep += [v, idx]
            ms[i].numVertices, ms[i].numTriangles = v.shape[0], idx.shape[0]
            ms[i].vertices = v.ctypes.data_as(C.POINTER(C.c_float))
            ms[i].indices = idx.ctypes.data_as(C.POINTER(C.c_int32))
            for key in ("normals", "uvs", "tangents", "bitangents"):
                if m.get(key) is not None:
                    a = np.ascontiguousarray(m[key], np.float32)
                    self._keep.append(a)
                    setattr(ms[i], key, a.ctypes.data_as(C.POINTER(C.c_float)))
            ms[i].material = int(m["material"])
        d.numMeshes, d.meshes = len(meshes), ms
        pg = (capi.opr_parallelogram * max(len(parallelograms), 1))()
        for i, p in enumerate(parallelograms):
            pg[i].anchor, pg[i].offset1, pg[i].offset2, pg[i].material = _f3(p[0]), _f3(p[1]), _f3(p[2]), int(p[3])
        d.numParallelograms, d.parallelograms = len(parallelograms), pg
        sp = (capi.opr_sphere * max(len(spheres), 1))()
        for i, s in enumerate(spheres):
            sp[i].center, sp[i].radius, sp[i].material = _f3(s[0]), float(s[1]), int(s[2])
        d.numSpheres, d.spheres = len(spheres), sp
        mt = (capi.opr_material * max(len(materials), 1))()
        for i, m in enumerate(materials):
            mt[i] = m
        d.numMaterials, d.materials = len(materials), mt
        tx = (capi.opr_texture * max(len(textures), 1))()
        for i, t in enumerate(textures):
            a = np.ascontiguousarray(t, np.uint8)
            assert a.ndim == 3 and a.shape[2] == 4
            self._keep.append(a)
            tx[i].width, tx[i].height = a.shape[1], a.shape[0]
            tx[i].rgba = a.ctypes.data_as(C.POINTER(C.c_uint8))
        d.numTextures, d.textures = len(textures), tx
        lt = (capi.opr_light * max(len(lights), 1))()
        for i, l in enumerate(lights):
            lt[i] = l.pod if hasattr(l, "pod") else l
        d.numLights, d.lights = len(lights), lt
        d.aabbMin, d.aabbMax = _f3(aabb_min), _f3(aabb_max)
        self._keep += [ms, pg, sp, mt, tx, lt]
        self._desc = d
        self.num_triangles = sum(int(np.asarray(m["indices"]).shape[0]) for m in meshes)
        self.num_meshes = len(meshes)

    def getSceneDescription(self):
        return C.pointer(self._desc)

    def getDefaultCamera(self):
        return self._camera

    def getSceneName(self):
        return self._name

    def getSceneAABB(self):
        return np.array(list(self._desc.aabbMin), np.float32), np.array(list(self._desc.aabbMax), np.float32)

    def getSceneInitialPPMRadiusEstimate(self):
        return capi.load_library().opr_initial_ppm_radius(self._desc.aabbMin, self._desc.aabbMax)

    def getNumTriangles(self):
        return self.num_triangles

    def getNumMeshes(self):
        return self.num_meshes


def make_material(kind, Kd=(0, 0, 0), Ks=(0, 0, 0), exponent=0.0, Kr=(0, 0, 0), Kt=(0, 0, 0), ior=0.0, power=(0, 0, 0),
                  inverseArea=0.0, diffuseTexture=-1, normalTexture=-1):
    m = capi.opr_material()
    m.kind = kind
    m.Kd, m.Ks, m.Kr, m.Kt, m.power = _f3(Kd), _f3(Ks), _f3(Kr), _f3(Kt), _f3(power)
    m.exponent, m.indexOfRefraction, m.inverseArea = exponent, ior, inverseArea
    m.diffuseTexture, m.normalTexture = diffuseTexture, normalTexture
    return m


def ppm_radius_schedule(r0, n, alpha=2.0 / 3.0):
    """Driver-side radius schedule, Standalone/StandaloneRenderManager.cpp:105-107 (host double)."""
    lib = capi.load_library()
    return [lib.opr_ppm_radius_for_iteration(float(r0), i, alpha) for i in range(n)]


class OptixRenderer:
    """RenderEngine/renderer/OptixRenderer.h:21-44 over the C ABI."""
    PHOTON_GRID_MAX_SIZE = 100 * 100 * 100
    MAX_PHOTON_COUNT = 4
    PHOTON_LAUNCH_WIDTH = 1024
    PHOTON_LAUNCH_HEIGHT = 1024
    EMITTED_PHOTONS_PER_ITERATION = 1024 * 1024
    NUM_PHOTONS = 4 * 1024 * 1024
    VCM_MAX_PATH_LENGTH = 10

    def __init__(self, photon_launch=None, vcm_max_path_length=0, seed=None):
        self._lib = capi.load_library()
        cfg = capi.opr_config()
        if photon_launch:
            cfg.photonLaunchWidth, cfg.photonLaunchHeight = photon_launch
        cfg.vcmMaxPathLength = vcm_max_path_length
        if seed is not None:
            cfg.randomSeed, cfg.seedIsSet = seed, 1
        self.photon_launch = tuple(photon_launch) if photon_launch else (1024, 1024)
        self._h = C.c_void_p()
        self._check(self._lib.opr_create(C.byref(self._h), C.byref(cfg)))
        self._scene = None

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(self._lib.opr_last_error().decode(errors="replace"))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.opr_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def initialize(self, device):
        self._check(self._lib.opr_initialize(self._h, device.getDeviceId()))

    def initScene(self, scene):
        self._scene = scene  # keep host arrays alive during the upload
        self._check(self._lib.opr_init_scene(self._h, scene.getSceneDescription()))

    def renderNextIteration(self, iterationNumber, localIterationNumber, PPMRadius, createOutput, details):
        pod = details.pod() if hasattr(details, "pod") else details
        self._check(self._lib.opr_render_next_iteration(self._h, iterationNumber, localIterationNumber, PPMRadius,
                                                        1 if createOutput else 0, C.byref(pod)))

    def getOutputBuffer(self, data=None):
        if data is None:
            data = np.empty((self.getHeight(), self.getWidth(), 3), np.float32)
        self._check(self._lib.opr_get_output_buffer(self._h, data.ctypes.data_as(C.c_void_p)))
        return data

    def getOutputBufferAsync(self, data):
        """Starts the read-back of the running sum as of the last iteration into `data` (pinned host memory) and returns at once;
        waitOutput() blocks until every started read-back has landed."""
        self._check(self._lib.opr_get_output_buffer_async(self._h, data.ctypes.data_as(C.c_void_p)))
        return data

    def waitOutput(self):
        self._check(self._lib.opr_wait_output(self._h))

    def getWidth(self):
        return self._lib.opr_get_width(self._h)

    def getHeight(self):
        return self._lib.opr_get_height(self._h)

    def getScreenBufferSizeBytes(self):
        return self._lib.opr_get_screen_buffer_size_bytes(self._h)

    # ---- additions (multi-GPU, measurement, parity) ----
    def output_device_ptr(self):
        return self._lib.opr_get_output_device_ptr(self._h)

    def synchronize(self):
        self._check(self._lib.opr_synchronize(self._h))

    # multi-GPU, one process per GPU: the NCCL communicator lives in the C++ library; the launcher only carries the 128-byte id
    def comm_init_rank(self, unique_id, rank, world):
        buf = C.create_string_buffer(bytes(unique_id), 128)
        self._check(self._lib.opr_comm_init_rank(self._h, buf, rank, world))

    def comm_destroy(self):
        self._check(self._lib.opr_comm_destroy(self._h))

    def reduce_output(self, root=0, data=None):
        """ncclReduce(sum) of all ranks' running sums onto `root`, queued on the renderer's stream; with `data` (root only) the
        total is copied to the host and the call blocks."""
        self._check(self._lib.opr_reduce_output(self._h, root, data.ctypes.data_as(C.c_void_p) if data is not None else None))
        return data

    def reduced_device_ptr(self):
        return self._lib.opr_get_reduced_device_ptr(self._h)

    def join(self):
        """Order all queued work (both internal streams) before later work on the renderer's public stream."""
        self._check(self._lib.opr_join_streams(self._h))

    def stats(self):
        a = (C.c_uint64 * 4)()
        self._check(self._lib.opr_get_stats(self._h, a, 4))
        return dict(closest_rays=a[0], any_rays=a[1], photons_stored=a[2], launches=a[3])

    def set_timing(self, on):
        """False / 0: off; True / 1: every pass timed alone (serialised); 2: timeline mode (events only, streams overlap)."""
        self._lib.opr_set_timing(self._h, int(on))

    def timeline(self):
        """[(name, stream, start_ms, end_ms)] of every pass since set_timing(2), on the device's clock."""
        cap = 4096
        a, b, s = (C.c_float * cap)(), (C.c_float * cap)(), C.create_string_buffer(cap)
        n = self._lib.opr_get_timeline(self._h, a, b, s, cap)
        if n < 0:
            raise RuntimeError(self._lib.opr_last_error().decode(errors="replace"))
        return [(self._lib.opr_get_timeline_name(self._h, i).decode(), s.raw[i:i + 1].decode(), a[i], b[i]) for i in range(min(n, cap))]

    def timings(self):
        ms = (C.c_float * 64)()
        n = self._lib.opr_get_timings_ms(self._h, ms, 64)
        return [(self._lib.opr_get_timing_name(self._h, i).decode(), ms[i]) for i in range(min(n, 64))]

    def debug_buffer(self, name, dtype, count):
        a = np.empty(count, dtype)
        self._check(self._lib.opr_debug_get_buffer(self._h, name.encode(), a.ctypes.data_as(C.c_void_p), a.nbytes))
        return a

    def set_trace_nodes(self, quantized):
        self._check(self._lib.opr_debug_set_trace_nodes(self._h, 1 if quantized else 0))

    def set_connection_counts(self, on):
        self._check(self._lib.opr_debug_set_connection_counts(self._h, 1 if on else 0))

    def light_vertex_positions(self):
        cap = self.getWidth() * self.getHeight() * 16 + 1
        a = self.debug_buffer("light_vertex_positions", np.float32, cap * 4).reshape(-1, 4)
        n = int(a[0, 0:1].view(np.uint32)[0])
        a = a[1:1 + min(n, cap - 1)]
        return a[:, 0].copy().view(np.uint32), a[:, 1:4].copy()

    def trace_closest(self, rays8):
        rays8 = np.ascontiguousarray(rays8, np.float32)
        n = rays8.shape[0]
        t, mat, ng = np.empty(n, np.float32), np.empty(n, np.int32), np.empty((n, 3), np.float32)
        PF, PI = C.POINTER(C.c_float), C.POINTER(C.c_int32)
        self._check(self._lib.opr_trace_closest(self._h, rays8.ctypes.data_as(PF), n, t.ctypes.data_as(PF), mat.ctypes.data_as(PI),
                                                ng.ctypes.data_as(PF)))
        return t, mat, ng

    def trace_any(self, rays8):
        rays8 = np.ascontiguousarray(rays8, np.float32)
        n = rays8.shape[0]
        occ = np.empty(n, np.int32)
        self._check(self._lib.opr_trace_any(self._h, rays8.ctypes.data_as(C.POINTER(C.c_float)), n, occ.ctypes.data_as(C.POINTER(C.c_int32))))
        return occ


def comm_unique_id():
    """128-byte NCCL id made by the C++ library (rank 0 calls this; the launcher broadcasts the bytes)."""
    lib = capi.load_library()
    buf = C.create_string_buffer(128)
    if lib.opr_comm_get_unique_id(buf) != 0:
        raise RuntimeError(lib.opr_last_error().decode(errors="replace"))
    return buf.raw


class RendererGroup:
    """One process, G GPUs (opr_group_*): the C++ library owns a renderer and a host thread per device; replaces the client's
    hand-out loop of Client/DistributedApplication.cpp:96-122."""

    def __init__(self, devices, photon_launch=None, vcm_max_path_length=0, seed=None):
        self._lib = capi.load_library()
        cfg = capi.opr_config()
        if photon_launch:
            cfg.photonLaunchWidth, cfg.photonLaunchHeight = photon_launch
        cfg.vcmMaxPathLength = vcm_max_path_length
        if seed is not None:
            cfg.randomSeed, cfg.seedIsSet = seed, 1
        self._h = C.c_void_p()
        devs = (C.c_int32 * len(devices))(*devices)
        self._check(self._lib.opr_group_create(C.byref(self._h), len(devices), devs, C.byref(cfg)))
        self._scene = None

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(self._lib.opr_last_error().decode(errors="replace"))

    def size(self):
        return self._lib.opr_group_size(self._h)

    def initScene(self, scene):
        self._scene = scene
        self._check(self._lib.opr_group_init_scene(self._h, scene.getSceneDescription()))

    def renderNextIterations(self, first, count, initial_radius, alpha, details):
        pod = details.pod() if hasattr(details, "pod") else details
        self._check(self._lib.opr_group_render_next_iterations(self._h, first, count, initial_radius, alpha, C.byref(pod)))

    def reduceOutput(self, data):
        self._check(self._lib.opr_group_reduce_output(self._h, data.ctypes.data_as(C.c_void_p)))
        return data

    def close(self):
        if getattr(self, "_h", None):
            self._lib.opr_group_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
