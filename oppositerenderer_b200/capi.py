"""ctypes mirror of ``include/opr_scene.h`` and ``include/opr.h`` (the C-ABI drop-in boundary).

Layouts follow the reference PODs byte for byte: ``opr_camera`` == ``Camera``
(RenderEngine/renderer/Camera.h:72-79, 96 B), ``opr_light`` == ``Light``
(RenderEngine/renderer/Light.h:27-53, 88 B).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# OPR_LIBRARY: another build of the same library (scripts/build_variant.sh), for A/B timing
LIB_PATH = os.environ.get("OPR_LIBRARY") or os.path.join(_HERE, "csrc", "liboppositerenderer_b200.so")

F3 = C.c_float * 3
F2 = C.c_float * 2

# RenderEngine/renderer/RenderMethod.h:13-19
PATH_TRACING = 0
VCM_BIDIRECTIONAL_PATH_TRACING = 1
PROGRESSIVE_PHOTON_MAPPING = 2
RENDER_METHOD_NONE = 3

KEEP_VERTICAL, KEEP_HORIZONTAL, KEEP_NONE = 0, 1, 2
LIGHT_AREA, LIGHT_POINT, LIGHT_SPOT = 0, 1, 2
MAT_DIFFUSE, MAT_DIFFUSE_EMITTER, MAT_GLASS, MAT_MIRROR, MAT_TEXTURE, MAT_GLOSSY = range(6)


class opr_camera(C.Structure):
    _fields_ = [("eye", F3), ("lookat", F3), ("up", F3), ("hfov", C.c_float), ("vfov", C.c_float),
                ("aperture", C.c_float), ("lookdir", F3), ("camera_u", F3), ("camera_v", F3),
                ("imagePlaneSize", F2), ("aspectRatioMode", C.c_int32)]


class opr_light(C.Structure):
    _fields_ = [("power", F3), ("position", F3), ("v1", F3), ("v2", F3), ("isDelta", C.c_uint8),
                ("isFinite", C.c_uint8), ("_pad", C.c_uint8 * 2), ("inverseArea", C.c_float),
                ("area_or_angle", C.c_float), ("Lemit", F3), ("normal", F3), ("lightType", C.c_int32)]


class opr_material(C.Structure):
    _fields_ = [("kind", C.c_int32), ("Kd", F3), ("Ks", F3), ("exponent", C.c_float), ("Kr", F3), ("Kt", F3),
                ("indexOfRefraction", C.c_float), ("power", F3), ("inverseArea", C.c_float),
                ("diffuseTexture", C.c_int32), ("normalTexture", C.c_int32)]


class opr_texture(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("rgba", C.POINTER(C.c_uint8))]


class opr_mesh(C.Structure):
    _fields_ = [("numVertices", C.c_int32), ("numTriangles", C.c_int32), ("vertices", C.POINTER(C.c_float)),
                ("normals", C.POINTER(C.c_float)), ("uvs", C.POINTER(C.c_float)), ("tangents", C.POINTER(C.c_float)),
                ("bitangents", C.POINTER(C.c_float)), ("indices", C.POINTER(C.c_int32)), ("material", C.c_int32)]


class opr_parallelogram(C.Structure):
    _fields_ = [("anchor", F3), ("offset1", F3), ("offset2", F3), ("material", C.c_int32)]


class opr_sphere(C.Structure):
    _fields_ = [("center", F3), ("radius", C.c_float), ("material", C.c_int32)]


class opr_scene_desc(C.Structure):
    _fields_ = [("name", C.c_char_p),
                ("numMeshes", C.c_int32), ("meshes", C.POINTER(opr_mesh)),
                ("numParallelograms", C.c_int32), ("parallelograms", C.POINTER(opr_parallelogram)),
                ("numSpheres", C.c_int32), ("spheres", C.POINTER(opr_sphere)),
                ("numMaterials", C.c_int32), ("materials", C.POINTER(opr_material)),
                ("numTextures", C.c_int32), ("textures", C.POINTER(opr_texture)),
                ("numLights", C.c_int32), ("lights", C.POINTER(opr_light)),
                ("aabbMin", F3), ("aabbMax", F3)]


class opr_request_details(C.Structure):
    _fields_ = [("camera", opr_camera), ("renderMethod", C.c_int32), ("width", C.c_uint32), ("height", C.c_uint32),
                ("ppmAlpha", C.c_float), ("sceneName", C.c_char_p)]


class opr_config(C.Structure):
    _fields_ = [("photonLaunchWidth", C.c_uint32), ("photonLaunchHeight", C.c_uint32), ("vcmMaxPathLength", C.c_uint32),
                ("randomSeed", C.c_uint32), ("seedIsSet", C.c_uint32)]


assert C.sizeof(opr_camera) == 96
assert C.sizeof(opr_light) == 88

_lib = None


def load_library():
    """Load the CUDA shared library.  There is no CPU fallback: a missing library is an error."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "oppositerenderer_b200: %s is missing — build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU fallback." % LIB_PATH)
    _lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    _declare(_lib)
    return _lib


def _declare(lib):
    vp, u32, u64, i32, f32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int32, C.c_float
    PF, PU, PI = C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.POINTER(C.c_int32)
    sig = {
        "opr_last_error": (C.c_char_p, []),
        "opr_create": (i32, [C.POINTER(vp), C.POINTER(opr_config)]),
        "opr_destroy": (None, [vp]),
        "opr_initialize": (i32, [vp, i32]),
        "opr_init_scene": (i32, [vp, C.POINTER(opr_scene_desc)]),
        "opr_render_next_iteration": (i32, [vp, u64, u64, f32, i32, C.POINTER(opr_request_details)]),
        "opr_get_output_buffer": (i32, [vp, vp]),
        "opr_get_width": (u32, [vp]),
        "opr_get_height": (u32, [vp]),
        "opr_get_screen_buffer_size_bytes": (u32, [vp]),
        "opr_get_output_device_ptr": (vp, [vp]),
        "opr_synchronize": (i32, [vp]),
        "opr_join_streams": (i32, [vp]),
        "opr_get_stream": (vp, [vp]),
        "opr_get_output_buffer_async": (i32, [vp, vp]),
        "opr_wait_output": (i32, [vp]),
        "opr_group_rank_seed": (u32, [u32, i32]),
        "opr_comm_get_unique_id": (i32, [vp]),
        "opr_comm_init_rank": (i32, [vp, vp, i32, i32]),
        "opr_comm_destroy": (i32, [vp]),
        "opr_reduce_output": (i32, [vp, i32, vp]),
        "opr_get_reduced_device_ptr": (vp, [vp]),
        "opr_group_create": (i32, [C.POINTER(vp), i32, PI, C.POINTER(opr_config)]),
        "opr_group_destroy": (None, [vp]),
        "opr_group_size": (i32, [vp]),
        "opr_group_renderer": (vp, [vp, i32]),
        "opr_group_init_scene": (i32, [vp, C.POINTER(opr_scene_desc)]),
        "opr_group_render_next_iterations": (i32, [vp, u64, u64, C.c_double, C.c_double, C.POINTER(opr_request_details)]),
        "opr_group_reduce_output": (i32, [vp, vp]),
        "opr_get_stats": (i32, [vp, C.POINTER(C.c_uint64), i32]),
        "opr_get_timings_ms": (i32, [vp, PF, i32]),
        "opr_get_timing_name": (C.c_char_p, [vp, i32]),
        "opr_set_timing": (i32, [vp, i32]),
        "opr_get_timeline": (i32, [vp, PF, PF, C.c_char_p, i32]),
        "opr_get_timeline_name": (C.c_char_p, [vp, i32]),
        "opr_debug_get_buffer": (i32, [vp, C.c_char_p, vp, u64]),
        "opr_debug_set_connection_counts": (i32, [vp, i32]),
        "opr_debug_set_trace_nodes": (i32, [vp, i32]),
        "opr_debug_slot_pixel": (u32, [u32, u32, u32]),
        "opr_debug_set_guards": (None, [i32]),
        "opr_debug_check_guards": (i32, [C.POINTER(C.c_int)]),
        "opr_rng_init": (i32, [PU, u32, u32, u32]),
        "opr_rng_uniform": (i32, [PU, PF, u32]),
        "opr_build_photon_grid": (i32, [PF, u32, PF, PU, PU, PF, PU, PF, PU]),
        "opr_gather": (i32, [u32, PF, u32, PF, PU, f32, PU, PF, f32, f32, PF]),
        "opr_trace_closest": (i32, [vp, PF, u32, PF, PI, PF]),
        "opr_trace_any": (i32, [vp, PF, u32, PI]),
        "opr_scene_desc_check": (i32, [C.POINTER(opr_scene_desc), C.c_char_p, i32]),
        "opr_builtin_scene_create": (vp, [C.c_char_p, u32]),
        "opr_builtin_scene_destroy": (None, [vp]),
        "opr_builtin_scene_desc": (C.POINTER(opr_scene_desc), [vp]),
        "opr_builtin_scene_camera": (None, [vp, C.POINTER(opr_camera)]),
        "opr_camera_setup": (None, [C.POINTER(opr_camera)]),
        "opr_camera_set_aspect_ratio": (None, [C.POINTER(opr_camera), f32]),
        "opr_make_area_light": (None, [C.POINTER(opr_light), PF, PF, PF, PF]),
        "opr_make_point_light": (None, [C.POINTER(opr_light), PF, PF]),
        "opr_initial_ppm_radius": (f32, [PF, PF]),
        "opr_ppm_radius_for_iteration": (C.c_double, [C.c_double, u64, C.c_double]),
        "opr_io_last_error": (C.c_char_p, []),
        "opr_display_bgr8": (i32, [PF, u32, u32, f32, f32, C.POINTER(C.c_uint8)]),
        "opr_save_bmp": (i32, [C.c_char_p, PF, u32, u32, f32, f32]),
        "opr_image_info": (i32, [C.c_char_p, PI, PI]),
        "opr_image_read": (i32, [C.c_char_p, C.POINTER(C.c_uint8), u64]),
        "opr_scene_file_load": (vp, [C.c_char_p, i32]),
        "opr_scene_file_destroy": (None, [vp]),
        "opr_scene_file_desc": (C.POINTER(opr_scene_desc), [vp]),
        "opr_scene_file_camera": (i32, [vp, C.POINTER(opr_camera)]),
        "opr_scene_file_num_triangles": (u32, [vp]),
        "opr_scene_file_num_meshes": (u32, [vp]),
        "opr_scene_file_warnings": (C.c_char_p, [vp]),
        "opr_scene_file_last_error": (C.c_char_p, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)   # AttributeError here = header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args


EXPORTED_SYMBOLS = None  # filled by tests from include/opr.h
