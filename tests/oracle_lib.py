"""ctypes wrapper around oracle/liboracle.so (the CPU restatement) and, when present,
oracle/_ref/liboppref.so (the reference's own device programs host-compiled behind the shim).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from oppositerenderer_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "liboppref.so")


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR])


def _declare(lib):
    vp, u32, u64, i32, f32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int32, C.c_float
    PF, PU, PI = C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.POINTER(C.c_int32)
    sig = {
        "ora_last_error": (C.c_char_p, []),
        "ora_create": (vp, [C.POINTER(capi.opr_config)]),
        "ora_destroy": (None, [vp]),
        "ora_initialize": (i32, [vp]),
        "ora_init_scene": (i32, [vp, C.POINTER(capi.opr_scene_desc)]),
        "ora_set_threads": (i32, [vp, i32]),
        "ora_render_next_iteration": (i32, [vp, u64, u64, f32, i32, C.POINTER(capi.opr_request_details)]),
        "ora_get_output_buffer": (i32, [vp, vp]),
        "ora_get_width": (u32, [vp]),
        "ora_get_height": (u32, [vp]),
        "ora_get_hitpoints": (i32, [vp, vp]),
        "ora_get_photons": (i32, [vp, vp]),
        "ora_get_photon_keys": (i32, [vp, vp]),
        "ora_get_offsets": (i32, [vp, vp, u32]),
        "ora_get_grid": (i32, [vp, PF, PU, PF, PU]),
        "ora_get_indirect": (i32, [vp, vp]),
        "ora_get_direct": (i32, [vp, vp]),
        "ora_get_random_states": (i32, [vp, vp, u64, u64]),
        "ora_num_random_states": (u64, [vp]),
        "ora_get_ray_counts": (i32, [vp, C.POINTER(u64), C.POINTER(u64)]),
        "ora_get_photon_path_lengths": (i32, [vp, vp]),
        "ora_get_light_vertex_counts": (i32, [vp, vp]),
        "ora_get_photons_as_emitted": (i32, [vp, vp]),
        "ora_get_direct_sample_mask": (i32, [vp, vp]),
        "ora_get_connection_counts": (i32, [vp, vp]),
        "ora_get_pt_signature": (i32, [vp, vp]),
        "ora_get_path_signatures": (i32, [vp, vp]),
        "ora_rng_init": (None, [PU, u32, u32, u32]),
        "ora_rng_uniform": (None, [PU, PF, u32]),
        "ora_build_photon_grid": (i32, [PF, u32, PU, PU, PF, PU, PF, PU]),
        "ora_gather": (i32, [u32, PF, u32, PF, PU, f32, PU, PF, f32, f32, PF]),
        "ora_trace_closest": (i32, [vp, PF, u32, PF, PI, PF]),
        "ora_trace_any": (i32, [vp, PF, u32, PI]),
        "ora_camera_setup": (None, [C.POINTER(capi.opr_camera)]),
        "ora_camera_set_aspect_ratio": (None, [C.POINTER(capi.opr_camera), f32]),
        "ora_make_area_light": (None, [C.POINTER(capi.opr_light), PF, PF, PF, PF]),
        "ora_make_point_light": (None, [C.POINTER(capi.opr_light), PF, PF]),
        "ora_initial_ppm_radius": (f32, [PF, PF]),
        "ora_bounding_sphere": (None, [PF, PF, PF]),
        "ora_max_threads": (i32, []),
        "ora_builtin_scene_create": (vp, [C.c_char_p, u32]),
        "ora_builtin_scene_destroy": (None, [vp]),
        "ora_builtin_scene_desc": (C.POINTER(capi.opr_scene_desc), [vp]),
        "ora_builtin_scene_camera": (None, [vp, C.POINTER(capi.opr_camera)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args


_libs = {}


def load(which="oracle"):
    """which: 'oracle' (restatement) or 'ref' (host-compiled reference programs)."""
    if which in _libs:
        return _libs[which]
    path = ORACLE_SO if which == "oracle" else REF_SO
    if which == "oracle" and not os.path.exists(path):
        build_oracle()
    lib = C.CDLL(path)
    _declare(lib)
    _libs[which] = lib
    return lib


def have_ref():
    return os.path.exists(REF_SO)


PHOTON_DT = np.dtype([("power", "<f4", 3), ("position", "<f4", 3), ("rayDirection", "<f4", 3)])
HITPOINT_DT = np.dtype([("position", "<f4", 3), ("normal", "<f4", 3), ("attenuation", "<f4", 3), ("radiance", "<f4", 3),
                        ("flags", "<u4"), ("volumetricRadiance", "<f4", 3)])
assert PHOTON_DT.itemsize == 36 and HITPOINT_DT.itemsize == 64

PRD_HIT_EMITTER, PRD_ERROR, PRD_MISS = 1 << 31, 1 << 30, 1 << 29
PRD_HIT_SPECULAR, PRD_HIT_NON_SPECULAR = 1 << 28, 1 << 27


def fptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def uptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint32))


def iptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


class OracleRenderer:
    """Same call sequence as the reference's OptixRenderer (RenderEngine/renderer/OptixRenderer.h:21-44)."""

    def __init__(self, which="oracle", photon_launch=(1024, 1024), vcm_max_path_length=10, seed=1645301512, threads=0):
        self.lib = load(which)
        cfg = capi.opr_config(photon_launch[0], photon_launch[1], vcm_max_path_length, seed, 1)
        self.photon_launch = photon_launch
        self.h = self.lib.ora_create(C.byref(cfg))
        if threads:
            self.lib.ora_set_threads(self.h, threads)
        self._chk(self.lib.ora_initialize(self.h))

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError(self.lib.ora_last_error().decode())

    def close(self):
        if self.h:
            self.lib.ora_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def init_scene(self, desc_ptr):
        self._chk(self.lib.ora_init_scene(self.h, desc_ptr))

    def render_next_iteration(self, iteration, local_iteration, radius, details):
        self._chk(self.lib.ora_render_next_iteration(self.h, iteration, local_iteration, radius, 1, C.byref(details)))
        self.width, self.height = details.width, details.height

    def output(self):
        out = np.empty((self.height, self.width, 3), np.float32)
        self.lib.ora_get_output_buffer(self.h, out.ctypes.data_as(C.c_void_p))
        return out

    def hitpoints(self):
        a = np.empty(self.width * self.height, HITPOINT_DT)
        self.lib.ora_get_hitpoints(self.h, a.ctypes.data_as(C.c_void_p))
        return a

    def photons(self):
        a = np.empty(self.photon_launch[0] * self.photon_launch[1] * 4, PHOTON_DT)
        self.lib.ora_get_photons(self.h, a.ctypes.data_as(C.c_void_p))
        return a

    def photons_as_emitted(self):
        """Photon slots before the grid build sorted them (slot = 4 * thread + k), like the product's "photons" buffer."""
        a = np.empty(self.photon_launch[0] * self.photon_launch[1] * 4, PHOTON_DT)
        self.lib.ora_get_photons_as_emitted(self.h, a.ctypes.data_as(C.c_void_p))
        return a

    def path_signatures(self):
        """(light paths, camera paths) signature arrays of the last BDPT iteration."""
        a = np.empty(self.width * self.height * 2, np.uint32)
        self.lib.ora_get_path_signatures(self.h, a.ctypes.data_as(C.c_void_p))
        return a[:self.width * self.height], a[self.width * self.height:]

    def photon_path_lengths(self):
        """Depth reached by every photon thread of the last photon pass (debugPhotonPathLengthBuffer)."""
        a = np.empty(self.photon_launch[0] * self.photon_launch[1], np.uint32)
        self.lib.ora_get_photon_path_lengths(self.h, a.ctypes.data_as(C.c_void_p))
        return a

    def pt_signature(self):
        a = np.empty(self.width * self.height, np.uint32)
        self.lib.ora_get_pt_signature(self.h, a.ctypes.data_as(C.c_void_p))
        return a

    def direct_sample_mask(self):
        a = np.empty(self.width * self.height, np.uint8)
        self.lib.ora_get_direct_sample_mask(self.h, a.ctypes.data_as(C.c_void_p))
        return a

    def connection_counts(self):
        a = np.empty(self.width * self.height, np.uint32)
        self.lib.ora_get_connection_counts(self.h, a.ctypes.data_as(C.c_void_p))
        return a

    def random_states(self):
        n = self.lib.ora_num_random_states(self.h)
        a = np.zeros((n, 6), np.uint32)
        self.lib.ora_get_random_states(self.h, a.ctypes.data_as(C.c_void_p), 0, n)
        return a

    def light_vertex_positions(self):
        """(path u32, hit point float3) of every light vertex stored by the last light pass (restatement only)."""
        fn = self.lib.ora_get_light_vertex_positions
        fn.restype, fn.argtypes = C.c_uint32, [C.c_void_p, C.POINTER(C.c_float), C.c_uint32]
        cap = self.width * self.height * 16
        a = np.zeros((cap, 4), np.float32)
        n = fn(self.h, fptr(a), cap)
        a = a[:min(n, cap)]
        return a[:, 0].copy().view(np.uint32), a[:, 1:4].copy()

    def photon_keys(self):
        a = np.empty(self.photon_launch[0] * self.photon_launch[1] * 4, np.uint32)
        self.lib.ora_get_photon_keys(self.h, a.ctypes.data_as(C.c_void_p))
        return a

    def grid(self):
        cell = C.c_float()
        gs = (C.c_uint32 * 3)()
        org = (C.c_float * 3)()
        nv = C.c_uint32()
        self.lib.ora_get_grid(self.h, C.byref(cell), gs, org, C.byref(nv))
        return dict(cellSize=np.float32(cell.value), gridSize=tuple(gs), origo=np.array(list(org), np.float32), numValid=nv.value)

    def offsets(self):
        g = self.grid()
        n = g["gridSize"][0] * g["gridSize"][1] * g["gridSize"][2] + 1
        a = np.empty(n, np.uint32)
        self.lib.ora_get_offsets(self.h, a.ctypes.data_as(C.c_void_p), n)
        return a

    def indirect(self):
        a = np.empty((self.height, self.width, 3), np.float32)
        self.lib.ora_get_indirect(self.h, a.ctypes.data_as(C.c_void_p))
        return a

    def direct(self):
        a = np.empty((self.height, self.width, 3), np.float32)
        self.lib.ora_get_direct(self.h, a.ctypes.data_as(C.c_void_p))
        return a

    def ray_counts(self):
        a, b = C.c_uint64(), C.c_uint64()
        self.lib.ora_get_ray_counts(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def trace_closest(self, rays8):
        rays8 = np.ascontiguousarray(rays8, np.float32)
        n = rays8.shape[0]
        t = np.empty(n, np.float32)
        mat = np.empty(n, np.int32)
        ng = np.empty((n, 3), np.float32)
        self.lib.ora_trace_closest(self.h, fptr(rays8), n, fptr(t), iptr(mat), fptr(ng))
        return t, mat, ng

    def trace_any(self, rays8):
        rays8 = np.ascontiguousarray(rays8, np.float32)
        n = rays8.shape[0]
        occ = np.empty(n, np.int32)
        self.lib.ora_trace_any(self.h, fptr(rays8), n, iptr(occ))
        return occ


class BuiltinScene:
    def __init__(self, name, config=0, which="oracle"):
        self.lib = load(which)
        self.h = self.lib.ora_builtin_scene_create(name.encode(), config)
        if not self.h:
            raise ValueError("unknown built-in scene %r" % name)
        self.desc = self.lib.ora_builtin_scene_desc(self.h)

    def camera(self, aspect=None):
        cam = capi.opr_camera()
        self.lib.ora_builtin_scene_camera(self.h, C.byref(cam))
        if aspect is not None:
            self.lib.ora_camera_set_aspect_ratio(C.byref(cam), aspect)
        return cam

    def initial_radius(self):
        d = self.desc.contents
        return self.lib.ora_initial_ppm_radius(d.aabbMin, d.aabbMax)


def ppm_radius_schedule(r0, n, alpha=2.0 / 3.0):
    """ST/StandaloneRenderManager.cpp:105-107 — host double: r^2_{i+1} = r^2_i (i+alpha)/(i+1)."""
    out = []
    r = float(r0)
    for i in range(n):
        out.append(r)
        r2 = r * r
        r2 = r2 * (i + alpha) / (i + 1)
        r = r2 ** 0.5
    return out


def make_details(camera, method, width, height, alpha=2.0 / 3.0, scene_name=b""):
    return capi.opr_request_details(camera, method, width, height, alpha, scene_name)
