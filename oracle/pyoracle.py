"""ctypes binding of the CPU oracle (oracle/oracle.c).  TEST INFRASTRUCTURE ONLY.

Importers allowed: tests/, __graft_entry__.smoke(), bench.py (cpu_baseline / --impl reference).
The product package nori_b200 never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_DIR, "_build", "liboracle.so")


def build(force: bool = False) -> str:
    src = [os.path.join(_DIR, f) for f in ("oracle.c", "oracle.h", "Makefile")]
    stale = (not os.path.exists(_LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src)
    if force or stale:
        subprocess.run(["make", "-C", _DIR], check=True, capture_output=True)
    return _LIB_PATH


class Bsdf(C.Structure):
    _fields_ = [("type", C.c_int32), ("albedo", C.c_float * 3), ("alpha", C.c_float),
                ("intIOR", C.c_float), ("extIOR", C.c_float), ("ks", C.c_float)]


class Emitter(C.Structure):
    _fields_ = [("type", C.c_int32), ("radiance", C.c_float * 3)]


class Integrator(C.Structure):
    _fields_ = [("type", C.c_int32), ("rr_start", C.c_int32), ("max_depth", C.c_int32), ("reserved", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("samples", C.c_uint64), ("rays", C.c_uint64), ("node_visits", C.c_uint64),
                ("tri_tests", C.c_uint64), ("seconds", C.c_double)]


class Pcg32(C.Structure):
    _fields_ = [("state", C.c_uint64), ("inc", C.c_uint64)]


RAY_DTYPE = np.dtype([("o", np.float32, 3), ("mint", np.float32), ("d", np.float32, 3), ("maxt", np.float32)])
HIT_DTYPE = np.dtype([("t", np.float32), ("u", np.float32), ("v", np.float32), ("prim", np.uint32), ("mesh", np.uint32)])

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        fp, u32p, vp = C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.c_void_p
        L.orc_scene_create.restype = vp
        L.orc_scene_destroy.argtypes = [vp]
        L.orc_scene_add_mesh.argtypes = [vp, vp, C.c_uint32, vp, vp, vp, C.c_uint32, C.POINTER(Bsdf), C.POINTER(Emitter)]
        L.orc_scene_build.argtypes = [vp]
        L.orc_scene_set_camera.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_float, C.c_float]
        L.orc_scene_set_filter.argtypes = [vp, vp, C.c_float]
        L.orc_scene_set_sampler.argtypes = [vp, C.c_uint32, C.c_int, C.c_uint64]
        L.orc_scene_set_integrator.argtypes = [vp, C.POINTER(Integrator)]
        L.orc_scene_set_point_light.argtypes = [vp, vp, vp]
        L.orc_scene_set_tiles.argtypes = [vp, C.c_int, C.c_int]
        L.orc_render.argtypes = [vp, vp, C.c_int, C.c_int, C.POINTER(Stats)]
        L.orc_intersect.argtypes = [vp, vp, C.c_uint64, vp, C.c_int, C.c_int, C.POINTER(Stats)]
        L.orc_intersect_full.argtypes = [vp, vp, C.c_uint64, vp, C.c_int]
        L.orc_ttest_scene.argtypes = [vp, C.c_uint64, C.c_int, vp]
        L.orc_li_samples.argtypes = [vp, C.c_uint64, C.c_int, vp]
        L.orc_film_to_rgb.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp]
        L.orc_block_order.argtypes = [C.c_int, C.c_int, C.c_int, vp]
        L.orc_pcg32_init.argtypes = [C.POINTER(Pcg32)]
        L.orc_pcg32_seed.argtypes = [C.POINTER(Pcg32), C.c_uint64, C.c_uint64]
        L.orc_pcg32_next_uint.argtypes = [C.POINTER(Pcg32)]; L.orc_pcg32_next_uint.restype = C.c_uint32
        L.orc_pcg32_next_float.argtypes = [C.POINTER(Pcg32)]; L.orc_pcg32_next_float.restype = C.c_float
        L.orc_pcg32_advance.argtypes = [C.POINTER(Pcg32), C.c_int64]
        L.orc_sincos2pi.argtypes = [C.c_float, fp, fp]
        L.orc_logf.argtypes = [C.c_float]; L.orc_logf.restype = C.c_float
        L.orc_expf.argtypes = [C.c_float]; L.orc_expf.restype = C.c_float
        L.orc_square_to_cosine_hemisphere.argtypes = [vp, vp]
        L.orc_square_to_cosine_hemisphere_pdf.argtypes = [vp]; L.orc_square_to_cosine_hemisphere_pdf.restype = C.c_float
        L.orc_square_to_beckmann.argtypes = [vp, C.c_float, vp]
        L.orc_square_to_beckmann_pdf.argtypes = [vp, C.c_float]; L.orc_square_to_beckmann_pdf.restype = C.c_float
        L.orc_fresnel.argtypes = [C.c_float] * 3; L.orc_fresnel.restype = C.c_float
        L.orc_coordinate_system.argtypes = [vp, vp, vp]
        L.orc_bsdf_sample.argtypes = [C.POINTER(Bsdf), vp, vp, vp, fp, C.POINTER(C.c_int), vp]
        L.orc_bsdf_eval.argtypes = [C.POINTER(Bsdf), vp, vp, vp]
        L.orc_bsdf_pdf.argtypes = [C.POINTER(Bsdf), vp, vp]; L.orc_bsdf_pdf.restype = C.c_float
        L.orc_bsdf_sample_batch.argtypes = [C.POINTER(Bsdf), vp, C.c_uint64, C.POINTER(Pcg32), vp, vp]
        L.orc_bsdf_eval_pdf_batch.argtypes = [C.POINTER(Bsdf), vp, vp, C.c_uint64, vp]
        L.orc_filter_table.argtypes = [C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, vp, fp]
        L.orc_camera_matrices.argtypes = [C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, vp]
        L.orc_sample_ray.argtypes = [vp, C.c_float, C.c_float, vp]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def bsdf_struct(b) -> Bsdf:
    s = Bsdf()
    s.type = int(b.type)
    for i in range(3):
        s.albedo[i] = float(b.albedo[i])
    s.alpha, s.intIOR, s.extIOR, s.ks = float(b.alpha), float(b.intIOR), float(b.extIOR), float(b.ks)
    return s


class OracleScene:
    """Builds an orc_scene from a nori_b200.scene.Scene description."""

    def __init__(self, scene):
        L = lib()
        self.scene = scene
        self.h = L.orc_scene_create()
        for m in scene.meshes:
            b = bsdf_struct(m.bsdf)
            e = Emitter()
            if m.radiance is not None:
                e.type = 1
                for i in range(3):
                    e.radiance[i] = float(m.radiance[i])
            L.orc_scene_add_mesh(self.h, _p(m.V), m.V.shape[0], _p(m.N), _p(m.UV), _p(m.F), m.F.shape[0],
                                 C.byref(b), C.byref(e))
        L.orc_scene_build(self.h)
        self.update(scene)

    def update(self, scene):
        L = lib()
        self.scene = scene
        cam = scene.camera
        s2c = np.ascontiguousarray(cam.s2c, dtype=np.float32)
        c2w = np.ascontiguousarray(cam.c2w, dtype=np.float32)
        L.orc_scene_set_camera(self.h, _p(s2c), _p(c2w), cam.width, cam.height, cam.nearClip, cam.farClip)
        tab = np.ascontiguousarray(scene.filter_table, dtype=np.float32)
        L.orc_scene_set_filter(self.h, _p(tab), scene.filter_radius)
        L.orc_scene_set_sampler(self.h, scene.spp, scene.seed_mode, scene.seed)
        it = Integrator(int(scene.integrator), int(scene.rr_start), int(scene.max_depth), 0)
        L.orc_scene_set_integrator(self.h, C.byref(it))
        if getattr(scene, "light_pos", None) is not None:
            lp = np.ascontiguousarray(scene.light_pos, dtype=np.float32)
            le = np.ascontiguousarray(scene.light_energy, dtype=np.float32)
            L.orc_scene_set_point_light(self.h, _p(lp), _p(le))

    def close(self):
        if self.h:
            lib().orc_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_tiles(self, rank, nranks):
        lib().orc_scene_set_tiles(self.h, rank, nranks)

    def render(self, accel=1, nthreads=None):
        nthreads = nthreads or os.cpu_count() or 1
        film = np.zeros(self.scene.film_shape, dtype=np.float32)
        st = Stats()
        lib().orc_render(self.h, _p(film), accel, nthreads, C.byref(st))
        return film, st

    def intersect(self, rays: np.ndarray, shadow=False, accel=1):
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        hits = np.zeros(rays.shape[0], dtype=HIT_DTYPE)
        st = Stats()
        lib().orc_intersect(self.h, _p(rays), rays.shape[0], _p(hits), int(shadow), accel, C.byref(st))
        return hits, st

    def intersect_full(self, rays: np.ndarray, accel=1):
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        out = np.zeros((rays.shape[0], 16), dtype=np.float32)
        lib().orc_intersect_full(self.h, _p(rays), rays.shape[0], _p(out), accel)
        return out

    def ttest_lum(self, n, accel=1):
        lum = np.zeros(n, dtype=np.float64)
        lib().orc_ttest_scene(self.h, n, accel, _p(lum))
        return lum

    def li_samples(self, n, accel=1):
        lum = np.zeros(n, dtype=np.float32)
        lib().orc_li_samples(self.h, n, accel, _p(lum))
        return lum

    def sample_ray(self, sx, sy):
        r = np.zeros(1, dtype=RAY_DTYPE)
        lib().orc_sample_ray(self.h, sx, sy, _p(r))
        return r[0]


def film_to_rgb(film: np.ndarray, W, H, border):
    rgb = np.zeros((H, W, 3), dtype=np.float32)
    film = np.ascontiguousarray(film, dtype=np.float32)
    lib().orc_film_to_rgb(_p(film), W, H, border, _p(rgb))
    return rgb
