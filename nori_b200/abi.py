"""ctypes binding of the C-ABI in include/nori_b200.h (libnori_b200.so).

The library is the product: there is no Python or CPU fallback behind these calls.  If the shared
object is missing or no CUDA device is usable the functions raise -- loudly.
"""
from __future__ import annotations

import ctypes as C
import os
import re

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NORI_B200_LIB") or os.path.join(_HERE, "lib", "libnori_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "nori_b200.h")


class NoriError(RuntimeError):
    """Mirror of NoriException (ref: include/nori/common.h:135-140) for errors crossing the C-ABI."""


class BsdfDesc(C.Structure):
    _fields_ = [("type", C.c_int32), ("albedo", C.c_float * 3), ("alpha", C.c_float),
                ("intIOR", C.c_float), ("extIOR", C.c_float), ("ks", C.c_float)]


class EmitterDesc(C.Structure):
    _fields_ = [("type", C.c_int32), ("radiance", C.c_float * 3)]


class IntegratorDesc(C.Structure):
    _fields_ = [("type", C.c_int32), ("rr_start", C.c_int32), ("max_depth", C.c_int32), ("reserved", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("samples", C.c_uint64), ("rays", C.c_uint64), ("node_visits", C.c_uint64), ("tri_tests", C.c_uint64),
                ("hits_shaded", C.c_uint64), ("kernel_ms", C.c_double), ("total_ms", C.c_double),
                ("launches", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


RAY_DTYPE = np.dtype([("o", np.float32, 3), ("mint", np.float32), ("d", np.float32, 3), ("maxt", np.float32)])
HIT_DTYPE = np.dtype([("t", np.float32), ("u", np.float32), ("v", np.float32), ("prim", np.uint32), ("mesh", np.uint32)])

_lib = None


def declared_symbols():
    """Every entry point include/nori_b200.h declares."""
    src = open(HEADER_PATH).read()
    return sorted(set(re.findall(r"\b(nb_[a-z_0-9]+)\s*\(", src)) - {"nb_ctx"})


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NoriError(f"{LIB_PATH} is missing: run `python -m nori_b200.build` (there is no fallback path)")
        L = C.CDLL(LIB_PATH)
        vp, i, u32, u64, f = C.c_void_p, C.c_int, C.c_uint32, C.c_uint64, C.c_float
        sp = C.POINTER(Stats)
        L.nb_create.argtypes = [i]; L.nb_create.restype = vp
        L.nb_create_multi.argtypes = [C.POINTER(i), i]; L.nb_create_multi.restype = vp
        L.nb_device_count.argtypes = [vp]
        L.nb_comm_get_unique_id.argtypes = [vp]
        L.nb_comm_init_rank.argtypes = [vp, vp, i, i]
        L.nb_render_gather.argtypes = [vp, vp, vp, sp]
        L.nb_destroy.argtypes = [vp]; L.nb_destroy.restype = None
        L.nb_last_error.restype = C.c_char_p
        L.nb_abi_version.restype = i
        L.nb_node_bytes.restype = i
        L.nb_add_mesh.argtypes = [vp, vp, u32, vp, vp, vp, u32, C.POINTER(BsdfDesc), C.POINTER(EmitterDesc)]
        L.nb_clear_meshes.argtypes = [vp]
        L.nb_build_accel.argtypes = [vp]
        L.nb_upload_scene.argtypes = [vp]
        L.nb_set_accel_cache.argtypes = [vp, C.c_char_p]
        L.nb_accel_cache_hit.argtypes = [vp]
        L.nb_debug_build_wide.argtypes = [vp, vp, u32, vp, u64, vp, u64, vp]
        L.nb_debug_wide_intersect.argtypes = [vp, u32, vp, vp, u64, i, vp, vp]
        L.nb_debug_tile_order.argtypes = [i, i, i, vp, u64]
        L.nb_debug_unit_plan.argtypes = [i, u32, C.c_int64, C.c_int64, C.c_int64, C.c_int64, vp]
        L.nb_debug_bvh_cache.argtypes = [vp, vp, u32, i, C.c_int64, C.c_char_p, vp, u64, vp, u64, vp]
        L.nb_build_stats.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(i)]
        L.nb_set_camera.argtypes = [vp, vp, vp, i, i, f, f]
        L.nb_set_filter.argtypes = [vp, vp, f]
        L.nb_set_sampler.argtypes = [vp, u32, i, u64]
        L.nb_set_integrator.argtypes = [vp, C.POINTER(IntegratorDesc)]
        L.nb_set_point_light.argtypes = [vp, vp, vp]
        L.nb_set_tiles.argtypes = [vp, i, i]
        L.nb_render.argtypes = [vp, vp, sp]
        L.nb_render_device.argtypes = [vp, vp, vp, sp]
        L.nb_last_kernel_ms.argtypes = [vp, C.POINTER(C.c_double)]
        L.nb_render_blocks_device.argtypes = [vp, vp, vp, sp]
        L.nb_tile_count.argtypes = [vp, i, i, C.POINTER(i), C.POINTER(i)]
        L.nb_merge_blocks_device.argtypes = [vp, vp, i, i, vp, vp]
        L.nb_merge_all_blocks_device.argtypes = [vp, vp, i, i, vp, vp]
        L.nb_li_samples.argtypes = [vp, u64, vp, sp]
        L.nb_bsdf_sample.argtypes = [vp, C.POINTER(BsdfDesc), vp, i, vp, u64, vp]
        L.nb_bsdf_eval_pdf.argtypes = [vp, C.POINTER(BsdfDesc), vp, i, vp, u64, vp]
        L.nb_intersect.argtypes = [vp, vp, u64, vp, i, sp]
        L.nb_intersect_device.argtypes = [vp, vp, u64, vp, i, vp, sp]
        L.nb_intersect_full.argtypes = [vp, vp, u64, vp]
        L.nb_film_to_rgb.argtypes = [vp, vp, vp]
        L.nb_last_film_to_srgb8.argtypes = [vp, vp]
        L.nb_render_begin.argtypes = [vp]
        L.nb_render_pass.argtypes = [vp, u32, sp]
        L.nb_render_preview.argtypes = [vp, vp, vp]
        L.nb_render_end.argtypes = [vp]
        L.nb_set_option.argtypes = [vp, C.c_char_p, C.c_int64]
        L.nb_debug_counters.argtypes = [vp, vp]
        L.nb_debug_build_bvh.argtypes = [vp, vp, u32, i, C.c_int64, vp, u64, vp, u64, vp]
        L.nb_scene_info.argtypes = [vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64), C.POINTER(i)]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _check(rc):
    if rc != 0:
        raise NoriError(lib().nb_last_error().decode())


def debug_build_bvh(V: np.ndarray, F: np.ndarray, max_leaf=3, bfs_nodes=2048):
    """Runs the product's host SAH builder without a GPU (nb_debug_build_bvh).  V (nv,3) float32, F (nf,3) uint32.
    Returns (nodes [n,16] float32, tris [m,12] float32, info dict)."""
    L = lib()
    v4 = np.zeros((V.shape[0], 4), dtype=np.float32); v4[:, :3] = V
    f4 = np.zeros((F.shape[0], 4), dtype=np.uint32); f4[:, :3] = F
    info = np.zeros(4, dtype=np.uint32)
    _check_rc = L.nb_debug_build_bvh(_p(v4), _p(f4), F.shape[0], max_leaf, bfs_nodes, None, 0, None, 0, _p(info))
    if _check_rc:
        raise NoriError(f"nb_debug_build_bvh failed ({_check_rc})")
    nodes = np.zeros((int(info[0]), 16), dtype=np.float32); tris = np.zeros((int(info[1]), 12), dtype=np.float32)
    rc = L.nb_debug_build_bvh(_p(v4), _p(f4), F.shape[0], max_leaf, bfs_nodes, _p(nodes), nodes.size, _p(tris), tris.size, _p(info))
    if rc:
        raise NoriError(f"nb_debug_build_bvh failed ({rc})")
    return nodes, tris, dict(nodes=int(info[0]), tris=int(info[1]), top_nodes=int(info[2]), depth=int(info[3]))


def debug_build_wide(V: np.ndarray, F: np.ndarray):
    """The 8-wide compressed hierarchy of the product (nb_wide.h), built without a GPU: (nodes [n,20] uint32, tris [m,12] float32, info)."""
    L = lib()
    v4 = np.zeros((V.shape[0], 4), dtype=np.float32); v4[:, :3] = V
    f4 = np.zeros((F.shape[0], 4), dtype=np.uint32); f4[:, :3] = F
    info = np.zeros(4, dtype=np.uint32)
    rc = L.nb_debug_build_wide(_p(v4), _p(f4), F.shape[0], None, 0, None, 0, _p(info))
    if rc:
        raise NoriError(f"nb_debug_build_wide failed ({rc})")
    nodes = np.zeros((int(info[0]), 20), dtype=np.uint32); tris = np.zeros((int(info[1]), 12), dtype=np.float32)
    rc = L.nb_debug_build_wide(_p(v4), _p(f4), F.shape[0], _p(nodes), nodes.size, _p(tris), tris.size, _p(info))
    if rc:
        raise NoriError(f"nb_debug_build_wide failed ({rc})")
    return nodes, tris, dict(nodes=int(info[0]), tris=int(info[1]), depth=int(info[2]), binary_nodes=int(info[3]))


def debug_wide_intersect(nodes: np.ndarray, tris: np.ndarray, rays: np.ndarray, any_hit=False):
    """Host reference walk over a wide hierarchy with the kernels' node step: rays RAY_DTYPE -> (hits HIT_DTYPE-like t,u,v,prim; counts)."""
    rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
    out = np.zeros((rays.shape[0], 4), dtype=np.float32)
    counts = np.zeros(2, dtype=np.uint64)
    rc = lib().nb_debug_wide_intersect(_p(nodes), nodes.shape[0], _p(tris), _p(rays), rays.shape[0], int(any_hit), _p(out), _p(counts))
    if rc:
        raise NoriError(f"nb_debug_wide_intersect failed ({rc})")
    return out, counts


def debug_tile_order(width, height, nranks):
    """nb_debug_tile_order (no GPU): [(bx, by)] in the device path's tile numbering for a group of `nranks`."""
    L = lib()
    n = ((width + 31) // 32) * ((height + 31) // 32)
    out = np.zeros(n, dtype=np.uint32)
    rc = L.nb_debug_tile_order(width, height, nranks, _p(out), out.size)
    if rc != 0:
        raise NoriError(f"nb_debug_tile_order failed ({rc})")
    return [(int(v & 0xffff), int(v >> 16)) for v in out]


def debug_unit_plan(n_tiles, spp, resident_warps, chunk=0, guided=-1, coarse=0):
    """nb_debug_unit_plan (no GPU): the work-unit schedule render_blocks hands the render kernel."""
    out = np.zeros(7, dtype=np.uint32)
    rc = lib().nb_debug_unit_plan(n_tiles, spp, resident_warps, chunk, guided, coarse, _p(out))
    if rc != 0:
        raise NoriError(f"nb_debug_unit_plan failed ({rc})")
    return dict(zip(("chunk", "nchunks", "split_sample", "chunk_a", "nchunks_a", "split_units", "n_units"), (int(v) for v in out)))


def debug_bvh_cache(V: np.ndarray, F: np.ndarray, path: str, max_leaf=3, bfs_nodes=2048):
    """nb_debug_bvh_cache (no GPU): the load-or-build-and-save step of nb_build_accel with nb_set_accel_cache.
    Returns (nodes, tris, info) with info["hit"]."""
    L = lib()
    v4 = np.zeros((V.shape[0], 4), dtype=np.float32); v4[:, :3] = V
    f4 = np.zeros((F.shape[0], 4), dtype=np.uint32); f4[:, :3] = F
    info = np.zeros(5, dtype=np.uint32)
    rc = L.nb_debug_bvh_cache(_p(v4), _p(f4), F.shape[0], max_leaf, bfs_nodes, os.fspath(path).encode(), None, 0, None, 0, _p(info))
    if rc:
        raise NoriError(f"nb_debug_bvh_cache failed ({rc})")
    nodes = np.zeros((int(info[0]), 16), dtype=np.float32); tris = np.zeros((int(info[1]), 12), dtype=np.float32)
    first_hit = int(info[4])
    rc = L.nb_debug_bvh_cache(_p(v4), _p(f4), F.shape[0], max_leaf, bfs_nodes, os.fspath(path).encode(), _p(nodes), nodes.size, _p(tris), tris.size, _p(info))
    if rc:
        raise NoriError(f"nb_debug_bvh_cache failed ({rc})")
    return nodes, tris, dict(nodes=int(info[0]), tris=int(info[1]), top_nodes=int(info[2]), depth=int(info[3]), hit=bool(first_hit))


class Context:
    """One GPU's render context (nb_ctx).  Mirrors the role of Scene's Accel + render() in the reference."""

    def __init__(self, device=-1):
        """device: one CUDA device index (nb_create), or a list of them -- ONE context driving all of them, tiles sharded
        tile_id % N and gathered over NCCL behind nb_render (nb_create_multi)."""
        L = lib()
        if isinstance(device, (list, tuple)):
            arr = (C.c_int * len(device))(*[int(d) for d in device])
            self.h = L.nb_create_multi(arr, len(device))
        else:
            self.h = L.nb_create(int(device))
        if not self.h:
            raise NoriError(L.nb_last_error().decode())
        self.scene = None

    @property
    def device_count(self) -> int:
        return int(lib().nb_device_count(self.h))

    # ---- one process per GPU (torchrun): attach this context to an NCCL group
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = (C.c_uint8 * 128)()
        _check(lib().nb_comm_get_unique_id(buf))
        return bytes(buf)

    def comm_init_rank(self, uid: bytes, rank: int, nranks: int):
        assert len(uid) == 128
        buf = (C.c_uint8 * 128).from_buffer_copy(uid)
        _check(lib().nb_comm_init_rank(self.h, buf, rank, nranks))

    def render_gather(self, film_ptr: int, stream: int = 0, want_stats: bool = True):
        """nb_render_gather: this rank's tile shard -> ONE NCCL gather of finished blocks on rank 0 -> ONE merge launch there.
        film_ptr: device film on rank 0 (0 elsewhere).  want_stats=False only enqueues."""
        if not want_stats:
            _check(lib().nb_render_gather(self.h, C.c_void_p(film_ptr) if film_ptr else None, C.c_void_p(stream), None))
            return None
        st = Stats()
        _check(lib().nb_render_gather(self.h, C.c_void_p(film_ptr) if film_ptr else None, C.c_void_p(stream), C.byref(st)))
        return st

    def close(self):
        if getattr(self, "h", None):
            lib().nb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- scene assembly (Scene::addChild / activate, ref: src/scene.cpp:27-79)
    def set_option(self, key: str, value: int):
        _check(lib().nb_set_option(self.h, key.encode(), int(value)))

    def load(self, scene, build=True):
        L = lib()
        _check(L.nb_clear_meshes(self.h))
        for m in scene.meshes:
            b = BsdfDesc()
            b.type = int(m.bsdf.type)
            for k in range(3):
                b.albedo[k] = float(m.bsdf.albedo[k])
            b.alpha, b.intIOR, b.extIOR, b.ks = float(m.bsdf.alpha), float(m.bsdf.intIOR), float(m.bsdf.extIOR), float(m.bsdf.ks)
            e = EmitterDesc()
            if m.radiance is not None:
                e.type = 1
                for k in range(3):
                    e.radiance[k] = float(m.radiance[k])
            if L.nb_add_mesh(self.h, _p(m.V), m.V.shape[0], _p(m.N), _p(m.UV), _p(m.F), m.F.shape[0], C.byref(b), C.byref(e)) < 0:
                raise NoriError(L.nb_last_error().decode())
        if build:
            _check(L.nb_build_accel(self.h))
        self.configure(scene)

    def configure(self, scene):
        """Camera / filter / sampler / integrator state (cheap; no geometry work)."""
        L = lib()
        self.scene = scene
        cam = scene.camera
        s2c = np.ascontiguousarray(cam.s2c, dtype=np.float32)
        c2w = np.ascontiguousarray(cam.c2w, dtype=np.float32)
        _check(L.nb_set_camera(self.h, _p(s2c), _p(c2w), cam.width, cam.height, cam.nearClip, cam.farClip))
        tab = np.ascontiguousarray(scene.filter_table, dtype=np.float32)
        _check(L.nb_set_filter(self.h, _p(tab), scene.filter_radius))
        _check(L.nb_set_sampler(self.h, scene.spp, scene.seed_mode, scene.seed))
        it = IntegratorDesc(int(scene.integrator), int(scene.rr_start), int(scene.max_depth), 0)
        _check(L.nb_set_integrator(self.h, C.byref(it)))
        if getattr(scene, "light_pos", None) is not None:      # `simple` integrator only
            lp = np.ascontiguousarray(scene.light_pos, dtype=np.float32)
            le = np.ascontiguousarray(scene.light_energy, dtype=np.float32)
            _check(L.nb_set_point_light(self.h, _p(lp), _p(le)))

    def set_accel_cache(self, path):
        _check(lib().nb_set_accel_cache(self.h, os.fspath(path).encode() if path else None))

    @property
    def accel_cache_hit(self) -> bool:
        return bool(lib().nb_accel_cache_hit(self.h))

    def build_stats(self):
        sec, b = C.c_double(), C.c_int()
        _check(lib().nb_build_stats(self.h, C.byref(sec), C.byref(b)))
        return dict(seconds=sec.value, builder="device-lbvh" if b.value == 1 else "host-sah")

    def upload(self):
        _check(lib().nb_upload_scene(self.h))

    def set_tiles(self, rank, nranks):
        _check(lib().nb_set_tiles(self.h, rank, nranks))

    def tile_count(self, rank, nranks):
        n, e = C.c_int(), C.c_int()
        _check(lib().nb_tile_count(self.h, rank, nranks, C.byref(n), C.byref(e)))
        return n.value, e.value

    def debug_counters(self):
        out = np.zeros(8, dtype=np.uint64)
        _check(lib().nb_debug_counters(self.h, _p(out)))
        return out

    def scene_info(self):
        a, b, c, d = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_int()
        _check(lib().nb_scene_info(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return dict(tris=a.value, nodes=b.value, bytes=c.value, depth=d.value)

    # ---- the path
    def render(self, film: np.ndarray | None = None):
        """nb_render: host film out (H+2b, W+2b, 4) -- the reference-facing call."""
        if film is None:
            film = np.zeros(self.scene.film_shape, dtype=np.float32)
        st = Stats()
        _check(lib().nb_render(self.h, _p(film), C.byref(st)))
        return film, st

    def render_host_ptr(self, host_ptr: int):
        st = Stats()
        _check(lib().nb_render(self.h, C.c_void_p(host_ptr), C.byref(st)))
        return st

    def render_device(self, film_ptr: int, stream: int = 0, want_stats: bool = True):
        if not want_stats:
            _check(lib().nb_render_device(self.h, C.c_void_p(film_ptr), C.c_void_p(stream), None))
            return None
        st = Stats()
        _check(lib().nb_render_device(self.h, C.c_void_p(film_ptr), C.c_void_p(stream), C.byref(st)))
        return st

    def last_kernel_ms(self) -> float:
        ms = C.c_double()
        _check(lib().nb_last_kernel_ms(self.h, C.byref(ms)))
        return ms.value

    def render_blocks_device(self, blocks_ptr: int, stream: int = 0, want_stats: bool = True):
        """want_stats=False only enqueues the work on `stream` (no counter read-back, no host synchronisation)."""
        if not want_stats:
            _check(lib().nb_render_blocks_device(self.h, C.c_void_p(blocks_ptr), C.c_void_p(stream), None))
            return None
        st = Stats()
        _check(lib().nb_render_blocks_device(self.h, C.c_void_p(blocks_ptr), C.c_void_p(stream), C.byref(st)))
        return st

    def merge_blocks_device(self, blocks_ptr: int, rank: int, nranks: int, film_ptr: int, stream: int = 0):
        _check(lib().nb_merge_blocks_device(self.h, C.c_void_p(blocks_ptr), rank, nranks, C.c_void_p(film_ptr), C.c_void_p(stream)))

    def merge_all_blocks_device(self, blocks_ptr: int, nranks: int, stride_tiles: int, film_ptr: int, stream: int = 0):
        _check(lib().nb_merge_all_blocks_device(self.h, C.c_void_p(blocks_ptr), nranks, stride_tiles, C.c_void_p(film_ptr), C.c_void_p(stream)))

    def li_samples(self, n: int):
        """nb_li_samples: luminance of Li for n independent camera paths (t-test scene mode, ref: src/ttest.cpp:153-167)."""
        lum = np.zeros(n, dtype=np.float32)
        st = Stats()
        _check(lib().nb_li_samples(self.h, n, _p(lum), C.byref(st)))
        return lum, st

    def bsdf_sample(self, bsdf: "BsdfDesc", wi: np.ndarray, xi: np.ndarray):
        """nb_bsdf_sample: wi (3,) shared or (n,3); xi (n,2) -> (n,8) = wo, weight, pdf, measure."""
        wi = np.ascontiguousarray(wi, dtype=np.float32); xi = np.ascontiguousarray(xi, dtype=np.float32)
        out = np.zeros((xi.shape[0], 8), dtype=np.float32)
        _check(lib().nb_bsdf_sample(self.h, C.byref(bsdf), _p(wi), int(wi.ndim == 2), _p(xi), xi.shape[0], _p(out)))
        return out

    def bsdf_eval_pdf(self, bsdf: "BsdfDesc", wi: np.ndarray, wo: np.ndarray):
        """nb_bsdf_eval_pdf: wi (3,) shared or (n,3); wo (n,3) -> (n,4) = eval rgb, pdf."""
        wi = np.ascontiguousarray(wi, dtype=np.float32); wo = np.ascontiguousarray(wo, dtype=np.float32)
        out = np.zeros((wo.shape[0], 4), dtype=np.float32)
        _check(lib().nb_bsdf_eval_pdf(self.h, C.byref(bsdf), _p(wi), int(wi.ndim == 2), _p(wo), wo.shape[0], _p(out)))
        return out

    def intersect(self, rays: np.ndarray, shadow=False):
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        hits = np.zeros(rays.shape[0], dtype=HIT_DTYPE)
        st = Stats()
        _check(lib().nb_intersect(self.h, _p(rays), rays.shape[0], _p(hits), int(shadow), C.byref(st)))
        return hits, st

    def intersect_device(self, rays_ptr: int, n: int, hits_ptr: int, shadow=False, stream: int = 0):
        st = Stats()
        _check(lib().nb_intersect_device(self.h, C.c_void_p(rays_ptr), n, C.c_void_p(hits_ptr), int(shadow), C.c_void_p(stream), C.byref(st)))
        return st

    def intersect_full(self, rays: np.ndarray):
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        out = np.zeros((rays.shape[0], 16), dtype=np.float32)
        _check(lib().nb_intersect_full(self.h, _p(rays), rays.shape[0], _p(out)))
        return out

    def render_progressive(self, samples_per_pass: int, want_rgb8: bool = False):
        """Generator over the passes of a progressive frame (nb_render_begin / _pass / _preview / _end): yields
        (samples_done, film, rgb8 or None, stats) after every pass; the last film is the full frame."""
        _check(lib().nb_render_begin(self.h))
        try:
            done, spp = 0, int(self.scene.spp)
            while done < spp:
                n = min(int(samples_per_pass), spp - done)
                st = Stats()
                _check(lib().nb_render_pass(self.h, n, C.byref(st)))
                done += n
                film = np.zeros(self.scene.film_shape, dtype=np.float32)
                rgb8 = np.zeros((self.scene.camera.height, self.scene.camera.width, 3), dtype=np.uint8) if want_rgb8 else None
                _check(lib().nb_render_preview(self.h, _p(film), _p(rgb8)))
                yield done, film, rgb8, st
        finally:
            lib().nb_render_end(self.h)

    def last_film_to_srgb8(self):
        """nb_last_film_to_srgb8: the last nb_render's film, normalised + sRGB-tonemapped + quantised on the device: (H, W, 3) uint8."""
        out = np.zeros((self.scene.camera.height, self.scene.camera.width, 3), dtype=np.uint8)
        _check(lib().nb_last_film_to_srgb8(self.h, _p(out)))
        return out

    def film_to_rgb(self, film: np.ndarray):
        film = np.ascontiguousarray(film, dtype=np.float32)
        rgb = np.zeros((self.scene.camera.height, self.scene.camera.width, 3), dtype=np.float32)
        _check(lib().nb_film_to_rgb(self.h, _p(film), _p(rgb)))
        return rgb
