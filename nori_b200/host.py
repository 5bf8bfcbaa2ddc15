"""ctypes binding of the C++ host mirror of Nori (libnori_host.so): NoriObject registry, XML scene pipeline,
plugins, and render() on top of the C-ABI.  Used by the tests and the bench; the CLI is nori_b200/lib/nori."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import abi
from . import scene as S

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libnori_host.so")
CLI_PATH = os.path.join(_HERE, "lib", "nori")
_lib = None


def lib():
    global _lib
    if _lib is None:
        abi.lib()   # libnori_host links against libnori_b200 (rpath $ORIGIN)
        L = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        L.nori_host_last_error.restype = C.c_char_p
        L.nori_host_load.argtypes = [C.c_char_p]; L.nori_host_load.restype = vp
        L.nori_host_free.argtypes = [vp]; L.nori_host_free.restype = None
        L.nori_host_info.argtypes = [vp, vp]
        L.nori_host_camera.argtypes = [vp, vp]
        L.nori_host_mesh.argtypes = [vp, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(vp), C.POINTER(vp),
                                     C.POINTER(vp), C.POINTER(vp), C.POINTER(abi.BsdfDesc), C.POINTER(abi.EmitterDesc)]
        L.nori_host_render.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, C.POINTER(abi.Stats)]
        L.nori_host_render_gpus.argtypes = [vp, C.c_int, C.c_int, C.c_char_p, vp, vp, vp, C.POINTER(abi.Stats)]
        L.nori_host_is_registered.argtypes = [C.c_char_p]
        L.nori_host_block_order.argtypes = [C.c_int, C.c_int, vp]
        _lib = L
    return _lib


def is_registered(name: str) -> bool:
    return bool(lib().nori_host_is_registered(name.encode()))


def block_order(W, H):
    n = lib().nori_host_block_order(W, H, None)
    xy = np.zeros((n, 4), dtype=np.int32)
    lib().nori_host_block_order(W, H, xy.ctypes.data_as(C.c_void_p))
    return xy


class HostScene:
    """A scene parsed by the C++ XML pipeline (loadFromXML)."""

    def __init__(self, xml_path: str):
        L = lib()
        self.h = L.nori_host_load(os.fspath(xml_path).encode())
        if not self.h:
            raise abi.NoriError(L.nori_host_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            lib().nori_host_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self):
        out = np.zeros(8, dtype=np.int64)
        if lib().nori_host_info(self.h, out.ctypes.data_as(C.c_void_p)):
            raise abi.NoriError(lib().nori_host_last_error().decode())
        keys = ["width", "height", "border", "spp", "n_meshes", "n_triangles", "integrator", "seed_mode"]
        return dict(zip(keys, (int(v) for v in out)))

    def camera(self):
        out = np.zeros(68, dtype=np.float32)
        if lib().nori_host_camera(self.h, out.ctypes.data_as(C.c_void_p)):
            raise abi.NoriError(lib().nori_host_last_error().decode())
        return dict(s2c=out[:16].reshape(4, 4).copy(), c2w=out[16:32].reshape(4, 4).copy(), nearClip=float(out[32]),
                    farClip=float(out[33]), filter_radius=float(out[34]), filter_table=out[35:68].copy())

    def mesh(self, i):
        nv, nf = C.c_uint32(), C.c_uint32()
        V, N, UV, F = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        b, e = abi.BsdfDesc(), abi.EmitterDesc()
        if lib().nori_host_mesh(self.h, i, C.byref(nv), C.byref(nf), C.byref(V), C.byref(N), C.byref(UV), C.byref(F), C.byref(b), C.byref(e)):
            raise abi.NoriError(lib().nori_host_last_error().decode())

        def arr(p, n, ty):
            if not p.value:
                return None
            return np.ctypeslib.as_array(C.cast(p, C.POINTER(ty)), shape=(n,)).copy()
        return dict(V=arr(V, nv.value * 3, C.c_float).reshape(-1, 3), F=arr(F, nf.value * 3, C.c_uint32).reshape(-1, 3),
                    N=None if not N.value else arr(N, nv.value * 3, C.c_float).reshape(-1, 3),
                    UV=None if not UV.value else arr(UV, nv.value * 2, C.c_float).reshape(-1, 2),
                    bsdf=dict(type=b.type, albedo=tuple(b.albedo), alpha=b.alpha, intIOR=b.intIOR, extIOR=b.extIOR, ks=b.ks),
                    emitter=dict(type=e.type, radiance=tuple(e.radiance)))

    def render(self, device=0, tile_rank=0, tile_ranks=1, gpus=1, accel_cache=None, srgb8=False):
        """gpus > 1: devices device..device+gpus-1 behind ONE context (nb_create_multi), the path of `nori --gpus N`.
        accel_cache: file caching the built hierarchy (`nori --cache`).  srgb8=True also returns the 8-bit tonemapped image
        twice -- from the device kernel and from the host loop -- as (film, stats, dev8, host8)."""
        i = self.info()
        b = i["border"]
        film = np.zeros((i["height"] + 2 * b, i["width"] + 2 * b, 4), dtype=np.float32)
        st = abi.Stats()
        if gpus > 1 or accel_cache or srgb8:
            d8 = np.zeros((i["height"], i["width"], 3), dtype=np.uint8) if srgb8 else None
            h8 = np.zeros((i["height"], i["width"], 3), dtype=np.uint8) if srgb8 else None
            rc = lib().nori_host_render_gpus(self.h, device, gpus, os.fspath(accel_cache).encode() if accel_cache else None,
                                             film.ctypes.data_as(C.c_void_p), d8.ctypes.data_as(C.c_void_p) if srgb8 else None,
                                             h8.ctypes.data_as(C.c_void_p) if srgb8 else None, C.byref(st))
        else:
            rc = lib().nori_host_render(self.h, device, tile_rank, tile_ranks, film.ctypes.data_as(C.c_void_p), C.byref(st))
        if rc:
            raise abi.NoriError(lib().nori_host_last_error().decode())
        return (film, st, d8, h8) if srgb8 else (film, st)


# ----------------------------------------------------------------------------- XML writer (scene description -> Nori XML + OBJ)
_BSDF_XML = {S.BSDF_DIFFUSE: "diffuse", S.BSDF_MIRROR: "mirror", S.BSDF_DIELECTRIC: "dielectric", S.BSDF_MICROFACET: "microfacet"}
_INT_XML = {v: k for k, v in S.INTEGRATORS.items()}


def _f(x):
    return "%.9g" % float(np.float32(x))


def write_xml(scene: S.Scene, directory: str, name: str = "scene", filter_xml: str | None = None) -> str:
    """Writes <directory>/<name>.xml plus one OBJ per mesh, in the reference's scene grammar (ref: src/parser.cpp:78-102)."""
    os.makedirs(directory, exist_ok=True)
    if scene.integrator == S.INT_SIMPLE:
        integ = ['\t<integrator type="simple">', '\t\t<point name="position" value="' + ", ".join(_f(v) for v in scene.light_pos) + '"/>',
                 '\t\t<color name="energy" value="' + ", ".join(_f(v) for v in scene.light_energy) + '"/>', "\t</integrator>"]
    else:
        integ = [f'\t<integrator type="{_INT_XML[scene.integrator]}"/>']
    lines = ["<?xml version='1.0' encoding='utf-8'?>", "<scene>"] + integ + [
             '\t<sampler type="independent">', f'\t\t<integer name="sampleCount" value="{scene.spp}"/>',
             f'\t\t<string name="seedMode" value="{"block" if scene.seed_mode == S.SEED_PER_BLOCK else "sample"}"/>', "\t</sampler>"]
    cam = scene.camera
    lines += ['\t<camera type="perspective">', f'\t\t<float name="fov" value="{_f(cam.fov)}"/>',
              f'\t\t<float name="nearClip" value="{_f(cam.nearClip)}"/>', f'\t\t<float name="farClip" value="{_f(cam.farClip)}"/>',
              '\t\t<transform name="toWorld">', '\t\t\t<matrix value="' + " ".join(_f(v) for v in np.asarray(cam.c2w, dtype=np.float32).reshape(-1)) + '"/>',
              "\t\t</transform>", f'\t\t<integer name="width" value="{cam.width}"/>', f'\t\t<integer name="height" value="{cam.height}"/>']
    if filter_xml:
        lines.append("\t\t" + filter_xml)
    lines.append("\t</camera>")
    for i, m in enumerate(scene.meshes):
        obj = f"{name}_mesh{i}.obj"
        S.write_obj(os.path.join(directory, obj), m)
        lines += ['\t<mesh type="obj">', f'\t\t<string name="filename" value="{obj}"/>']
        b = m.bsdf
        lines.append(f'\t\t<bsdf type="{_BSDF_XML[b.type]}">')
        if b.type == S.BSDF_DIFFUSE:
            lines.append('\t\t\t<color name="albedo" value="' + " ".join(_f(v) for v in b.albedo) + '"/>')
        elif b.type == S.BSDF_MICROFACET:
            lines += ['\t\t\t<color name="kd" value="' + " ".join(_f(v) for v in b.albedo) + '"/>', f'\t\t\t<float name="alpha" value="{_f(b.alpha)}"/>',
                      f'\t\t\t<float name="intIOR" value="{_f(b.intIOR)}"/>', f'\t\t\t<float name="extIOR" value="{_f(b.extIOR)}"/>']
        elif b.type == S.BSDF_DIELECTRIC:
            lines += [f'\t\t\t<float name="intIOR" value="{_f(b.intIOR)}"/>', f'\t\t\t<float name="extIOR" value="{_f(b.extIOR)}"/>']
        lines.append("\t\t</bsdf>")
        if m.radiance is not None:
            lines += ['\t\t<emitter type="area">', '\t\t\t<color name="radiance" value="' + " ".join(_f(v) for v in m.radiance) + '"/>', "\t\t</emitter>"]
        lines.append("\t</mesh>")
    lines.append("</scene>")
    path = os.path.join(directory, name + ".xml")
    with open(path, "w") as fh:
        fh.write("\n".join(lines) + "\n")
    return path
