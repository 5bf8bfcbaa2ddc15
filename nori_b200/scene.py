"""Host-side scene description (plain data) and the generators for the BASELINE.json workloads.

This is plumbing for tests / bench: it produces exactly the arrays the C-ABI consumes
(include/nori_b200.h) -- meshes in the reference's layout (packed xyz positions / normals,
uint32 index triples; ref: include/nori/mesh.h:159-166), the two camera matrices of
PerspectiveCamera (ref: src/perspective.cpp:41-68), the tabulated reconstruction filter
(ref: src/block.cpp:19-27) and POD descriptors for the BSDF / emitter / integrator plugins.
The C++ host (nori_b200/csrc/host) produces the same arrays from Nori XML scenes.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

BSDF_DIFFUSE, BSDF_MIRROR, BSDF_DIELECTRIC, BSDF_MICROFACET = 0, 1, 2, 3
INT_NORMALS, INT_AO, INT_WHITTED, INT_PATH_MATS, INT_PATH_EMS, INT_PATH_MIS, INT_SIMPLE = range(7)
INTEGRATORS = {"normals": INT_NORMALS, "ao": INT_AO, "whitted": INT_WHITTED,
               "path_mats": INT_PATH_MATS, "path_ems": INT_PATH_EMS, "path_mis": INT_PATH_MIS, "simple": INT_SIMPLE}
SEED_PER_SAMPLE, SEED_PER_BLOCK = 0, 1
BLOCK = 32
FILTER_RES = 32

_HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(_HERE)
GOLDEN_MESHES = os.path.join(REPO, "tests", "golden", "ref_meshes.npz")


# ----------------------------------------------------------------------------- descriptors
@dataclass
class Bsdf:
    type: int = BSDF_DIFFUSE
    albedo: tuple = (0.5, 0.5, 0.5)     # ref: src/diffuse.cpp:19 / src/microfacet.cpp:27 (kd)
    alpha: float = 0.1                  # ref: src/microfacet.cpp:18
    intIOR: float = 1.5046              # ref: src/microfacet.cpp:21
    extIOR: float = 1.000277            # ref: src/microfacet.cpp:24

    @property
    def ks(self) -> float:              # ref: src/microfacet.cpp:36
        return float(np.float32(1.0) - np.float32(max(np.float32(a) for a in self.albedo)))


def diffuse(albedo=(0.5, 0.5, 0.5)):
    return Bsdf(BSDF_DIFFUSE, tuple(albedo))


def microfacet(kd=(0.5, 0.5, 0.5), alpha=0.1, intIOR=1.5046, extIOR=1.000277):
    return Bsdf(BSDF_MICROFACET, tuple(kd), alpha, intIOR, extIOR)


def mirror():
    return Bsdf(BSDF_MIRROR)


def dielectric(intIOR=1.5046, extIOR=1.000277):
    return Bsdf(BSDF_DIELECTRIC, (0, 0, 0), 0.1, intIOR, extIOR)


@dataclass
class Mesh:
    V: np.ndarray                       # (nv, 3) float32, world space (toWorld applied at load; ref: src/obj.cpp:52)
    F: np.ndarray                       # (nf, 3) uint32
    N: Optional[np.ndarray] = None      # (nv, 3) float32
    UV: Optional[np.ndarray] = None     # (nv, 2) float32
    bsdf: Bsdf = field(default_factory=Bsdf)
    radiance: Optional[tuple] = None    # area emitter radiance, None = no emitter
    name: str = ""

    def __post_init__(self):
        self.V = np.ascontiguousarray(self.V, dtype=np.float32)
        self.F = np.ascontiguousarray(self.F, dtype=np.uint32)
        if self.N is not None:
            self.N = np.ascontiguousarray(self.N, dtype=np.float32)
        if self.UV is not None:
            self.UV = np.ascontiguousarray(self.UV, dtype=np.float32)


@dataclass
class Camera:
    c2w: np.ndarray                     # (4,4) float32 row-major cameraToWorld
    fov: float = 30.0                   # ref: src/perspective.cpp:32
    width: int = 1280                   # ref: src/perspective.cpp:24-25
    height: int = 720
    nearClip: float = 1e-4              # ref: src/perspective.cpp:35-36
    farClip: float = 1e4

    @property
    def s2c(self) -> np.ndarray:
        return sample_to_camera(self.fov, self.nearClip, self.farClip, self.width, self.height)


@dataclass
class Scene:
    meshes: List[Mesh]
    camera: Camera
    integrator: int = INT_NORMALS
    spp: int = 1
    seed_mode: int = SEED_PER_SAMPLE
    seed: int = 0
    filter_table: Optional[np.ndarray] = None   # (33,) float32
    filter_radius: float = 2.0
    rr_start: int = 3
    max_depth: int = 0
    name: str = ""
    light_pos: Optional[Sequence[float]] = None      # point light of the `simple` integrator (position, energy)
    light_energy: Optional[Sequence[float]] = None

    def __post_init__(self):
        if self.filter_table is None:
            self.filter_table, self.filter_radius = gaussian_table()

    @property
    def border(self) -> int:            # ref: src/block.cpp:20
        return int(math.ceil(np.float32(self.filter_radius) - np.float32(0.5)))

    @property
    def film_shape(self):
        b = self.border
        return (self.camera.height + 2 * b, self.camera.width + 2 * b, 4)

    @property
    def n_samples(self) -> int:
        return self.camera.width * self.camera.height * self.spp

    @property
    def n_tris(self) -> int:
        return int(sum(m.F.shape[0] for m in self.meshes))


# ----------------------------------------------------------------------------- camera / transforms
def sample_to_camera(fov, near, far, W, H) -> np.ndarray:
    """ref: src/perspective.cpp:41-68.  Composed and inverted in float64, rounded once to fp32."""
    aspect = float(np.float32(W) / np.float32(H))
    recip = 1.0 / (float(np.float32(far)) - float(np.float32(near)))
    cot = 1.0 / math.tan(math.radians(float(np.float32(fov)) / 2.0))
    f, n = float(np.float32(far)), float(np.float32(near))
    P = np.array([[cot, 0, 0, 0], [0, cot, 0, 0], [0, 0, f * recip, -n * f * recip], [0, 0, 1, 0]], dtype=np.float64)
    T = np.eye(4); T[0, 3] = -1.0; T[1, 3] = -1.0 / aspect
    S = np.diag([-0.5, -0.5 * aspect, 1.0, 1.0])
    return np.linalg.inv(S @ T @ P).astype(np.float32)


def _norm(v):
    v = np.asarray(v, dtype=np.float64)
    return v / np.linalg.norm(v)


def lookat(origin, target, up) -> np.ndarray:
    """ref: src/parser.cpp:274-289 -- columns [left, newUp, dir, origin]."""
    origin = np.asarray(origin, dtype=np.float64); target = np.asarray(target, dtype=np.float64)
    d = _norm(target - origin)
    left = _norm(np.cross(_norm(up), d))
    new_up = _norm(np.cross(d, left))
    m = np.eye(4)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = left, new_up, d, origin
    return m


def scale(v) -> np.ndarray:
    return np.diag([v[0], v[1], v[2], 1.0])


def translate(v) -> np.ndarray:
    m = np.eye(4); m[:3, 3] = v; return m


# ----------------------------------------------------------------------------- filters (host plugins' eval, tabulated)
def _table(fn, radius):
    r = np.float32(radius)
    tab = np.zeros(FILTER_RES + 1, dtype=np.float32)
    for i in range(FILTER_RES):
        pos = np.float32(r * np.float32(i)) / np.float32(FILTER_RES)
        tab[i] = fn(pos)
    return tab, float(r)


def gaussian_table(radius=2.0, stddev=0.5):
    """ref: src/rfilter.cpp:16-30"""
    r, s = np.float32(radius), np.float32(stddev)
    alpha = np.float32(-1.0) / (np.float32(2.0) * s * s)

    def ev(x):
        return max(np.float32(0.0), np.float32(np.exp(alpha * x * x)) - np.float32(np.exp(alpha * r * r)))
    return _table(ev, radius)


def tent_table():
    """ref: src/rfilter.cpp:79-91"""
    return _table(lambda x: max(np.float32(0.0), np.float32(1.0) - abs(x)), 1.0)


def box_table():
    """ref: src/rfilter.cpp:94-106"""
    return _table(lambda x: np.float32(1.0), 0.5)


def mitchell_table(radius=2.0, B=1.0 / 3.0, C=1.0 / 3.0):
    """ref: src/rfilter.cpp:43-73"""
    B, C, r = np.float32(B), np.float32(C), np.float32(radius)

    def ev(x):
        x = abs(np.float32(2.0) * x / r)
        x2, x3 = x * x, x * x * x
        if x < 1:
            return np.float32(1.0 / 6.0) * ((12 - 9 * B - 6 * C) * x3 + (-18 + 12 * B + 6 * C) * x2 + (6 - 2 * B))
        if x < 2:
            return np.float32(1.0 / 6.0) * ((-B - 6 * C) * x3 + (6 * B + 30 * C) * x2 + (-12 * B - 48 * C) * x + (8 * B + 24 * C))
        return np.float32(0.0)
    return _table(ev, radius)


# ----------------------------------------------------------------------------- OBJ (test plumbing; the product loader is C++)
def load_obj(path, to_world: Optional[np.ndarray] = None) -> Mesh:
    """Wavefront OBJ with the rules of ref: src/obj.cpp:43-112 -- v/vt/vn/f, quads split as
    (0,1,2),(3,0,2), vertices deduplicated on the (p,uv,n) index triple in order of first use."""
    P, T, Nn, idx, verts, vmap = [], [], [], [], [], {}
    M = np.eye(4) if to_world is None else np.asarray(to_world, dtype=np.float64)
    Nm = np.linalg.inv(M)[:3, :3].T
    with open(path) as fh:
        for line in fh:
            tok = line.split()
            if not tok:
                continue
            if tok[0] == "v":
                p = M @ np.array([float(tok[1]), float(tok[2]), float(tok[3]), 1.0])
                P.append(p[:3] / p[3])
            elif tok[0] == "vt":
                T.append([float(tok[1]), float(tok[2])])
            elif tok[0] == "vn":
                n = Nm @ np.array([float(tok[1]), float(tok[2]), float(tok[3])])
                Nn.append(n / np.linalg.norm(n))
            elif tok[0] == "f":
                vs = tok[1:5]
                order = [0, 1, 2] if len(vs) == 3 else [0, 1, 2, 3, 0, 2]
                for k in order:
                    parts = vs[k].split("/")
                    key = (int(parts[0]),
                           int(parts[1]) if len(parts) >= 2 and parts[1] else -1,
                           int(parts[2]) if len(parts) >= 3 and parts[2] else -1)
                    if key not in vmap:
                        vmap[key] = len(verts)
                        verts.append(key)
                    idx.append(vmap[key])
    V = np.array([P[k[0] - 1] for k in verts], dtype=np.float32)
    N = np.array([Nn[k[2] - 1] for k in verts], dtype=np.float32) if Nn else None
    UV = np.array([T[k[1] - 1] for k in verts], dtype=np.float32) if T else None
    F = np.array(idx, dtype=np.uint32).reshape(-1, 3)
    return Mesh(V, F, N, UV, name=os.path.basename(path))


def write_obj(path, mesh: Mesh):
    with open(path, "w") as fh:
        for v in mesh.V:
            fh.write("v %.9g %.9g %.9g\n" % tuple(v))
        if mesh.UV is not None:
            for t in mesh.UV:
                fh.write("vt %.9g %.9g\n" % tuple(t))
        if mesh.N is not None:
            for n in mesh.N:
                fh.write("vn %.9g %.9g %.9g\n" % tuple(n))
        for f in mesh.F:
            a, b, c = (int(i) + 1 for i in f)
            if mesh.N is not None and mesh.UV is not None:
                fh.write(f"f {a}/{a}/{a} {b}/{b}/{b} {c}/{c}/{c}\n")
            elif mesh.N is not None:
                fh.write(f"f {a}//{a} {b}//{b} {c}//{c}\n")
            elif mesh.UV is not None:
                fh.write(f"f {a}/{a} {b}/{b} {c}/{c}\n")
            else:
                fh.write(f"f {a} {b} {c}\n")


# ----------------------------------------------------------------------------- reference-derived fixture meshes
_golden_cache = None


def golden_mesh(name: str) -> Mesh:
    """Meshes extracted from the reference's scenes by tests/golden/make_fixtures.py (bunny, cbox, test meshes)."""
    global _golden_cache
    if _golden_cache is None:
        _golden_cache = np.load(GOLDEN_MESHES)
    z = _golden_cache
    N = z[name + ".N"] if (name + ".N") in z.files else None
    UV = z[name + ".UV"] if (name + ".UV") in z.files else None
    return Mesh(z[name + ".V"], z[name + ".F"], N, UV, name=name)


# ----------------------------------------------------------------------------- generated geometry
def vertex_normals(V, F):
    fn = np.cross(V[F[:, 1]] - V[F[:, 0]], V[F[:, 2]] - V[F[:, 0]]).astype(np.float64)
    N = np.zeros((V.shape[0], 3))
    for k in range(3):
        np.add.at(N, F[:, k], fn)
    ln = np.linalg.norm(N, axis=1, keepdims=True)
    ln[ln == 0] = 1.0
    return (N / ln).astype(np.float32)


def subdivide(V, F):
    """One level of midpoint (1->4) subdivision with shared edge vertices; deterministic."""
    V = np.asarray(V, dtype=np.float64); F = np.asarray(F, dtype=np.int64)
    e = np.concatenate([F[:, [0, 1]], F[:, [1, 2]], F[:, [2, 0]]], axis=0)
    e_sorted = np.sort(e, axis=1)
    uniq, inv = np.unique(e_sorted, axis=0, return_inverse=True)
    inv = inv.reshape(-1)
    mid = 0.5 * (V[uniq[:, 0]] + V[uniq[:, 1]])
    nv = V.shape[0]; nf = F.shape[0]
    m01, m12, m20 = nv + inv[:nf], nv + inv[nf:2 * nf], nv + inv[2 * nf:]
    F2 = np.concatenate([
        np.stack([F[:, 0], m01, m20], 1), np.stack([m01, F[:, 1], m12], 1),
        np.stack([m20, m12, F[:, 2]], 1), np.stack([m01, m12, m20], 1)], axis=0)
    return np.concatenate([V, mid], 0), F2


def ajax_standin(levels: int = 4, seed: int = 7) -> Mesh:
    """Deterministic stand-in for ajax.obj, which the reference does not ship.

    bunny (2 000 tris, ref: scenes/pa1/bunny.obj) x `levels` midpoint subdivisions (4 -> 512 000 tris,
    the Ajax bust has ~544 k) with a fixed-seed normal displacement at every level so the surface
    carries fine-scale relief (occlusion detail), then scaled and placed where the Ajax cameras
    (ref: scenes/pa3/ajax-ao.xml:19-23) look.
    """
    b = golden_mesh("bunny")
    V, F = b.V.astype(np.float64), b.F.astype(np.int64)
    rng = np.random.default_rng(seed)
    edge = np.mean(np.linalg.norm(V[F[:, 1]] - V[F[:, 0]], axis=1))
    for lvl in range(levels):
        n_old = V.shape[0]
        V, F = subdivide(V, F)
        N = vertex_normals(V, F).astype(np.float64)
        amp = 0.12 * edge / (2 ** lvl)
        disp = rng.uniform(-1.0, 1.0, size=V.shape[0]) * amp
        disp[:n_old] *= 0.25
        V = V + N * disp[:, None]
    # place: the Ajax cameras sit at (-65.6, 47.6, 24.4) looking along (0.789, -0.355, -0.501)
    lo, hi = V.min(0), V.max(0)
    c = 0.5 * (lo + hi)
    s = 27.0 / (hi[1] - lo[1])
    V = (V - c) * s + np.array([-10.0, 20.0, -12.0])
    # turn the model to face the camera: rotate about y by ~ -50 degrees
    ang = math.radians(-58.0)
    R = np.array([[math.cos(ang), 0, math.sin(ang)], [0, 1, 0], [-math.sin(ang), 0, math.cos(ang)]])
    ctr = np.array([-10.0, 20.0, -12.0])
    V = (V - ctr) @ R.T + ctr
    Vf = V.astype(np.float32)
    return Mesh(Vf, F.astype(np.uint32), vertex_normals(Vf, F), None, name=f"ajax_standin_l{levels}")


def random_triangles(n: int, s: float = 0.01, seed: int = 1) -> Mesh:
    """BASELINE config 5: centres uniform in [-1,1]^3, edge vectors uniform in [-s,s]^3, no normals."""
    rng = np.random.default_rng(seed)
    c = rng.uniform(-1.0, 1.0, size=(n, 3)).astype(np.float32)
    e1 = rng.uniform(-s, s, size=(n, 3)).astype(np.float32)
    e2 = rng.uniform(-s, s, size=(n, 3)).astype(np.float32)
    V = np.empty((3 * n, 3), dtype=np.float32)
    V[0::3], V[1::3], V[2::3] = c, c + e1, c + e2
    F = np.arange(3 * n, dtype=np.uint32).reshape(n, 3)
    return Mesh(V, F, None, None, name=f"random_{n}")


def quad(p0, p1, p2, p3) -> Mesh:
    V = np.array([p0, p1, p2, p3], dtype=np.float32)
    F = np.array([[0, 1, 2], [0, 2, 3]], dtype=np.uint32)
    return Mesh(V, F, name="quad")


def with_(m: Mesh, bsdf: Optional[Bsdf] = None, radiance=None) -> Mesh:
    return Mesh(m.V, m.F, m.N, m.UV, bsdf if bsdf is not None else m.bsdf, radiance, m.name)


# ----------------------------------------------------------------------------- BASELINE.json workloads
def config_bunny(spp=1, seed_mode=SEED_PER_BLOCK) -> Scene:
    """configs[0]: ref scenes/pa1/bunny.xml verbatim -- 768x768, 1 spp, normals, fov 16."""
    cam = Camera(lookat([-0.0315182, 0.284011, 0.7331], [-0.0123771, 0.0540913, -0.239922],
                        [0.00717446, 0.973206, -0.229822]).astype(np.float32), 16.0, 768, 768)
    return Scene([golden_mesh("bunny")], cam, INT_NORMALS, spp, seed_mode, name="bunny-normals")


_AJAX_CAM = dict(origin=[-65.6055, 47.5762, 24.3583], target=[-64.8161, 47.2211, 23.8576], up=[0.299858, 0.934836, -0.190177])


def empty_mesh() -> Mesh:
    """Placeholder geometry for the ranks of a multi-GPU group that receive the scene arrays from rank 0 over NVLink."""
    return Mesh(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.uint32), name="(replicated from rank 0)")


def config_ajax_ao(width=800, height=600, spp=64, levels=4, geometry=True) -> Scene:
    """configs[1]: Ajax(stand-in) ambient occlusion, camera of ref scenes/pa3/ajax-ao.xml:19-23, fov 30."""
    cam = Camera(lookat(**_AJAX_CAM).astype(np.float32), 30.0, width, height)
    return Scene([ajax_standin(levels) if geometry else empty_mesh()], cam, INT_AO, spp, name=f"ajax-ao-{width}x{height}x{spp}")


def config_cbox(width=512, height=512, spp=256, integrator=INT_PATH_MIS) -> Scene:
    """configs[2]: Cornell box of ref scenes/pa4/cbox/cbox-distributed.xml:6-63, both spheres diffuse, radiance 40."""
    c2w = (lookat([0, 0.919769, 5.41159], [0, 0.893051, 4.41198], [0, 1, 0]) @ scale([-1, 1, 1])).astype(np.float32)
    cam = Camera(c2w, 27.7856, width, height)
    meshes = [
        with_(golden_mesh("cbox_walls"), diffuse((0.725, 0.71, 0.68))),
        with_(golden_mesh("cbox_rightwall"), diffuse((0.161, 0.133, 0.427))),
        with_(golden_mesh("cbox_leftwall"), diffuse((0.630, 0.065, 0.05))),
        with_(golden_mesh("cbox_sphere1"), diffuse()),
        with_(golden_mesh("cbox_sphere2"), diffuse()),
        with_(golden_mesh("cbox_light"), diffuse(), radiance=(40, 40, 40)),
    ]
    return Scene(meshes, cam, integrator, spp, name=f"cbox-{width}x{height}x{spp}")


def config_ajax_microfacet(width=768, height=768, spp=1024, levels=4, integrator=INT_PATH_MIS, geometry=True) -> Scene:
    """configs[3]: ref scenes/pa5/ajax/ajax-rough.xml:14-28 -- microfacet intIOR 1.7, kd .2 .2 .4, alpha .28; light quad radiance 20."""
    cam = Camera(lookat(**_AJAX_CAM).astype(np.float32), 30.0, width, height)
    meshes = [
        with_(ajax_standin(levels) if geometry else empty_mesh(), microfacet((0.2, 0.2, 0.4), 0.28, 1.7)),
        with_(golden_mesh("ajax_light"), diffuse(), radiance=(20, 20, 20)),
    ]
    return Scene(meshes, cam, integrator, spp, name=f"ajax-rough-{width}x{height}x{spp}")


def config_random_tris(n=10_000_000, width=1920, height=1080, spp=1, integrator=INT_AO, s=0.01, geometry=True) -> Scene:
    """configs[4]: synthetic n random triangles, camera at (0,0,4) looking at the origin, fov 40."""
    cam = Camera(lookat([0, 0, 4], [0, 0, 0], [0, 1, 0]).astype(np.float32), 40.0, width, height)
    return Scene([random_triangles(n, s) if geometry else empty_mesh()], cam, integrator, spp, name=f"random{n}-{width}x{height}x{spp}")


def rel_l2(a: np.ndarray, b: np.ndarray) -> float:
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
