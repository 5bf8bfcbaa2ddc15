"""Drives nb_debug_build_bvh of a sanitizer-instrumented build of nb_bvh.cpp (tools/sanitizers/run.sh)."""
import os, sys, ctypes as C, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from nori_b200 import scene as S
L = C.CDLL(__import__('os').environ.get('NB_SAN_BVH', '/tmp/nb_san/libbvh.so'))
vp=C.c_void_p
L.nb_debug_build_bvh.argtypes = [vp, vp, C.c_uint32, C.c_int, C.c_int64, vp, C.c_uint64, vp, C.c_uint64, vp]
def build(V, F, ml, bfs):
    v4 = np.zeros((V.shape[0], 4), np.float32); v4[:, :3] = V
    f4 = np.zeros((F.shape[0], 4), np.uint32); f4[:, :3] = F
    info = np.zeros(4, np.uint32)
    p = lambda a: a.ctypes.data_as(vp)
    assert L.nb_debug_build_bvh(p(v4), p(f4), F.shape[0], ml, bfs, None, 0, None, 0, p(info)) == 0
    nodes = np.zeros((int(info[0]), 16), np.float32); tris = np.zeros((int(info[1]), 12), np.float32)
    assert L.nb_debug_build_bvh(p(v4), p(f4), F.shape[0], ml, bfs, p(nodes), nodes.size, p(tris), tris.size, p(info)) == 0
    return info
m = S.golden_mesh("bunny")
for ml, bfs in [(1, 0), (3, 2048), (4, 16), (8, -1)]:
    print(build(m.V, m.F, ml, bfs))
rng = np.random.default_rng(0)
for n in (0, 1, 2, 3, 17, 1000, 200000):
    c = rng.uniform(-1, 1, size=(max(n,1), 1, 3)); V = (c + rng.uniform(-.05, .05, size=(max(n,1), 3, 3))).reshape(-1, 3).astype(np.float32)
    F = np.arange(3 * max(n,1), dtype=np.uint32).reshape(-1, 3)[:n]
    print(n, build(V, F, 3, 2048), build(V, F, 3, -1))
a = S.ajax_standin(3)
print("ajax", build(a.V, a.F, 3, 2048))
