"""The 8-wide compressed hierarchy (nori_b200/csrc/nb_wide.h), checked WITHOUT a GPU: structure of the collapsed tree and a
host walk that uses the kernels' own node step (wide_node_test) against the oracle's brute-force loop -- the reference's
Accel::rayIntersect (ref: src/accel.cpp:23-43) -- bit for bit, closest hit and any hit."""
import numpy as np
import pytest

from nori_b200 import abi
from nori_b200 import scene as S


def random_rays(n, lo, hi, seed=0):
    rng = np.random.default_rng(seed)
    rays = np.zeros(n, dtype=abi.RAY_DTYPE)
    c = 0.5 * (lo + hi); ext = float(np.max(hi - lo))
    o = c + rng.normal(size=(n, 3)) * ext * 1.5
    tgt = lo + rng.random((n, 3)) * (hi - lo)
    d = tgt - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays["o"], rays["d"] = o.astype(np.float32), d.astype(np.float32)
    rays["mint"], rays["maxt"] = 1e-4, np.inf
    rays["d"][:6] = [[1, 0, 0], [0, 1, 0], [0, 0, 1], [-1, 0, 0], [0, -1, 0], [0, 0, -1]]      # zero components (ref: bbox.h:331-333)
    rays["o"][:6] = (c - rays["d"][:6] * ext * 2).astype(np.float32)
    # rays that start INSIDE the geometry's box and leave in all octants
    k = n // 4
    rays["o"][6:6 + k] = (lo + rng.random((k, 3)) * (hi - lo)).astype(np.float32)
    return rays


def scene_arrays(sc):
    V = np.concatenate([m.V for m in sc.meshes]); off = np.cumsum([0] + [m.V.shape[0] for m in sc.meshes[:-1]])
    F = np.concatenate([m.F + o for m, o in zip(sc.meshes, off)]).astype(np.uint32)
    return V, F


def check_structure(nodes, tris, info, nf):
    seen_tri = np.zeros(tris.shape[0], dtype=np.int32)
    seen_node = np.zeros(nodes.shape[0], dtype=np.int32)
    prim = tris[:, 3].copy().view(np.uint32)

    def decode(nd):
        p = nd[:3].copy().view(np.float32)
        e = [(int(nd[3]) >> (8 * a)) & 0xff for a in range(3)]
        imask = int(nd[3]) >> 24
        step = [np.float32(2.0) ** np.float32(x - 127) for x in e]
        meta = [(int(nd[6 + s // 4]) >> (8 * (s % 4))) & 0xff for s in range(8)]
        q = [[(int(nd[8 + 2 * k + s // 4]) >> (8 * (s % 4))) & 0xff for s in range(8)] for k in range(6)]   # lox loy loz hix hiy hiz
        return p, step, imask, meta, q, int(nd[4]), int(nd[5])

    def walk(ni, depth):
        seen_node[ni] += 1
        p, step, imask, meta, q, cbase, tbase = decode(nodes[ni])
        lo = np.full(3, np.inf); hi = np.full(3, -np.inf); maxd = depth
        rel = 0
        for s in range(8):
            m = meta[s]
            if m == 0:
                assert not (imask >> s) & 1
                continue
            blo = np.array([float(p[a]) + q[a][s] * float(step[a]) for a in range(3)]); bhi = np.array([float(p[a]) + q[3 + a][s] * float(step[a]) for a in range(3)])
            if (m & 0x1f) >= 24:                   # inner
                assert (imask >> s) & 1 and (m & 0x1f) == 24 + s and (m >> 5) == 1
                clo, chi, d = walk(cbase + rel, depth + 1); rel += 1
                maxd = max(maxd, d)
            else:
                assert not (imask >> s) & 1
                cnt = {1: 1, 3: 2, 7: 3}[m >> 5]; first = tbase + (m & 0x1f)
                seen_tri[first:first + cnt] += 1
                pts = tris[first:first + cnt].reshape(-1, 4)[:, :3]
                clo, chi = pts.min(0), pts.max(0)
            assert np.all(blo <= clo) and np.all(bhi >= chi), "quantised child box does not contain its subtree"
            lo = np.minimum(lo, clo); hi = np.maximum(hi, chi)
        assert np.all(p.astype(np.float64) <= lo)
        return lo, hi, maxd
    import sys
    sys.setrecursionlimit(10000)
    _, _, d = walk(0, 1)
    assert d == info["depth"] and 2 * d + 2 <= 40
    assert np.all(seen_node == 1) and np.all(seen_tri == 1)
    if tris.shape[0] == nf:                                           # (tiny trees carry one degenerate filler triangle)
        assert sorted(prim.tolist()) == list(range(nf))               # every triangle exactly once


@pytest.mark.parametrize("name", ["bunny", "cbox", "ajax2", "tiny"])
def test_wide_hierarchy_structure_and_hits(oracle, name):
    if name == "bunny":
        sc = S.config_bunny()
    elif name == "cbox":
        sc = S.config_cbox(64, 64, 1)
    elif name == "ajax2":
        sc = S.Scene([S.ajax_standin(2)], S.config_bunny().camera)
    else:
        tri = S.Mesh(np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32), np.array([[0, 1, 2]], np.uint32))
        sc = S.Scene([tri, tri, tri], S.config_bunny().camera)       # coincident duplicates: the tie rule (highest index) must hold
    V, F = scene_arrays(sc)
    nodes, tris, info = abi.debug_build_wide(V, F)
    assert info["tris"] >= F.shape[0] and info["nodes"] >= 1
    if name != "tiny":
        assert info["nodes"] < 0.45 * info["binary_nodes"]           # the collapse really widens the tree
    check_structure(nodes, tris, info, F.shape[0])
    lo, hi = V.min(0), V.max(0)
    rays = random_rays(30000 if name != "tiny" else 2000, lo, hi + (1e-3 if name == "tiny" else 0), seed=4)
    o = oracle.OracleScene(sc)
    got, counts = abi.debug_wide_intersect(nodes, tris, rays)
    ref, _ = o.intersect(rays, accel=0)                              # brute force: the reference's loop
    gp = got[:, 3].copy().view(np.uint32)
    assert np.array_equal(gp, ref["prim"])
    hit = gp != 0xffffffff
    assert np.array_equal(got[hit, 0].tobytes(), ref["t"][hit].tobytes())
    assert np.array_equal(got[hit, 1].tobytes(), ref["u"][hit].tobytes()) and np.array_equal(got[hit, 2].tobytes(), ref["v"][hit].tobytes())
    assert hit.mean() > 0.15
    sh, _ = abi.debug_wide_intersect(nodes, tris, rays, any_hit=True)
    assert np.array_equal(sh[:, 3].copy().view(np.uint32) != 0xffffffff, hit)
    # maxt clipping exactly at the hit distance keeps the hit (t <= maxt, ref: src/mesh.cpp:75); just below it loses it
    r2 = rays[hit][:500].copy(); r2["maxt"] = ref["t"][hit][:500]
    g2, _ = abi.debug_wide_intersect(nodes, tris, r2)
    assert np.array_equal(g2[:, 3].copy().view(np.uint32), ref["prim"][hit][:500])
    if name == "ajax2":
        print("wide: %.1f node visits / ray, %.1f triangle tests / ray" % (counts[0] / rays.shape[0], counts[1] / rays.shape[0]))
