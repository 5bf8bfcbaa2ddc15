"""The product's host SAH builder (nori_b200/csrc/nb_bvh.cpp), checked WITHOUT a GPU through nb_debug_build_bvh:
structural invariants of the 64 B node / 48 B triangle layout the kernels consume, and a numpy emulation of the
device walk (same slab test, same leaf decoding, same tie rule) against the oracle's brute-force loop
(the reference's Accel::rayIntersect, ref: src/accel.cpp:23-43)."""
import numpy as np
import pytest

from nori_b200 import abi
from nori_b200 import scene as S

STACK = 64    # nb_kernels.cuh: kStack


def decode(nodes, tris):
    refs = nodes[:, 12:14].copy().view(np.int32)
    return refs


def leaf_range(ref):
    payload = (~int(ref)) & 0xffffffff
    return payload >> 3, (payload & 7) + 1


def child_box(node, c):
    if c == 0:
        return np.array([node[0], node[2], node[8]]), np.array([node[1], node[3], node[9]])
    return np.array([node[4], node[6], node[10]]), np.array([node[5], node[7], node[11]])


def check_tree(V, F, max_leaf, bfs_nodes):
    nodes, tris, info = abi.debug_build_bvh(V, F, max_leaf, bfs_nodes)
    nf = F.shape[0]
    refs = decode(nodes, tris)
    prim_ids = tris[:, 3].copy().view(np.uint32)
    seen_tri = np.zeros(tris.shape[0], dtype=np.int32)
    seen_node = np.zeros(nodes.shape[0], dtype=np.int32)
    max_depth = 0
    # iterative DFS returning the bounds of every subtree
    def bounds_of(ref, depth):
        nonlocal max_depth
        max_depth = max(max_depth, depth)
        if ref < 0:
            first, count = leaf_range(ref)
            assert 1 <= count <= max(max_leaf, 1)
            seen_tri[first:first + count] += 1
            pts = tris[first:first + count].reshape(-1, 4)[:, :3]
            return pts.min(0), pts.max(0)
        seen_node[ref] += 1
        lo = np.full(3, np.inf); hi = np.full(3, -np.inf)
        for c in range(2):
            clo, chi = bounds_of(int(refs[ref, c]), depth + 1)
            blo, bhi = child_box(nodes[ref], c)
            assert np.all(blo <= clo) and np.all(bhi >= chi), "child box does not contain its subtree"
            lo = np.minimum(lo, clo); hi = np.maximum(hi, chi)
        return lo, hi
    import sys
    sys.setrecursionlimit(10000)
    bounds_of(0, 1)
    if bfs_nodes >= 0:
        assert np.all(seen_node == 1), "every inner node is reachable exactly once"
    else:
        # sibling-pair layout: two inner children of one node share an aligned 128-byte line (slots 2k, 2k+1);
        # a slot whose sibling is a leaf stays unused
        assert np.all(seen_node <= 1) and seen_node[0] == 1
        both = (refs[:, 0] >= 0) & (refs[:, 1] >= 0) & (seen_node == 1)
        a, b = refs[both, 0], refs[both, 1]
        assert np.all((np.minimum(a, b) % 2 == 0) & (np.abs(a - b) == 1))
    assert max_depth <= info["depth"] + 1 and info["depth"] < STACK
    if nf >= 2:
        assert np.all(seen_tri == 1), "every leaf triangle is referenced by exactly one leaf"
        real = np.ones(tris.shape[0], dtype=bool)
        if tris.shape[0] == nf + 1:          # a tree with a single leaf: the absent sibling is one all-zero triangle (nb_bvh.h)
            dummy = np.flatnonzero(np.all(tris == 0, axis=1))
            assert dummy.size >= 1 and info["nodes"] == 1
            real[dummy[-1]] = False
        assert real.sum() == nf, "every input triangle sits in exactly one leaf"
        ids = prim_ids[real]
        assert np.array_equal(np.sort(ids), np.arange(nf, dtype=np.uint32))
        # the leaf-ordered triangles are the input triangles, gathered
        assert np.array_equal(tris[real, 0:3], V[F[ids, 0]]) and np.array_equal(tris[real, 4:7], V[F[ids, 1]])
        assert np.array_equal(tris[real, 8:11], V[F[ids, 2]])
    return nodes, tris, info


def soup(n, seed, scale=0.05):
    rng = np.random.default_rng(seed)
    c = rng.uniform(-1, 1, size=(n, 1, 3))
    V = (c + rng.uniform(-scale, scale, size=(n, 3, 3))).reshape(-1, 3).astype(np.float32)
    F = np.arange(3 * n, dtype=np.uint32).reshape(n, 3)
    return V, F


@pytest.mark.parametrize("max_leaf,bfs_nodes", [(1, 0), (3, 2048), (4, 16), (8, -1)])
def test_builder_invariants_bunny(max_leaf, bfs_nodes):
    m = S.golden_mesh("bunny")
    nodes, tris, info = check_tree(m.V, m.F, max_leaf, bfs_nodes)
    assert info["nodes"] >= (m.F.shape[0] + max_leaf - 1) // max_leaf - 1
    if bfs_nodes > 0:
        assert info["top_nodes"] == min(bfs_nodes, info["nodes"])


def test_builder_invariants_soups_and_degenerate_inputs():
    for n, seed in [(2, 0), (3, 1), (17, 2), (1000, 3), (20000, 4)]:
        check_tree(*soup(n, seed), 3, 2048)
    # many identical triangles (no spatial split possible), collinear / zero-area triangles, one huge + many tiny
    V, F = soup(1, 5); check_tree(np.tile(V, (64, 1)), np.arange(192, dtype=np.uint32).reshape(64, 3), 3, 2048)
    V = np.zeros((300, 3), dtype=np.float32); V[:, 0] = np.arange(300)
    check_tree(V, np.arange(300, dtype=np.uint32).reshape(100, 3), 2, 2048)
    V, F = soup(500, 6, 0.001)
    V = np.vstack([V, np.array([[-50, -50, 0], [50, -50, 0], [0, 50, 0]], dtype=np.float32)])
    F = np.vstack([F, np.array([[1500, 1501, 1502]], dtype=np.uint32)])
    check_tree(V, F, 3, 2048)
    # fewer than two triangles: the absent child is a leaf with one all-zero triangle that can never be hit
    for nf in (0, 1):
        V, F = soup(max(nf, 1), 7)
        nodes, tris, info = abi.debug_build_bvh(V, F[:nf], 3, 2048)
        assert info["nodes"] >= 1 and np.all(nodes[:, 12:14].view(np.int32) < 0)
        assert any(np.all(tris[i] == 0) for i in range(tris.shape[0]))


def walk(nodes, refs, tris, o, d, mint=1e-4, maxt=np.inf, any_hit=False):
    """numpy restatement of nb_kernels.cuh:trav_run (order of leaf visits differs; the result does not depend on it)."""
    f32 = np.float32
    o = o.astype(f32); d = d.astype(f32)
    dd = np.where(np.abs(d) > 1e-24, d, np.copysign(f32(1e-24), d)).astype(f32)
    inv = (f32(1) / dd).astype(f32); ood = (o * inv).astype(f32)
    best = (f32(maxt), 0xffffffff, f32(0), f32(0))
    stack = [0]
    while stack:
        ref = stack.pop()
        if ref >= 0:
            n = nodes[ref]
            for c in (0, 1):
                lo, hi = child_box(n, c)
                t0 = (lo.astype(np.float64) * inv - ood).astype(f32); t1 = (hi.astype(np.float64) * inv - ood).astype(f32)   # fma: one rounding
                tmin = max(np.minimum(t0, t1).max(), f32(mint)); tmax = min(np.maximum(t0, t1).min(), best[0])
                if tmin <= tmax:
                    stack.append(int(refs[ref, c]))
                    assert len(stack) < STACK
            continue
        first, count = leaf_range(ref)
        for i in range(first, first + count):
            p0, p1, p2 = tris[i, 0:3], tris[i, 4:7], tris[i, 8:11]
            prim = int(tris[i, 3:4].view(np.uint32)[0])
            e1 = p1 - p0; e2 = p2 - p0
            pv = np.cross(d, e2).astype(f32)
            det = f32(e1[0] * pv[0] + (e1[1] * pv[1] + e1[2] * pv[2]))
            if -1e-8 < det < 1e-8:
                continue
            inv_det = f32(1) / det
            tv = o - p0
            u = f32(f32(tv[0] * pv[0] + (tv[1] * pv[1] + tv[2] * pv[2])) * inv_det)
            if u < 0 or u > 1:
                continue
            q = np.cross(tv, e1).astype(f32)
            v = f32(f32(d[0] * q[0] + (d[1] * q[1] + d[2] * q[2])) * inv_det)
            if v < 0 or u + v > 1:
                continue
            t = f32(f32(e2[0] * q[0] + (e2[1] * q[1] + e2[2] * q[2])) * inv_det)
            if not (t >= mint and t <= best[0]):
                continue
            if any_hit:
                return (t, prim, u, v)
            if best[1] == 0xffffffff or t < best[0] or prim > best[1]:
                best = (t, prim, u, v)
    return best


def test_emulated_device_walk_matches_brute_force(oracle):
    m = S.golden_mesh("bunny")
    sc = S.config_bunny()
    nodes, tris, info = abi.debug_build_bvh(m.V, m.F, 3, 2048)
    refs = decode(nodes, tris)
    rng = np.random.default_rng(21)
    lo, hi = m.V.min(0), m.V.max(0)
    n = 300
    rays = np.zeros(n, dtype=oracle.RAY_DTYPE)
    org = 0.5 * (lo + hi) + rng.normal(size=(n, 3)) * float(np.max(hi - lo)) * 1.5
    dirs = (lo + rng.random((n, 3)) * (hi - lo)) - org
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    rays["o"], rays["d"], rays["mint"], rays["maxt"] = org.astype(np.float32), dirs.astype(np.float32), 1e-4, np.inf
    rays["d"][:3] = [[1, 0, 0], [0, 1, 0], [0, 0, -1]]          # zero components (ref: include/nori/bbox.h:331-333)
    rays["o"][:3] = (0.5 * (lo + hi) - rays["d"][:3] * 2).astype(np.float32)
    hits, _ = oracle.OracleScene(sc).intersect(rays, accel=0)     # the reference's brute-force loop
    n_hit = 0
    for i in range(n):
        t, prim, u, v = walk(nodes, refs, tris, rays["o"][i], rays["d"][i])
        assert prim == int(hits["prim"][i]), i
        if prim != 0xffffffff:
            n_hit += 1
            assert np.float32(t) == hits["t"][i] and np.float32(u) == hits["u"][i] and np.float32(v) == hits["v"][i], i
    assert n_hit > 50


def test_hierarchy_cache_round_trip(tmp_path):
    """SURVEY 8f row 3 / VERDICT r1 item 9: the built BVH (nodes + leaf-ordered triangles) is cached on disk, keyed on the
    geometry and the build parameters; a hit returns the very same arrays, anything else rebuilds."""
    m = S.ajax_standin(2)
    path = tmp_path / "scene.nbbvh"
    n0, t0, i0 = abi.debug_bvh_cache(m.V, m.F, path)              # miss: builds and writes
    assert not i0["hit"] and path.exists() and path.stat().st_size == 64 + n0.nbytes + t0.nbytes
    ref_nodes, ref_tris, ref_info = abi.debug_build_bvh(m.V, m.F)
    assert n0.tobytes() == ref_nodes.tobytes() and t0.tobytes() == ref_tris.tobytes()
    n1, t1, i1 = abi.debug_bvh_cache(m.V, m.F, path)              # hit: same bytes, no build
    assert i1["hit"] and n1.tobytes() == n0.tobytes() and t1.tobytes() == t0.tobytes()
    assert {k: i1[k] for k in ("nodes", "tris", "top_nodes", "depth")} == {k: ref_info[k] for k in ("nodes", "tris", "top_nodes", "depth")}
    # other build parameters or other geometry: the key differs, the file is rebuilt
    _, _, i2 = abi.debug_bvh_cache(m.V, m.F, path, max_leaf=4)
    assert not i2["hit"]
    _, _, i3 = abi.debug_bvh_cache(m.V, m.F, path, max_leaf=4)
    assert i3["hit"]
    V2 = m.V.copy(); V2[17, 1] += 1e-3
    n4, t4, i4 = abi.debug_bvh_cache(V2, m.F, path, max_leaf=4)
    assert not i4["hit"]
    # a damaged file (flipped payload byte, truncation) misses and is replaced
    raw = bytearray(path.read_bytes()); raw[len(raw) // 2] ^= 0x40; path.write_bytes(bytes(raw))
    n5, t5, i5 = abi.debug_bvh_cache(V2, m.F, path, max_leaf=4)
    assert not i5["hit"] and n5.tobytes() == n4.tobytes()
    path.write_bytes(path.read_bytes()[:1000])
    assert not abi.debug_bvh_cache(V2, m.F, path, max_leaf=4)[2]["hit"]
    assert abi.debug_bvh_cache(V2, m.F, path, max_leaf=4)[2]["hit"]
    # a read-only location is not an error: the build just is not cached
    n6, _, i6 = abi.debug_bvh_cache(V2, m.F, "/proc/nope/scene.nbbvh", max_leaf=4)
    assert not i6["hit"] and n6.tobytes() == n4.tobytes()
