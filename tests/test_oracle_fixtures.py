"""Pins the CPU oracle to every known-answer fixture the reference holds for the path."""
import ctypes as C

import numpy as np
import pytest

from nori_b200 import scene as S
from tests import fixtures as FX

N_PATHS = 100000   # ref: src/ttest.cpp:64 (sampleCount default)


def test_pcg32_kat(oracle):
    """pcg-random.org demo vector for seed(42, 54) -- the only RNG golden available (ext/pcg32 is un-vendored)."""
    L = oracle.lib()
    r = oracle.Pcg32()
    L.orc_pcg32_seed(C.byref(r), 42, 54)
    got = [L.orc_pcg32_next_uint(C.byref(r)) for _ in range(6)]
    assert got == [0xa15c02b7, 0x7b47f409, 0xba1d3330, 0x83d2f293, 0xbfa4784b, 0xcbed606e]


def test_pcg32_float_and_advance(oracle):
    L = oracle.lib()
    a, b = oracle.Pcg32(), oracle.Pcg32()
    L.orc_pcg32_seed(C.byref(a), 7, 3); L.orc_pcg32_seed(C.byref(b), 7, 3)
    seq = [L.orc_pcg32_next_float(C.byref(a)) for _ in range(100)]
    assert all(0.0 <= x < 1.0 for x in seq)
    L.orc_pcg32_advance(C.byref(b), 57)
    assert L.orc_pcg32_next_float(C.byref(b)) == seq[57]


def test_deterministic_math_accuracy(oracle):
    L = oracle.lib()
    rng = np.random.default_rng(0)
    s, c = C.c_float(), C.c_float()
    for u in rng.random(4000).astype(np.float32):
        L.orc_sincos2pi(float(u), C.byref(s), C.byref(c))
        assert abs(s.value - np.sin(2 * np.pi * float(u))) < 2e-7
        assert abs(c.value - np.cos(2 * np.pi * float(u))) < 2e-7
    for x in (rng.random(2000).astype(np.float32) * 0.999 + 1e-7):
        assert abs(L.orc_logf(float(x)) - np.log(float(x))) <= 2e-7 * max(1.0, abs(np.log(float(x))))
    for x in (-rng.random(2000).astype(np.float32) * 80):
        assert abs(L.orc_expf(float(x)) - np.exp(float(x))) <= 3e-7 * np.exp(float(x))


def test_ttest_microfacet(oracle):
    """ref: scenes/pa5/tests/ttest-microfacet.xml -- one default-seeded stream across the five angles (src/ttest.cpp:93,116)."""
    L = oracle.lib()
    bs = oracle.bsdf_struct(S.microfacet((0.1, 0.2, 0.15), 0.1, 1.5, 1.000277))
    rng = oracle.Pcg32(); L.orc_pcg32_init(C.byref(rng))
    thr = FX.sidak(FX.SIGNIFICANCE, len(FX.MICROFACET_REFS))
    for ang, ref in zip(FX.MICROFACET_ANGLES, FX.MICROFACET_REFS):
        th = np.float32(np.radians(np.float32(ang)))
        wi = np.array([np.sin(th), 0, np.cos(th)], dtype=np.float32)
        w = np.zeros((N_PATHS, 3), np.float32)
        L.orc_bsdf_sample_batch(C.byref(bs), oracle._p(wi), N_PATHS, C.byref(rng), None, oracle._p(w))
        lum = w.astype(np.float64) @ np.array([0.212671, 0.715160, 0.072169])
        assert FX.t_test_pvalue(lum, ref) > thr, (ang, lum.mean(), ref)


CHI2_BSDFS = [  # ref: scenes/pa5/tests/chi2test-microfacet.xml:5-24
    dict(alpha=0.1, intIOR=1.33, extIOR=1.01, kd=(0.0, 0.0, 0.0)),
    dict(alpha=0.3, intIOR=1.5, extIOR=1.01, kd=(0.2, 0.1, 0.6)),
    dict(alpha=0.6, intIOR=1.8, extIOR=1.3, kd=(0.4, 0.2, 0.3)),
]


def test_chi2_microfacet(oracle):
    """ref: scenes/pa5/tests/chi2test-microfacet.xml through src/chi2test.cpp:79-173 -- sample() histogram against the
    integral of pdf() over 10 x 20 (cos theta, phi) cells, 5 incident directions per BSDF, ONE default-seeded pcg32
    consumed in the reference's order.  All 15 cases must be accepted at the Sidak-corrected 1 % level, the
    alpha = 0.1 pure-specular lobe at grazing incidence included (that one needs the converged quadrature of
    fixtures.integrate_cells; a fixed 8 x 8 midpoint rule rejects it)."""
    L = oracle.lib()
    rng = oracle.Pcg32(); L.orc_pcg32_init(C.byref(rng))
    ct_res, phi_res = 10, 20
    n = ct_res * phi_res * 5000
    thr = FX.sidak(0.01, 5 * len(CHI2_BSDFS))
    pvals = []
    for p in CHI2_BSDFS:
        bs = oracle.bsdf_struct(S.microfacet(p["kd"], p["alpha"], p["intIOR"], p["extIOR"]))
        for _ in range(5):
            cos_t = np.float32(L.orc_pcg32_next_float(C.byref(rng)))
            sin_t = np.float32(np.sqrt(max(np.float32(0.0), np.float32(1) - cos_t * cos_t)))
            ph = np.float32(2 * np.pi) * np.float32(L.orc_pcg32_next_float(C.byref(rng)))
            wi = np.array([np.cos(ph) * sin_t, np.sin(ph) * sin_t, cos_t], dtype=np.float32)
            wo = np.zeros((n, 3), np.float32); w = np.zeros((n, 3), np.float32)
            L.orc_bsdf_sample_batch(C.byref(bs), oracle._p(wi), n, C.byref(rng), oracle._p(wo), oracle._p(w))
            ok = ~np.all(w == 0, axis=1)
            ctb = np.clip(np.floor((wo[ok, 2] * 0.5 + 0.5) * ct_res).astype(int), 0, ct_res - 1)
            sp = np.arctan2(wo[ok, 1], wo[ok, 0]) / (2 * np.pi)
            sp[sp < 0] += 1
            pb = np.clip(np.floor(sp * phi_res).astype(int), 0, phi_res - 1)
            obs = np.bincount(ctb * phi_res + pb, minlength=ct_res * phi_res).astype(np.float64)

            def pdf_batch(dirs):
                out = np.zeros((dirs.shape[0], 4), np.float32)
                dirs = np.ascontiguousarray(dirs)
                L.orc_bsdf_eval_pdf_batch(C.byref(bs), oracle._p(wi), oracle._p(dirs), dirs.shape[0], oracle._p(out))
                return out[:, 3]
            exp = FX.integrate_cells(pdf_batch, ct_res, phi_res) * n
            pval = FX.chi2_pvalue(obs, exp)
            pvals.append(pval)
            assert pval > thr, (p, wi, pval)
    assert len(pvals) == 15


@pytest.mark.parametrize("integrator", ["whitted", "path_ems", "path_mats", "path_mis"])
def test_ttest_polygon_light(oracle, integrator):
    """ref: scenes/pa4/tests/test-mesh.xml (whitted) and scenes/pa5/tests/test-direct.xml (path_ems/mats/mis)."""
    thr = FX.sidak(FX.SIGNIFICANCE, 5 if integrator == "whitted" else 15)
    for i, ref in enumerate(FX.POLYLUM_REFS, start=1):
        sc = FX.polylum_scene(i, S.INTEGRATORS[integrator])
        o = oracle.OracleScene(sc)
        lum = o.ttest_lum(N_PATHS, accel=0 if i % 2 else 1)   # alternate brute force / BVH
        assert FX.t_test_pvalue(lum, ref) > thr, (integrator, i, lum.mean(), ref)


@pytest.mark.parametrize("integrator,albedo,ref", [
    ("whitted", 0.5, 1.5), ("whitted", 0.8, 1.8),             # ref: scenes/pa4/tests/test-mesh-furnace.xml:16
    ("path_ems", 0.5, 2.0), ("path_ems", 0.8, 5.0),           # ref: scenes/pa5/tests/test-furnace.xml:17
    ("path_mats", 0.5, 2.0), ("path_mats", 0.8, 5.0),
    ("path_mis", 0.5, 2.0), ("path_mis", 0.8, 5.0)])
def test_ttest_furnace(oracle, integrator, albedo, ref):
    thr = FX.sidak(FX.SIGNIFICANCE, 2 if integrator == "whitted" else 6)
    o = oracle.OracleScene(FX.furnace_scene(albedo, S.INTEGRATORS[integrator]))
    lum = o.ttest_lum(N_PATHS)
    assert FX.t_test_pvalue(lum, ref) > thr, (integrator, albedo, lum.mean(), ref)
