"""The reference's statistical known-answer fixtures, restated as scene descriptions.

Sources (ref = /root/reference): scenes/pa4/tests/test-mesh.xml, test-mesh-furnace.xml,
scenes/pa5/tests/test-direct.xml, test-furnace.xml, ttest-microfacet.xml, chi2test-microfacet.xml.
The reference runs them through src/ttest.cpp / src/chi2test.cpp with the `hypothesis` library
(absent); the same decisions are taken here with scipy.stats.
"""
import numpy as np
from scipy import stats

from nori_b200 import scene as S

POLYLUM_REFS = [0.0898394, 0.02292, 0.0534198, 0.0205314, 0.26174]   # ref: scenes/pa4/tests/test-mesh.xml:4-5
MICROFACET_ANGLES = [0, 45, 60, 80, 85]                              # ref: scenes/pa5/tests/ttest-microfacet.xml:4
MICROFACET_REFS = [0.207067, 0.215733, 0.247884, 0.430936, 0.519016] # ref: scenes/pa5/tests/ttest-microfacet.xml:5
SIGNIFICANCE = 0.01                                                  # ref: src/ttest.cpp:48


def polylum_scene(i: int, integrator: int) -> S.Scene:
    """ref: scenes/pa4/tests/test-mesh.xml:7-38 (camera 0.01 above a rho=0.5 floor, fov 1e-6, 1x1 px)."""
    cam = S.Camera(S.lookat([0, 0.01, 0], [0, 0, 0], [0, 0, 1]).astype(np.float32), 1e-6, 1, 1)
    meshes = [S.with_(S.golden_mesh("test_floor"), S.diffuse((0.5, 0.5, 0.5))),
              S.with_(S.golden_mesh(f"test_polylum{i}"), S.diffuse((0, 0, 0)), radiance=(1, 1, 1))]
    return S.Scene(meshes, cam, integrator, 1, name=f"polylum{i}")


def furnace_scene(albedo: float, integrator: int) -> S.Scene:
    """ref: scenes/pa5/tests/test-furnace.xml:19-37 (identity camera inside an emissive diffuse cube, fov 10)."""
    cam = S.Camera(np.eye(4, dtype=np.float32), 10.0, 1, 1)
    meshes = [S.with_(S.golden_mesh("test_furnace"), S.diffuse((albedo,) * 3), radiance=(1, 1, 1))]
    return S.Scene(meshes, cam, integrator, 1, name=f"furnace{albedo}")


def t_test_pvalue(lum: np.ndarray, reference: float) -> float:
    """Two-sided one-sample Student t-test (hypothesis::students_t_test as called at ref: src/ttest.cpp:126-128)."""
    n = lum.shape[0]
    mean, var = float(lum.mean()), float(lum.var(ddof=1))
    if var == 0.0:
        return 1.0 if abs(mean - reference) < 1e-5 * max(1.0, abs(reference)) else 0.0
    t = abs(mean - reference) * np.sqrt(n) / np.sqrt(var)
    return float(2.0 * stats.t.sf(t, n - 1))


def sidak(alpha: float, ntests: int) -> float:
    return 1.0 - (1.0 - alpha) ** (1.0 / ntests)


def chi2_pvalue(obs: np.ndarray, exp: np.ndarray, min_exp: float = 5.0):
    """hypothesis::chi2_test: pool low-expectation cells, Pearson chi^2 (as called at ref: src/chi2test.cpp:171-173)."""
    order = np.argsort(exp)
    obs, exp = obs[order], exp[order]
    pooled_o = pooled_e = 0.0
    chsq, dof = 0.0, 0
    for o, e in zip(obs, exp):
        if e == 0:
            if o > len(obs) * 1e-5:
                return 0.0
            continue
        if e < min_exp:
            pooled_o += o; pooled_e += e
        else:
            if pooled_e > 0 and pooled_e < min_exp:
                pooled_o += o; pooled_e += e
            else:
                chsq += (o - e) ** 2 / e; dof += 1
    if pooled_e > 0:
        chsq += (pooled_o - pooled_e) ** 2 / pooled_e; dof += 1
    dof -= 1
    if dof <= 0:
        return 1.0
    return float(stats.chi2.sf(chsq, dof))


def integrate_cells(pdf_batch, ct_res: int, phi_res: int, rel_tol: float = 1e-5):
    """Integral of a directional density over every (cos theta, phi) cell of the chi^2 contingency table
    (ref: src/chi2test.cpp:126-153, where hypothesis::adaptiveSimpson2D does it).  Composite Simpson in (theta, phi)
    -- theta, not cos theta: sin(theta) = sqrt(1 - c^2) has an unbounded derivative at c = 1 and stalls Simpson in the
    top row -- with the intervals per axis doubled per cell until the value moves by less than rel_tol (a narrow
    specular lobe at grazing incidence needs 128).  pdf_batch(wo[n,3] float32) -> pdf[n].  Same scheme as
    nori_b200/csrc/host/stat_tests.cpp:integrateCells."""
    res = ct_res * phi_res
    dc, dp = 2.0 / ct_res, 2 * np.pi / phi_res
    value = np.zeros(res); prev = np.full(res, -1.0)
    todo = np.arange(res)
    S = 16
    while todo.size:
        a = np.arange(S + 1)
        w1 = np.where((a == 0) | (a == S), 1.0, np.where(a % 2 == 1, 4.0, 2.0))
        i, j = todo // phi_res, todo % phi_res
        t0 = np.arccos(np.minimum(1.0, -1.0 + (i + 1) * dc)); t1 = np.arccos(np.maximum(-1.0, -1.0 + i * dc))
        th = t0[:, None] + (t1 - t0)[:, None] * a[None, :] / S                      # [cells, S+1]
        ph = (j[:, None] + a[None, :] / S) * dp                                     # [cells, S+1]
        sn = np.sin(th)
        # the density jumps at the horizon, which is a cell boundary: a node ON it must be evaluated from its own cell's
        # side (cos(pi/2) rounds to +6e-17, which would leak the upper-hemisphere value into the cell below)
        c_lo, c_hi = (-1.0 + i * dc)[:, None], (-1.0 + (i + 1) * dc)[:, None]
        cs = np.clip(np.cos(th), c_lo, c_hi)
        cs = np.where(c_lo >= 0, np.maximum(cs, 1e-6), cs)
        sd = np.sqrt(np.maximum(0.0, 1.0 - cs * cs))
        wo = np.stack([sd[:, :, None] * np.cos(ph)[:, None, :], sd[:, :, None] * np.sin(ph)[:, None, :],
                       np.broadcast_to(cs[:, :, None], (todo.size, S + 1, S + 1))], axis=-1).astype(np.float32)
        pdf = pdf_batch(wo.reshape(-1, 3)).reshape(todo.size, S + 1, S + 1).astype(np.float64)
        v = np.einsum("a,b,ca,cab->c", w1, w1, sn, pdf) * ((t1 - t0) / S / 3.0) * (dp / S / 3.0)
        done = (prev[todo] >= 0) & (np.abs(v - prev[todo]) <= rel_tol * np.abs(v) + 1e-10)
        value[todo] = v; prev[todo] = v
        todo = todo[~done] if S < 1024 else todo[:0]
        S *= 2
    return value
