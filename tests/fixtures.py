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
