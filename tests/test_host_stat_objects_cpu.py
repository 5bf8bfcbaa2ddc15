"""CPU-only end-to-end runs of the host's `ttest` / `chi2test` scene objects (nori_b200/csrc/host/stat_tests.cpp) against
a TEST DOUBLE of the C-ABI (tests/mock_device/mock_nb.c: closed-form diffuse BSDF, uniform `Li` luminances).  The double is
compiled into a temporary directory and loaded in a child process; the product libraries are not involved beyond a copy
of libnori_host.so.  On the GPU box the same objects run on the real device (tests/test_zz_gpu_*.py)."""
import os
import shutil
import subprocess
import sys
import textwrap

import pytest

from nori_b200 import host

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)

DRIVER = textwrap.dedent("""
    import ctypes as C, sys
    C.CDLL(sys.argv[1] + "/libnori_b200.so", mode=C.RTLD_GLOBAL)
    L = C.CDLL(sys.argv[1] + "/libnori_host.so")
    L.nori_host_load.restype = C.c_void_p; L.nori_host_load.argtypes = [C.c_char_p]
    L.nori_host_last_error.restype = C.c_char_p
    h = L.nori_host_load(sys.argv[2].encode())
    sys.stdout.flush()
    print("RESULT:", "loaded" if h else L.nori_host_last_error().decode())
""")


@pytest.fixture(scope="module")
def mock_dir(tmp_path_factory):
    d = tmp_path_factory.mktemp("mockdev")
    host.lib()                                       # makes sure libnori_host.so is built
    shutil.copy(host.LIB_PATH, d / "libnori_host.so")
    subprocess.run(["gcc", "-O1", "-shared", "-fPIC", "-I", os.path.join(REPO, "include"), "-o", str(d / "libnori_b200.so"),
                    os.path.join(HERE, "mock_device", "mock_nb.c"), "-lm"], check=True)
    (d / "tri.obj").write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3\n")
    return d


def run(mock_dir, xml_text, env=None):
    p = mock_dir / "case.xml"
    p.write_text(xml_text)
    e = dict(os.environ); e.update(env or {})
    r = subprocess.run([sys.executable, "-c", DRIVER, str(mock_dir), str(p)], capture_output=True, text=True, env=e, timeout=600)
    assert r.returncode == 0, r.stderr
    return r.stdout


SCENE = '<scene><integrator type="path_mis"/><mesh type="obj"><string name="filename" value="tri.obj"/></mesh><camera type="perspective"/></scene>'


def test_ttest_scene_mode_accepts_and_rejects(mock_dir):
    ok = run(mock_dir, '<test type="ttest"><string name="references" value="0.5, 0.5"/>' + SCENE + SCENE + '</test>')
    assert "Passed 2/2 tests." in ok and "is not a scene" in ok            # ran to the end; the root is a test object
    assert ok.count("accepted the null hypothesis") == 2
    bad = run(mock_dir, '<test type="ttest"><string name="references" value="0.5, 0.51"/>' + SCENE + SCENE + '</test>')
    assert "Passed 1/2 tests." in bad and "Some tests failed" in bad and "REJECTED" in bad
    shifted = run(mock_dir, '<test type="ttest"><string name="references" value="0.5"/><integer name="sampleCount" value="20000"/>' + SCENE + '</test>',
                  {"MOCK_NB_LI_SHIFT": "0.02"})
    assert "Passed 0/1 tests." in shifted


def test_ttest_bsdf_mode(mock_dir):
    # diffuse: sample() returns the albedo for every sample -> mean = luminance(albedo) exactly, variance 0
    lum = 0.2 * 0.212671 + 0.5 * 0.715160 + 0.7 * 0.072169
    xml = ('<test type="ttest"><string name="angles" value="0, 30, 80"/><string name="references" value="%.7f, %.7f, %.7f"/>'
           '<bsdf type="diffuse"><color name="albedo" value="0.2, 0.5, 0.7"/></bsdf></test>')
    ok = run(mock_dir, xml % (lum, lum, lum))
    assert "Passed 3/3 tests." in ok and "Testing (angle=30)" in ok
    bad = run(mock_dir, xml % (lum, lum * 1.01, lum))
    assert "Passed 2/3 tests." in bad and "Some tests failed" in bad


def test_chi2test_accepts_matching_and_rejects_skewed_sampling(mock_dir):
    xml = ('<test type="chi2test"><integer name="testCount" value="3"/>'
           '<bsdf type="diffuse"><color name="albedo" value="0.5, 0.5, 0.5"/></bsdf></test>')
    ok = run(mock_dir, xml)
    assert "Passed 3/3 tests." in ok and "10x20 contingency table" in ok
    skew = run(mock_dir, xml, {"MOCK_NB_PDF_SCALE": "1.1"})              # samples ~ cos^1.1, pdf says cos
    assert "Passed 0/3 tests." in skew and "Some tests failed" in skew
    small = run(mock_dir, '<test type="chi2test"><integer name="resolution" value="4"/><integer name="sampleCount" value="40000"/>'
                          '<integer name="testCount" value="2"/><bsdf type="diffuse"/></test>')
    assert "Passed 2/2 tests." in small and "4x8 contingency table" in small


def test_device_errors_surface_as_nori_exceptions(mock_dir):
    out = run(mock_dir, '<test type="chi2test"><bsdf type="mirror"/></test>')
    assert "nb_bsdf_sample: mock device: diffuse only" in out
