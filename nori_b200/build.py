"""Builds the native libraries in-tree (they travel to the GPU box with the snapshot; *.so is git-ignored).

  nori_b200/lib/libnori_b200.so  -- CUDA C-ABI (sm_100a only), nvcc
  nori_b200/lib/libnori_host.so  -- C++ host mirror of Nori's object/parser/plugin layer, g++
  nori_b200/lib/nori             -- CLI (ref: src/main.cpp), links both
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib")
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",                       # parity: every fp32 mul/add rounds separately
    "-Xcompiler", "-fPIC,-ffp-contract=off,-O3",
    "-Xptxas", "-v",
]


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd, log_name):
    r = subprocess.run(cmd, capture_output=True, text=True)
    with open(os.path.join(LIB, log_name), "w") as fh:
        fh.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError(f"build failed: {' '.join(cmd)}")
    return r


def build_cuda(force=False, variant="", extra_flags=()):
    """nvcc, one object per translation unit (compiled concurrently), linked into libnori_b200<variant>.so."""
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(LIB, exist_ok=True)
    target = os.path.join(LIB, f"libnori_b200{variant}.so")
    srcs = [os.path.join(CSRC, f) for f in ("nb_api.cu", "nb_aux.cu", "nb_wave.cu", "nb_bvh.cpp", "nb_wide.cpp")]
    deps = srcs + [os.path.join(CSRC, f) for f in ("nb_bvh.h", "nb_kernels.cuh", "nb_device.cuh", "nb_lbvh.cuh", "nb_ctx.h", "nb_multi.inl", "nb_wave.cuh", "nb_wide.h")
                   if os.path.exists(os.path.join(CSRC, f))] + [os.path.join(os.path.dirname(HERE), "include", "nori_b200.h")]
    if force or _stale(target, deps):
        objdir = os.path.join(LIB, f"obj{variant}")
        os.makedirs(objdir, exist_ok=True)
        objs = [os.path.join(objdir, os.path.splitext(os.path.basename(s))[0] + ".o") for s in srcs]

        def cc(pair):
            src, obj = pair
            return _run([NVCC] + NVCC_FLAGS + list(extra_flags) + ["-c", "-o", obj, src], f"build_cuda{variant}_{os.path.basename(src)}.log")
        with ThreadPoolExecutor(len(srcs)) as ex:
            logs = list(ex.map(cc, zip(srcs, objs)))
        with open(os.path.join(LIB, f"build_cuda{variant}.log"), "w") as fh:      # one combined ptxas -v log, as before
            for r in logs:
                fh.write(r.stdout + r.stderr)
        _run([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", target] + objs + ["-ldl"], f"link_cuda{variant}.log")
    return target


def build_host(force=False):
    host = os.path.join(CSRC, "host")
    if not os.path.isdir(host):
        return None
    os.makedirs(LIB, exist_ok=True)
    target = os.path.join(LIB, "libnori_host.so")
    srcs = sorted(os.path.join(host, f) for f in os.listdir(host) if f.endswith(".cpp") and f != "main.cpp")
    hdrs = [os.path.join(dp, f) for dp, _, fs in os.walk(host) for f in fs if f.endswith(".h")]
    inc = ["-I", host, "-I", os.path.join(os.path.dirname(HERE), "include")]
    flags = ["-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-Wall", "-pthread"]
    if force or _stale(target, srcs + hdrs):
        _run(["g++"] + flags + inc + ["-shared", "-o", target] + srcs + ["-L", LIB, "-lnori_b200", "-ldl", "-Wl,-rpath,$ORIGIN"], "build_host.log")
    exe = os.path.join(LIB, "nori")
    main = os.path.join(host, "main.cpp")
    if os.path.exists(main) and (force or _stale(exe, [main, target] + hdrs)):
        _run(["g++"] + flags + inc + ["-o", exe, main, "-L", LIB, "-lnori_host", "-lnori_b200", "-ldl", "-Wl,-rpath,$ORIGIN"], "build_cli.log")
    return target


def build_all(force=False):
    return build_cuda(force), build_host(force)


if __name__ == "__main__":
    print(build_all("--force" in sys.argv))
