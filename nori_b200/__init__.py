"""nori_b200 -- Blackwell-native render hot path for the Nori ray tracer.

    abi       ctypes binding of the C-ABI (include/nori_b200.h, libnori_b200.so) -- the product
    host      ctypes binding of the C++ host mirror of Nori (registry, XML pipeline, plugins, render())
    scene     plain-data scene descriptions + generators of the BASELINE.json workloads
    multigpu  tile sharding / block gather / merge logic for one-process-per-GPU runs
    build     in-tree builds (nvcc sm_100a, g++)

There is no CPU fallback in this package; the CPU oracle lives in oracle/ and is test infrastructure only.
"""
__version__ = "0.1.0"
