#!/bin/bash
# Host-side sanitizer passes (no GPU): ASan+UBSan over the C++ host mirror and the SAH builder, TSan over the builder's
# threaded top levels.  Builds instrumented copies under /tmp/nb_san; the product libraries are not touched.
set -e
R=$(cd "$(dirname "$0")/../.." && pwd); O=/tmp/nb_san; mkdir -p $O
ASAN=$(g++ -print-file-name=libasan.so); TSAN=$(g++ -print-file-name=libtsan.so); STD=$(g++ -print-file-name=libstdc++.so.6)
python -c "from nori_b200 import build; build.build_cuda()"    # libnori_host links against libnori_b200.so
g++ -O1 -g -std=c++17 -fPIC -ffp-contract=off -fsanitize=address,undefined -fno-omit-frame-pointer -pthread -I $R/nori_b200/csrc/host -I $R/include \
    -shared -o $O/libnori_host.so $(ls $R/nori_b200/csrc/host/*.cpp | grep -v main.cpp) -L $R/nori_b200/lib -lnori_b200 -ldl -Wl,-rpath,$R/nori_b200/lib
echo "== ASan+UBSan: host mirror"
LD_PRELOAD="$ASAN $STD" ASAN_OPTIONS=detect_leaks=0 UBSAN_OPTIONS=print_stacktrace=1 python $R/tools/sanitizers/host_driver.py 2>&1 | grep -E "loaded|reference scenes|fuzz|ERROR|runtime error|Sanitizer" 
g++ -O1 -g -std=c++17 -fPIC -fsanitize=address,undefined -fno-omit-frame-pointer -pthread -shared -o $O/libbvh.so $R/nori_b200/csrc/nb_bvh.cpp
echo "== ASan+UBSan: SAH builder"
LD_PRELOAD="$ASAN $STD" ASAN_OPTIONS=detect_leaks=0 python $R/tools/sanitizers/bvh_driver.py 2>&1 | tail -4
g++ -O1 -g -std=c++17 -fPIC -fsanitize=thread -fno-omit-frame-pointer -pthread -shared -o $O/libbvh.so $R/nori_b200/csrc/nb_bvh.cpp
echo "== TSan: SAH builder (threaded top levels, 200 k triangles)"
LD_PRELOAD="$TSAN $STD" python $R/tools/sanitizers/bvh_driver.py 2>&1 | grep -E "WARNING|data race|^200000|^ajax"
echo "== done"
