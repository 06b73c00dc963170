#!/bin/bash
# Host-side memory / UB check without a GPU: builds libbartint_cuda.so with AddressSanitizer + UBSan (host code only;
# the device code is compiled as usual) into gpurun_out/_asan/ and runs the CPU tests that drive the exporters, the
# group planner and the ABI argument checks through it (BARTINT_LIB selects the build, libasan is preloaded into python).
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=$ROOT/gpurun_out/_asan
mkdir -p "$OUT"
(cd "$ROOT/bart-int_b200/csrc" && nvcc -O1 -g -std=c++17 -gencode arch=compute_100a,code=sm_100a -DBARTINT_TEST_HOOKS \
    -Xcompiler -fPIC,-fopenmp,-pthread,-fsanitize=address,-fsanitize=undefined,-fno-omit-frame-pointer -shared -lgomp -ldl \
    -o "$OUT/libbartint_cuda.so" bartint_abi.cu forest_host.cu kernels_eval.cu kernels_gather.cu kernels_bitslice.cu \
    kernels_misc.cu group.cu)
cd "$ROOT"
BARTINT_LIB=$OUT/libbartint_cuda.so LD_PRELOAD="$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so)" \
ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 UBSAN_OPTIONS=print_stacktrace=1 \
python -m pytest tests/test_planner_cpu.py tests/test_exporter_cpu.py tests/test_wbart.py tests/test_abi.py \
    tests/test_host_logic.py tests/test_value_splits.py -q -m "not gpu" -k "not digests" -s 2>&1 | tee "$OUT/log.txt" | grep -i "runtime error\|AddressSanitizer\|passed\|failed"
