#!/bin/sh
# Development build of the library with phase time stamps (-DFN_TRACE): tools/micro/libfn_trace.so
cd "$(dirname "$0")/../.." || exit 1
HASH=$(python -c "from fitnoise_b200 import build; print(build.source_hash())")
exec nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -diag-suppress 550,1886 \
    -DFN_TRACE -DFN_SINGLE_TU -DFN_SOURCE_HASH="\"$HASH\"" -shared --cudart static \
    -o tools/micro/libfn_trace.so fitnoise_b200/csrc/fn_api.cu
