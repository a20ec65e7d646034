#!/bin/bash
# Kernel experiments: build a variant of libmptg.so with extra nvcc flags into mpt_b200/_lib/variants/<name>/
# (next to the product library, which is left alone).  Use it with MPTG_LIB=<path> python tools/....
#   tools/build_variant.sh probe -DMPTG_KNN_PROBE
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
out=$root/mpt_b200/_lib/variants/$name
mkdir -p "$out"
flags="-std=c++17 -O3 -lineinfo --fmad=false -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-ffp-contract=off,-O2,-fopenmp -ccbin /usr/bin/g++ --expt-relaxed-constexpr --extended-lambda -Xptxas -warn-spills"
pids=()
for src in "$root"/mpt_b200/csrc/*.cu; do
  /usr/local/cuda/bin/nvcc $flags "$@" -c "$src" -o "$out/$(basename "${src%.cu}").o" & pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
/usr/local/cuda/bin/nvcc -shared -o "$out/libmptg.so" "$out"/*.o -ccbin /usr/bin/g++ -lcudart_static -lpthread -ldl -lrt -lgomp
echo "$out/libmptg.so"
