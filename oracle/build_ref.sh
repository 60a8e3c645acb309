#!/usr/bin/env bash
# Builds oracle/_ref/*: the REFERENCE's own sources compiled where they lie under /root/reference.
# TEST INFRASTRUCTURE ONLY.  Outputs go to oracle/_ref/ (git-ignored, travels to the GPU box with the snapshot).
# No reference file is copied or patched: the broken, unused 2-D overload in qr.cuh:55 is skipped by
# pre-defining its include guard (-DQR_CUH), and the wrappers in this directory declare what it would have.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
R="${CLAYMORE_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
mkdir -p "$OUT"
[ -d "$R" ] || { echo "reference checkout not found at $R; keeping prebuilt $OUT"; exit 0; }
INC="-I$R/Library -I$R/Projects/GMPM -I$R/Externals/function_ref -I$R/Externals/optional -I$R/Externals/variant"
CXX=/usr/bin/g++; [ -x "$CXX" ] || CXX=g++
# (1) host build of the reference's 3x3 SVD + constitutive models (svd.cuh, constitutive_models.cuh)
if [ ! -f "$OUT/libclaymore_ref_math.so" ] || [ "$HERE/ref_math_host.cpp" -nt "$OUT/libclaymore_ref_math.so" ]; then
  $CXX -std=c++17 -O2 -fPIC -shared -ffp-contract=off -DQR_CUH $INC -idirafter /usr/local/cuda/include \
      -o "$OUT/libclaymore_ref_math.so" "$HERE/ref_math_host.cpp"
fi
# (2) sm_100a build of the reference's GMPM kernels + a minimal driver (ref_gpu_driver.cu), one .so per DOMAIN_BITS
if [ -f "$HERE/ref_gpu_driver.cu" ]; then
  for BITS in ${CLAYMORE_REF_BITS:-6 7 8 9}; do
    SO="$OUT/libclaymore_ref_gpu_d${BITS}.so"
    if [ ! -f "$SO" ] || [ "$HERE/ref_gpu_driver.cu" -nt "$SO" ] || [ "$HERE/ref_settings.h" -nt "$SO" ]; then
      /usr/local/cuda/bin/nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a --expt-extended-lambda --expt-relaxed-constexpr \
          --use_fast_math -lineinfo -Xcompiler -fPIC -shared -cudart static -DQR_CUH -DSETTINGS_H -DREF_DOMAIN_BITS=$BITS \
          -include "$HERE/ref_settings.h" $INC -o "$SO" "$HERE/ref_gpu_driver.cu"
    fi
  done
fi
echo "oracle/_ref up to date"
