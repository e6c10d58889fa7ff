#!/bin/bash
# development aid: tuning builds of the library into build/variants/ (git-ignored; they travel to the GPU box with gpurun).
#   tools/build_variants.sh            -> base + every variant below
# then on the box:  tools/tune_variants.sh   (times each build)
#                   HYPERELASTIC_B200_LIB=$PWD/build/variants/ch.so python -m pytest tests -m gpu -q   (parity of a variant)
set -e
cd "$(dirname "$0")/.."
mkdir -p build/variants
FLAGS="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --fmad=true"
build() { echo "== $1: $2"; nvcc $FLAGS $2 -shared -o build/variants/$1.so hyperelasticsolver_b200/csrc/hs_api.cu -lcudart; }
build base ""
build ch "-DHS_PHASE_CH=1"          # quadrature states through phase_state_row1 (B = A A^T + Cayley-Hamilton): -11 FP64 instr / state
build crow0 "-DHS_SP_CROW=0"        # 6-row single-phase cell cache (lo, hi instead of c_max)
