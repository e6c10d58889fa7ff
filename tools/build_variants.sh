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
build ch0 "-DHS_PHASE_CH=0"         # quadrature states through the full phase_state (the round-1 default)
