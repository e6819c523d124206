#!/bin/bash
# CPU side of the first GPU call of round 2: build the default library and every experiment variant
# (feature-3dgs_b200/variants/<name>/libf3dgs_b200.so, git-ignored, shipped by gpurun).
set -e
cd "$(dirname "$0")/.."
python feature-3dgs_b200/build.py > /dev/null
( cd tools/micro && for m in red_rate ffma_rate ffma2_rate fma_variants; do [ -f $m.cu ] && nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o $m $m.cu; done )
bash tools/build_variants.sh \
  u0f0  "-DF3DGS_UNIFORM_WARP=0 -DF3DGS_FFMA2=0" \
  u1f0  "-DF3DGS_FFMA2=0" \
  u1f1p "-DF3DGS_FEAT_PREFETCH=1" \
  ec    "-DF3DGS_EXACT_CULL=1" \
  pr    "-DF3DGS_PAIR_SKIP=1" \
  ecpr  "-DF3DGS_EXACT_CULL=1 -DF3DGS_PAIR_SKIP=1" \
  timing "-DF3DGS_TIMING_BUILD=1" \
  nofr  "-DF3DGS_DIAG_NO_FEAT_RED=1" \
  nogr  "-DF3DGS_DIAG_NO_GEOM_RED=1" \
  nofma "-DF3DGS_DIAG_NO_FMA=1" \
  hp    "-DF3DGS_BWD_HELPERS=1" \
  hpec  "-DF3DGS_BWD_HELPERS=1 -DF3DGS_EXACT_CULL=1"
