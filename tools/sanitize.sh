#!/usr/bin/env bash
# compute-sanitizer passes over the single-GPU kernel tests (run on a B200 box; slow: minutes per tool).
#   tools/sanitize.sh [memcheck|racecheck|synccheck|initcheck] [pytest -k expression]
# memcheck  : out-of-bounds / misaligned global, shared and TMEM-adjacent accesses, leaked allocations
# racecheck : shared-memory hazards inside a CTA (the mbarrier / TMA pipelines are the interesting part)
# synccheck : invalid barrier usage (divergent bar.sync / cluster barriers)
# Cross-GPU flag protocols are outside the sanitizer's view: use vescale_b200.comm.symm_debug (VESCALE_B200_SYMM_DEBUG=1).
set -euo pipefail
TOOL="${1:-memcheck}"
EXPR="${2:-rms_norm or swiglu or cross_entropy or adamw or gemm_nt}"
cd "$(dirname "$0")/.."
exec timeout 3000 compute-sanitizer --tool "$TOOL" --error-exitcode 1 --launch-timeout 120 \
  python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "$EXPR"
