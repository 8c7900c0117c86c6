#!/usr/bin/env bash
# round-2 call 1 (1 GPU): full GPU test suite, never-run kernels, reference arm, our arm with per-kernel profile
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
step() { local name="$1" t="$2"; shift 2; echo "== $name"; local t0=$SECONDS; timeout "$t" "$@" > "gpurun_out/c1_$name.log" 2>&1; echo "   rc=$? ($((SECONDS-t0))s)"; tail -3 "gpurun_out/c1_$name.log" | cut -c1-400; }
step pytest_gpu 900 python -m pytest tests -q -m gpu -p no:cacheprovider
step mxfp8 200 python benchmarks/mxfp8_check.py
step gemm_v3 150 python benchmarks/gemm_variant3_check.py
step ref_n1 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 3
step ours_n1 400 python bench.py --gpus 1 --steps 4 --warmup 3 --profile gpurun_out/c1_profile_n1.txt
