#!/usr/bin/env bash
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
step() { local name="$1" t="$2"; shift 2; echo "== $name"; local t0=$SECONDS; timeout "$t" "$@" > "gpurun_out/c3_$name.log" 2>&1; echo "   rc=$? ($((SECONDS-t0))s)"; tail -4 "gpurun_out/c3_$name.log" | cut -c1-600; }
step clc 200 python benchmarks/gemm_clc_check.py
step pytest_gpu 900 python -m pytest tests -q -m gpu -p no:cacheprovider
VESCALE_B200_GEMM_SCHED=clc step bench_clc 300 python bench.py --gpus 1 --steps 4 --warmup 3 --no-e2e
