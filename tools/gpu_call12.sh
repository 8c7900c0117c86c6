#!/usr/bin/env bash
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
step() { local name="$1" t="$2"; shift 2; echo "== $name"; local t0=$SECONDS; timeout "$t" "$@" > "gpurun_out/c12_$name.log" 2>&1; echo "   rc=$? ($((SECONDS-t0))s)"; tail -6 "gpurun_out/c12_$name.log" | cut -c1-800; }
step attn 150 python benchmarks/attn_check.py
step pytest_attn 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "attention" -p no:cacheprovider
step bench_own_attn 300 python bench.py --gpus 1 --steps 4 --warmup 3 --no-e2e --attn tcgen05
