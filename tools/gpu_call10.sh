#!/usr/bin/env bash
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
step() { local name="$1" t="$2"; shift 2; echo "== $name"; local t0=$SECONDS; timeout "$t" "$@" > "gpurun_out/c10_$name.log" 2>&1; echo "   rc=$? ($((SECONDS-t0))s)"; tail -8 "gpurun_out/c10_$name.log" | cut -c1-700; }
step attn 150 python benchmarks/attn_check.py
step bench_own_attn 300 python bench.py --gpus 1 --steps 4 --warmup 3 --no-e2e --attn tcgen05
step ncu_attn 300 ncu --set full --clock-control none --import-source on -k "regex:attn_bwd_kernel|attn_fwd2" -s 2 -c 2 -o gpurun_out/ncu_attn2 -f python benchmarks/ncu_targets.py attn
