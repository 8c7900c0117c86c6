#!/usr/bin/env bash
# ncu --set full captures (1 GPU): attention fwd/bwd; FSDP comm kernels + fused TP + vocab CE at world size 1
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NC="ncu --set full --clock-control none --import-source on"
step() { local name="$1" t="$2"; shift 2; echo "== $name"; local t0=$SECONDS; timeout "$t" "$@" > "gpurun_out/c8_$name.log" 2>&1; echo "   rc=$? ($((SECONDS-t0))s)"; tail -3 "gpurun_out/c8_$name.log" | cut -c1-300; }
step ncu_attn 400 $NC -k regex:attn_ -s 3 -c 3 -o gpurun_out/ncu_attn -f python benchmarks/ncu_targets.py attn
step ncu_comm 400 $NC -k "regex:reduce_scatter|all_gather_pull|fused_tp_kernel|vocab_ce|handshake" -c 10 -o gpurun_out/ncu_comm -f python -m pytest "tests/test_z_late_gpu.py::test_symmetric_memory_kernels_single_rank" -q -m gpu -p no:cacheprovider
ls -la gpurun_out/*.ncu-rep
