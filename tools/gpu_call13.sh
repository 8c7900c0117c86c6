#!/bin/bash
# attention backward v2 (drain warpgroup + pipelined MMA order): numerics + timing against v1 and cuDNN, then the flagship bench with it
mkdir -p gpurun_out
for v in 2 1; do
  VESCALE_B200_ATTN_BWD=$v timeout 240 python benchmarks/attn_check.py --bwd > gpurun_out/c13_attn_bwd$v.log 2>&1
  echo "attn_check bwd variant $v exit $?"
  tail -12 gpurun_out/c13_attn_bwd$v.log
done
timeout 300 python -m pytest tests -m gpu -x -q -k "attention" 2>&1 | tail -3
timeout 400 python bench.py --steps 8 --warmup 4 --attn tcgen05 > gpurun_out/c13_bench_attn_own.json 2> gpurun_out/c13_bench_attn_own.err
echo "bench own-attn exit $?"; cat gpurun_out/c13_bench_attn_own.json | cut -c1-400
