#!/usr/bin/env bash
# round-2 call 4 (2 GPUs): RS unroll + grid sweep, CLC GEMM scheduling in the real step, TP2 (config 3 shape) fused vs plain, B0 FSDP2, W=2 tests
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=2
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
step() { local name="$1" t="$2"; shift 2; echo "== $name"; local t0=$SECONDS; timeout "$t" "$@" > "gpurun_out/c4_$name.log" 2>&1; echo "   rc=$? ($((SECONDS-t0))s)"; grep -h '"metric"' "gpurun_out/c4_$name.log" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('   ', d.get('impl'), 'tok/s', round(d['value']), 'ms', round(d['ms_per_step'],1), 'exposed', d.get('exposed_comm_ms_per_step'), 'steps', d.get('step_ms'))
" 2>/dev/null || tail -3 "gpurun_out/c4_$name.log" | cut -c1-300; }
B="bench.py --gpus $N --steps 4 --warmup 3 --no-e2e"
step default 300 $T $B --profile gpurun_out/c4_profile_default.txt
VESCALE_B200_GEMM_SCHED=clc step clc 300 $T $B --profile gpurun_out/c4_profile_clc.txt
VESCALE_B200_RS_CTAS=16 step rs16 300 $T $B
VESCALE_B200_RS_CTAS=64 step rs64 300 $T $B
VESCALE_B200_MULTIMEM=0 step p2p 300 $T $B
step tp2_fused 300 $T $B --tp 2
step tp2_plain 300 $T $B --tp 2 --tp-impl plain
step fsdp2_b0 400 $T benchmarks/baseline_fsdp2.py --gpus $N --steps 4 --warmup 3
step tests_w2 600 python -m pytest tests/test_symm_multigpu.py -q -m gpu -p no:cacheprovider
