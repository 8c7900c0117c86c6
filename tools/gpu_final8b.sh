#!/usr/bin/env bash
# round-2 final, part B (8 GPUs): B1 (--comm nccl), B0 (torch FSDP2), config 3 (tp2 x fsdp4, fused and plain), config 4 (Mixtral EP), config 5 (70B fp8)
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=8
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
step() { local name="$1" t="$2"; shift 2; echo "== $name"; local t0=$SECONDS; timeout "$t" "$@" > "gpurun_out/f8_$name.log" 2>&1; echo "   rc=$? ($((SECONDS-t0))s)"; grep -h '"metric"' "gpurun_out/f8_$name.log" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('   ', d.get('impl'), d.get('metric','')[:40], 'tok/s', round(d['value']), 'ms', round(d['ms_per_step'],1), 'steps', d.get('step_ms'))
" 2>/dev/null || tail -3 "gpurun_out/f8_$name.log" | cut -c1-300; }
B="bench.py --gpus $N --steps 4 --warmup 3 --no-e2e"
step nccl 300 $T $B --comm nccl
step fsdp2_b0 300 $T benchmarks/baseline_fsdp2.py --gpus $N --steps 4 --warmup 3
step tp2_fused 300 $T $B --tp 2
step tp2_plain 300 $T $B --tp 2 --tp-impl plain
step mixtral_symm 400 $T benchmarks/mixtral_bench.py --steps 3 --warmup 2 --dispatch symm
step mixtral_nccl 400 $T benchmarks/mixtral_bench.py --steps 3 --warmup 2 --dispatch nccl
step llama70b_fp8 600 $T bench.py --gpus 8 --model llama3_70b --fp8 --ac full --fused-reduce --reshard yes --steps 2 --warmup 2 --no-e2e
