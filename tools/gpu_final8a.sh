#!/usr/bin/env bash
# round-2 final, part A (8 GPUs): headline (ours, with per-kernel profile + NVLink byte counters), reference arm, W=8 multi-GPU tests
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=8
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
step() { local name="$1" t="$2"; shift 2; echo "== $name"; local t0=$SECONDS; timeout "$t" "$@" > "gpurun_out/f8_$name.log" 2>&1; echo "   rc=$? ($((SECONDS-t0))s)"; grep -h '"metric"' "gpurun_out/f8_$name.log" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('   ', d.get('impl'), 'tok/s', round(d['value']), 'ms', round(d['ms_per_step'],1), 'e2e', round(d.get('e2e',{}).get('value',0)), 'exposed', d.get('exposed_comm_ms_per_step'), 'steps', d.get('step_ms'))
" 2>/dev/null || tail -3 "gpurun_out/f8_$name.log" | cut -c1-300; }
nvidia-smi nvlink -gt d -i 0 > gpurun_out/f8_nvlink_before.txt 2>&1
step ours 400 $T bench.py --gpus $N --steps 5 --warmup 3 --profile gpurun_out/f8_profile_n8.txt
nvidia-smi nvlink -gt d -i 0 > gpurun_out/f8_nvlink_after.txt 2>&1
step ref 500 $T bench.py --impl reference --gpus $N --steps 3 --warmup 3
step tests_w8 600 python -m pytest tests/test_symm_multigpu.py -q -m gpu -p no:cacheprovider --durations=0
