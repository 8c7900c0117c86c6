#!/usr/bin/env bash
# round-2 call 2 (2 GPUs): where does the N>1 communication tax come from?  AG transport (copy engine vs SM pull), RS grid size,
# reshard on/off, NCCL arm, reference arm; W=2 multi-GPU tests.
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=2
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
step() { local name="$1" t="$2"; shift 2; echo "== $name"; local t0=$SECONDS; timeout "$t" "$@" > "gpurun_out/c2_$name.log" 2>&1; echo "   rc=$? ($((SECONDS-t0))s)"; grep -h '"metric"' "gpurun_out/c2_$name.log" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('   ', d.get('impl'), 'tok/s', round(d['value']), 'ms', round(d['ms_per_step'],1), 'exposed', d.get('exposed_comm_ms_per_step'), 'steps', d.get('step_ms'))
" 2>/dev/null || tail -3 "gpurun_out/c2_$name.log" | cut -c1-300; }
B="bench.py --gpus $N --steps 4 --warmup 3 --no-e2e"
step ce_default 300 $T $B --profile gpurun_out/c2_profile_ce.txt
VESCALE_B200_AG_IMPL=pull step pull 300 $T $B --profile gpurun_out/c2_profile_pull.txt
VESCALE_B200_RS_CTAS=16 step ce_rs16 300 $T $B
VESCALE_B200_RS_CTAS=32 step ce_rs32 300 $T $B
VESCALE_B200_MULTIMEM=0 VESCALE_B200_RS_CTAS=32 step ce_rs32_p2p 300 $T $B
step ce_reshard 300 $T $B --reshard yes
step nccl 300 $T $B --comm nccl
step ref 400 $T bench.py --impl reference --gpus $N --steps 3 --warmup 3
step tests_w2 600 python -m pytest tests/test_symm_multigpu.py -q -m gpu -p no:cacheprovider
