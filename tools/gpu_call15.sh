#!/bin/bash
# N=2 headline after moving the pre-region housekeeping in front of the last warm-up step (no idle gap before the timed steps)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542"
timeout 170 $T bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/c15_bench_n2.json 2> gpurun_out/c15_bench_n2.err
echo "bench n2 exit $?"; python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/c15_bench_n2.json") if l.startswith("{")][-1])
    print("tok/s", round(d["value"]), "ms", d["ms_per_step"], "steps", d["step_ms"], "host", d.get("host_enqueue_ms"), "e2e", d["e2e"]["value"], "per_rank", d.get("per_rank"))
except Exception as e:
    print("no record:", e); print(open("gpurun_out/c15_bench_n2.err").read()[-1500:])
PY
