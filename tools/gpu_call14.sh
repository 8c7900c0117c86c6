#!/bin/bash
# round-2 closing 2-GPU call: multi-GPU test suite at W=2 (split fused-TP tests, fail-fast harness) and the N=2 headline with per-rank diagnostics
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
timeout 200 $T bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/c14_bench_n2.json 2> gpurun_out/c14_bench_n2.err
echo "bench n2 exit $?"; python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/c14_bench_n2.json") if l.startswith("{")][-1])
    print("tok/s", round(d["value"]), "ms", d["ms_per_step"], "steps", d["step_ms"], "host", d.get("host_enqueue_ms"), "e2e", d["e2e"]["value"], "per_rank", d.get("per_rank"), "exposed", d.get("exposed_comm_ms_per_step"))
except Exception as e:
    print("no record:", e); print(open("gpurun_out/c14_bench_n2.err").read()[-1500:])
PY
timeout 420 python -m pytest tests -q -m gpu -p no:cacheprovider --durations=8 > gpurun_out/c14_tests_w2.log 2>&1
echo "pytest w2 exit $?"; tail -22 gpurun_out/c14_tests_w2.log
