#!/bin/bash
# round-2 closing 1-GPU call: (1) headline bench on the final tree (event-pool timing path), (2) full 1-GPU test suite,
# (3) attention backward v2 vs v1 vs cuDNN
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/c13_bench_n1.json 2> gpurun_out/c13_bench_n1.err
echo "bench n1 exit $?"; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/c13_bench_n1.json").read().strip().splitlines()[-1])
    print("tok/s", round(d["value"]), "ms", d["ms_per_step"], "steps", d["step_ms"], "host", d.get("host_enqueue_ms"), "e2e", d["e2e"]["value"], "launches", d["gpu_launches"], "clocks", d["clocks"])
except Exception as e:
    print("no record:", e); print(open("gpurun_out/c13_bench_n1.err").read()[-1500:])
PY
timeout 400 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/c13_pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/c13_pytest_gpu.log
for v in 2 1; do
  ATTN_CHECK_OUT=gpurun_out/c13_attn_bwd$v.json VESCALE_B200_ATTN_BWD=$v timeout 150 python benchmarks/attn_check.py --bwd-only > gpurun_out/c13_attn_bwd$v.log 2>&1
  echo "attn bwd variant $v exit $?"; tail -6 gpurun_out/c13_attn_bwd$v.log | cut -c1-330
done
