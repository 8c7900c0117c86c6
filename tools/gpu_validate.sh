#!/usr/bin/env bash
# One-shot GPU validation of everything that was written without hardware access at the end of round 1, in
# increasing order of risk, each step under its own timeout so a hang costs seconds, not the box.
#   tools/gpu_validate.sh            # single GPU
#   tools/gpu_validate.sh 2|4|8      # also the multi-GPU steps on N GPUs
#   NCU=1 tools/gpu_validate.sh      # also ncu --set full captures of the fused kernels (single GPU)
# Results land in gpurun_out/validate_*.log / *.json.
set -uo pipefail
N="${1:-1}"
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
step() { local name="$1" t="$2"; shift 2; echo "== $name"; timeout "$t" "$@" > "gpurun_out/validate_$name.log" 2>&1; echo "   rc=$? (log: gpurun_out/validate_$name.log)"; tail -2 "gpurun_out/validate_$name.log" | cut -c1-300; }

step kernels 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu
step late_kernels 600 python -m pytest tests/test_z_late_gpu.py -q -m gpu
step bench_n1 400 python bench.py --steps 4 --warmup 3
step gemm_v3 150 python benchmarks/gemm_variant3_check.py
step mxfp8 200 python benchmarks/mxfp8_check.py
# ncu captures (single GPU, one capture per kernel family; the world-size-1 test drives the fused / symmetric kernels on one GPU,
# which is the only way to put them under ncu: multi-rank commands must never be wrapped in it).  Read back here with
#   ncu -i gpurun_out/<name>.ncu-rep --page raw --csv | grep -E 'dram__bytes|sm__pipe_tensor|gpu__dram_throughput|lts__t_sectors'
if [ "${NCU:-0}" = "1" ]; then
  NC="ncu --set full --clock-control none --import-source on -c 3"
  step ncu_fused_tp 300 $NC -k regex:fused_tp_kernel -o gpurun_out/fused_tp python -m pytest "tests/test_z_late_gpu.py::test_symmetric_memory_kernels_single_rank" -q -m gpu
  step ncu_mxfp8 300 $NC -k regex:gemm_mxfp8_kernel -s 2 -o gpurun_out/mxfp8 python benchmarks/mxfp8_check.py
  step ncu_vocab_ce 300 $NC -k regex:vocab_ce_kernel -o gpurun_out/vocab_ce python -m pytest "tests/test_z_late_gpu.py::test_symmetric_memory_kernels_single_rank" -q -m gpu
fi
if [ "$N" -gt 1 ]; then
  T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
  step multigpu_tests 900 python -m pytest tests/test_symm_multigpu.py -x -q -m gpu
  step bench_nN 400 $T bench.py --gpus "$N" --steps 4 --warmup 3
  step bench_nccl 400 $T bench.py --gpus "$N" --steps 4 --warmup 3 --no-e2e --comm nccl
  step coll_bench 200 $T benchmarks/coll_bench.py
  step tp_bench 200 $T benchmarks/tp_bench.py
  step moe_bench 200 $T benchmarks/moe_bench.py
  step baseline_fsdp2 500 $T benchmarks/baseline_fsdp2.py --gpus "$N" --steps 4 --warmup 3 --ac full
  if [ "$N" -ge 2 ]; then step bench_tp2 400 $T bench.py --gpus "$N" --steps 4 --warmup 3 --no-e2e --tp 2; fi
  if [ "$N" -ge 8 ]; then
    step mixtral 600 $T benchmarks/mixtral_bench.py --steps 3 --warmup 2
    step llama70b_fp8 900 $T bench.py --gpus 8 --model llama3_70b --fp8 --ac full --fused-reduce --reshard yes --steps 3 --warmup 3 --no-e2e
  fi
fi
