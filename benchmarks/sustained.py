"""Sustained (power-capped) GEMM throughput: back-to-back launches for ~2.5 s per variant, CUDA-event timed,
with nvidia-smi clocks/power sampled during the loop.  Burst numbers (benchmarks/micro.py) run at ~1.9 GHz; a
training step lives under the 1 kW cap at ~1.4-1.5 GHz, where FLOPs per joule decide."""
import json, os, subprocess, sys, threading, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vescale_b200.ops import _ext
_ext.load(required=True); ops = torch.ops.vescale_b200

def sample(stop, out):
    p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "100", "-i", "0"], stdout=subprocess.PIPE, text=True)
    while not stop.is_set():
        l = p.stdout.readline()
        if l:
            try:
                a, b = l.split(","); out.append((float(a), float(b)))
            except Exception: pass
    p.terminate()

def run(name, fn, flops, secs=2.5):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    stop, samples = threading.Event(), []
    t = threading.Thread(target=sample, args=(stop, samples), daemon=True); t.start()
    n = 0; e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    t0 = time.time(); e0.record()
    while time.time() - t0 < secs:
        for _ in range(20): fn()
        n += 20
        torch.cuda.current_stream().synchronize() if n % 200 == 0 else None
    e1.record(); torch.cuda.synchronize(); stop.set(); t.join()
    ms = e0.elapsed_time(e1)
    tail = samples[len(samples)//3:] or samples or [(0, 0)]
    r = {"kernel": name, "tflops_sustained": flops * n / ms / 1e9, "sm_mhz": sorted(x[0] for x in tail)[len(tail)//2], "power_w": sorted(x[1] for x in tail)[len(tail)//2], "launches": n}
    print(json.dumps(r), flush=True); return r

res = []
for (M, N, K) in [(8192, 6144, 4096), (8192, 28672, 4096), (8192, 4096, 14336)]:
    a = torch.randn(M, K, device="cuda").bfloat16(); b = torch.randn(N, K, device="cuda").bfloat16(); c = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    fl = 2.0 * M * N * K
    res.append(run(f"tcgen05_2cta {M}x{N}x{K}", lambda: ops.gemm_nt(a, b, c, False, 2), fl))
    res.append(run(f"cublas {M}x{N}x{K}", lambda: torch.mm(a, b.t(), out=c), fl))
    res.append(run(f"tcgen05_1cta {M}x{N}x{K}", lambda: ops.gemm_nt(a, b, c, False, 1), fl))
os.makedirs("gpurun_out", exist_ok=True); json.dump(res, open("gpurun_out/sustained.json", "w"), indent=1)
