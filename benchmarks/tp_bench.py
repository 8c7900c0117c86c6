"""Fused TP kernels vs NCCL+cuBLAS on the Llama-3-8B tensor-parallel shapes (run under torchrun, >= 2 GPUs).
Times on the device with CUDA events, max over ranks; roofline = max(FLOPs / measured cuBLAS sustained,
NVLink bytes / 770 GB/s measured peer-copy)."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timeit(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    t = torch.tensor([ts[len(ts) // 2]], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def main():
    dist.init_process_group("nccl")
    rank, W = dist.get_rank(), dist.get_world_size()
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    dev = torch.device("cuda", torch.cuda.current_device())
    from vescale_b200 import init_device_mesh
    from vescale_b200.comm.fused_tp import FusedTP

    mesh = init_device_mesh("cuda", (W,), mesh_dim_names=("TP",))
    tp = FusedTP(mesh, "TP", dev, rs_impl="staged")
    tp_nvls = FusedTP(mesh, "TP", dev, rs_impl="nvls") if os.environ.get("TP_BENCH_NVLS", "1") == "1" else None
    peak, link = 1462.2e12, 770e9
    T, H, F, QKV = 8192, 4096, 14336, 6144
    res = []
    for name, (Ml, K, Nr) in {"qkv (ag_gemm)": (T // W, H, QKV // W), "gate_up (ag_gemm)": (T // W, H, 2 * F // W)}.items():
        x = torch.randn(Ml, K, device=dev).bfloat16()
        w = torch.randn(Nr, K, device=dev).bfloat16()
        xf = torch.empty(Ml * W, K, device=dev, dtype=torch.bfloat16)

        def base():
            dist.all_gather_into_tensor(xf, x)
            return xf @ w.t()

        t_f = timeit(lambda: tp.ag_gemm(x, w))
        t_b = timeit(base)
        t_mm = timeit(lambda: xf @ w.t())
        fl = 2.0 * Ml * W * K * Nr
        nv = (W - 1) * Ml * K * 2
        roof = max(fl / peak, nv / link) * 1e3
        res.append({"op": name, "shape": [Ml * W, Nr, K], "fused_ms": t_f, "nccl_cublas_ms": t_b, "cublas_only_ms": t_mm, "roofline_ms": roof, "frac_of_roofline": roof / t_f, "speedup_vs_nccl": t_b / t_f})
    for name, (M, Kr, N) in {"wo (gemm_rs)": (T, H // W, H), "down (gemm_rs)": (T, F // W, H)}.items():
        x = torch.randn(M, Kr, device=dev).bfloat16()
        w = torch.randn(N, Kr, device=dev).bfloat16()
        out = torch.empty(M // W, N, device=dev, dtype=torch.bfloat16)

        def base():
            y = x @ w.t()
            dist.reduce_scatter_tensor(out, y)
            return out

        t_f = timeit(lambda: tp.gemm_rs(x, w))
        t_b = timeit(base)
        t_mm = timeit(lambda: x @ w.t())
        t_n, err = None, None
        if tp_nvls is not None:  # GEMM into symmetric memory + switch-reduced pull (one multimem.ld_reduce per 16 bytes)
            ref = base().float()
            err = (tp_nvls.gemm_rs(x, w).float() - ref).abs().max().item() / ref.abs().max().item()
            t_n = timeit(lambda: tp_nvls.gemm_rs(x, w))
        fl = 2.0 * M * Kr * N
        nv = (W - 1) * (M // W) * N * 2
        roof = max(fl / peak, nv / link) * 1e3
        res.append({"op": name, "shape": [M, N, Kr], "fused_ms": t_f, "nccl_cublas_ms": t_b, "cublas_only_ms": t_mm, "roofline_ms": roof, "frac_of_roofline": roof / t_f, "speedup_vs_nccl": t_b / t_f,
                    "nvls_rs_ms": t_n, "nvls_rs_rel_err": err, "nvls_rs_speedup_vs_nccl": (t_b / t_n) if t_n else None})
    if rank == 0:
        for r in res:
            print(json.dumps(r), flush=True)
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump({"world": W, "results": res}, open(f"gpurun_out/tp_bench_w{W}.json", "w"), indent=1)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
