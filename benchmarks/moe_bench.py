"""Mixtral-8x7B MoE layer (H=4096, F=14336, 8 experts, top-2, 8192 tokens per rank), expert parallel over all ranks:
device-side symmetric-memory dispatch / grouped tcgen05 GEMM / combine vs the NCCL all_to_all_single path with its host
sync (legacy ``moe/_scheduler.py:165-215``).  torchrun, >= 2 GPUs; device-timed, max over ranks."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from benchmarks.tp_bench import timeit  # noqa: E402


def main():
    dist.init_process_group("nccl")
    rank, W = dist.get_rank(), dist.get_world_size()
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    dev = torch.device("cuda", torch.cuda.current_device())
    from vescale_b200 import init_device_mesh
    from vescale_b200.parallel.moe import MoEConfig, MoELayer
    from vescale_b200.parallel.moe.symm_dispatch import SymmMoEDispatcher

    mesh = init_device_mesh("cuda", (W,), mesh_dim_names=("EP",))
    H, F, E, k, T = 4096, 14336, 8, 2, 8192
    cfg = MoEConfig(H, F, E, k, dtype=torch.bfloat16)
    layer = MoELayer(cfg, mesh.get_group(0), device=dev)
    layer.reset_parameters(torch.Generator(device=dev).manual_seed(1))
    disp = SymmMoEDispatcher(mesh, E, H, F, max_tokens=T, top_k=k, capacity_factor=2.0, device=dev)
    x = torch.randn(T, H, device=dev, generator=torch.Generator(device=dev).manual_seed(rank)).bfloat16().requires_grad_()

    def fwd():
        with torch.no_grad():
            return layer(x)

    def fwd_bwd():
        y = layer(x)
        y.backward(y.detach())
        layer.zero_grad(set_to_none=True)
        x.grad = None

    res = {}
    for name in ("nccl", "symm"):
        layer.symm_dispatcher = None
        if name == "symm":
            layer.use_symmetric_dispatch(disp)
        res[name] = {"fwd_ms": timeit(fwd), "fwd_bwd_ms": timeit(fwd_bwd)}
    flops = 2.0 * T * k * 3 * H * F  # per rank, forward
    out = {"world": W, "tokens_per_rank": T, "experts": E, "top_k": k, "results": res,
           "fwd_speedup": res["nccl"]["fwd_ms"] / res["symm"]["fwd_ms"], "fwd_bwd_speedup": res["nccl"]["fwd_bwd_ms"] / res["symm"]["fwd_bwd_ms"],
           "symm_fwd_model_tflops_per_gpu": flops / res["symm"]["fwd_ms"] / 1e9}
    if rank == 0:
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
