import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
dist.init_process_group("nccl"); rank, W = dist.get_rank(), dist.get_world_size()
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0))); dev = torch.device("cuda", torch.cuda.current_device())
from vescale_b200 import init_device_mesh
from vescale_b200.comm.fused_tp import FusedTP
mesh = init_device_mesh("cuda", (W,), mesh_dim_names=("TP",)); tp = FusedTP(mesh, "TP", dev)
g = torch.Generator(device=dev)
def P(*a):
    if rank == 0: print(*a, flush=True)
for (Ml, K, Nr) in ((256, 512, 256), (1024, 4096, 3072), (512, 1024, 264)):
    for it in range(3):
        g.manual_seed(1000 * it + rank)
        x = (torch.randn(Ml, K, device=dev, generator=g) * 0.5).bfloat16(); w = (torch.randn(Nr, K, device=dev, generator=g) * 0.05).bfloat16()
        xs = [torch.empty_like(x) for _ in range(W)]; dist.all_gather(xs, x)
        ref = torch.cat(xs).float() @ w.float().t()
        y, xf = tp.ag_gemm(x, w); torch.cuda.synchronize()
        P("ag", Ml, K, Nr, it, "gather ok" if torch.equal(xf, torch.cat(xs)) else "GATHER BAD", (y.float()-ref).abs().max().item(), ref.abs().max().item())
for (M, Kr, N) in ((256 * W, 512, 256), (2048 * W // 2, 2048, 4096), (512 * W, 1024, 264)):
    for it in range(3):
        g.manual_seed(77 * it + rank)
        x = (torch.randn(M, Kr, device=dev, generator=g) * 0.5).bfloat16(); w = (torch.randn(N, Kr, device=dev, generator=g) * 0.05).bfloat16()
        tot = (x.float() @ w.float().t()); dist.all_reduce(tot); ref = tot[rank * (M // W):(rank + 1) * (M // W)]
        y = tp.gemm_rs(x, w); torch.cuda.synchronize()
        P("rs", M, Kr, N, it, (y.float()-ref).abs().max().item(), ref.abs().max().item()); dist.barrier()
P("done")
dist.destroy_process_group()
