"""Checkpoint save under one layout, load under another (RaggedShard ↔ RaggedShard ↔ Shard ↔ Replicate),
FSDP model + optimizer round trip.  Parity: reference ``test/dtensor/checkpoint/test_ragged_shard_sl.py``."""
import os
import shutil
import tempfile

import torch
import torch.distributed as dist

from common import device_type, run_distributed


def _shared_dir(rank):
    box = [tempfile.mkdtemp(prefix="vb200_ckpt_") if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    return box[0]


def _reshard(rank, world):
    import vescale_b200.checkpoint as ckpt
    from vescale_b200 import Replicate, Shard, distribute_tensor, init_device_mesh
    from vescale_b200.dtensor import RaggedShard, zeros

    dev = device_type()
    mesh = init_device_mesh(dev, (world,))
    path = _shared_dir(rank)
    g = torch.Generator().manual_seed(0)
    full = {"w": torch.randn(12, 10, generator=g).to(dev), "v": torch.randn(24, generator=g).to(dev), "t": torch.randn(4, 6, 5, generator=g).to(dev)}
    src_pl = {"w": [RaggedShard((0,), (1, 5, 0, 6))], "v": [RaggedShard((0,), (3, 1, 1, 1))], "t": [RaggedShard((0, 1), (1, 1, 1, 3))]}
    state = {k: distribute_tensor(v, mesh, src_pl[k]) for k, v in full.items()}
    ckpt.save(path, {"model": state}, async_checkpoint=(world > 0))
    ckpt.wait_for_async()
    dist.barrier()
    assert os.path.exists(os.path.join(path, "model", ".metadata"))
    for dst_pl in (
        {"w": [RaggedShard((0,), (6, 0, 5, 1))], "v": [Shard(0)], "t": [RaggedShard((0,), (2, 0, 1, 1))]},
        {"w": [Shard(1)], "v": [Replicate()], "t": [Shard(2)]},
        {"w": [Replicate()], "v": [RaggedShard((0,), (0, 0, 1, 0))], "t": [Replicate()]},
    ):
        target = {k: zeros(*full[k].shape, device_mesh=mesh, placements=dst_pl[k]) for k in full}
        ckpt.load(path, {"model": target})
        for k in full:
            assert torch.equal(target[k].full_tensor(), full[k]), (k, dst_pl[k])
    # load on a different mesh shape (2 x 2)
    mesh2 = init_device_mesh(dev, (2, 2), mesh_dim_names=("a", "b"))
    target = {"w": zeros(12, 10, device_mesh=mesh2, placements=[RaggedShard((0,), (1, 2)), Shard(1)]), "v": zeros(24, device_mesh=mesh2, placements=[Shard(0), Shard(0)]), "t": zeros(4, 6, 5, device_mesh=mesh2, placements=[Replicate(), Shard(0)])}
    ckpt.load(path, {"model": target})
    for k in full:
        assert torch.equal(target[k].full_tensor(), full[k]), k
    dist.barrier()
    if rank == 0:
        shutil.rmtree(path, ignore_errors=True)


def _fsdp_roundtrip(rank, world):
    import vescale_b200.checkpoint as ckpt
    from vescale_b200 import init_device_mesh
    from vescale_b200.models import LlamaConfig, LlamaModel
    from vescale_b200.optim import FSDPAdamW
    from vescale_b200.parallel.fsdp import MixedPrecisionPolicy, fully_shard

    dev = device_type()
    cfg = LlamaConfig.tiny()
    mesh = init_device_mesh(dev, (world,))
    mp = MixedPrecisionPolicy(param_dtype=torch.float32)

    def build(seed):
        m = LlamaModel(cfg).reset_parameters(seed=seed).to(dev)
        for blk in m.layers:
            fully_shard(blk, mesh, mp_policy=mp)
        fully_shard(m.embed, mesh, mp_policy=mp)
        fully_shard(m.head, mesh, mp_policy=mp)
        fully_shard(m, mesh, mp_policy=mp)
        return m, FSDPAdamW(m, lr=1e-2)

    def step(m, o, s):
        g = torch.Generator().manual_seed(10 * s + rank)
        tok = torch.randint(0, cfg.vocab_size, (2, 17), generator=g).to(dev)
        loss = m(tok[:, :-1], tok[:, 1:])
        loss.backward()
        o.step()
        o.zero_grad()
        return loss.item()

    path = _shared_dir(rank)
    m1, o1 = build(1)
    for s in range(2):
        step(m1, o1, s)
    ckpt.save(path, {"model": m1, "optimizer": o1})
    dist.barrier()
    ref = [step(m1, o1, s) for s in range(2, 4)]
    m2, o2 = build(99)  # different init, then restored
    ckpt.load(path, {"model": m2, "optimizer": o2})
    for u in m2._fsdp_state.units:
        u.bf16_fresh = False  # master shards changed under the unit
    got = [step(m2, o2, s) for s in range(2, 4)]
    assert all(abs(a - b) < 1e-5 for a, b in zip(ref, got)), (ref, got)
    dist.barrier()
    if rank == 0:
        shutil.rmtree(path, ignore_errors=True)


def test_ragged_checkpoint_resharding():
    run_distributed(_reshard, 4)


def test_fsdp_model_optimizer_roundtrip():
    run_distributed(_fsdp_roundtrip, 4)


def test_mem_file_server_roundtrip(tmp_path):
    """In-memory file server: write / read / rename / listdir / persist / report (single process)."""
    from vescale_b200.checkpoint import MemFileClient, MemFileServer

    srv = MemFileServer().start()
    try:
        c = MemFileClient(srv.address)
        big = os.urandom(9 << 20)  # > 2 chunks
        assert c.write("ck/a.bin", big) == len(big) and c.write("ck/empty", b"") == 0
        assert c.read("ck/a.bin") == big and c.read("ck/empty") == b""
        assert c.exists("ck/a.bin") and c.exists("ck") and not c.exists("ck/zzz")
        c.rename("ck/a.bin", "ck/b.bin")
        assert c.listdir("ck") == ["ck/b.bin", "ck/empty"]
        assert c.persist("ck", str(tmp_path / "out")) == 2
        assert open(tmp_path / "out" / "b.bin", "rb").read() == big
        rep = c.report(who=3, status="saved", step=7)
        assert rep["reports"]["3"]["status"] == "saved" and rep["reports"]["3"]["step"] == 7 and rep["files"] == 2
        c.remove("ck/b.bin")
        try:
            c.read("ck/b.bin")
            raise AssertionError("expected FileNotFoundError")
        except FileNotFoundError:
            pass
    finally:
        srv.stop()


def _mem_ckpt(rank, world):
    """Save into the memory server (``mem://``), reload from it under another layout, persist to disk and reload from
    the persisted directory; second save of the same structure hits the plan cache; broadcast-on-load of replicated entries."""
    import vescale_b200.checkpoint as ckpt
    from vescale_b200 import Replicate, Shard, distribute_tensor, init_device_mesh
    from vescale_b200.checkpoint import MemFileClient, MemFileServer
    from vescale_b200.checkpoint.api import _PLANNERS
    from vescale_b200.dtensor import RaggedShard, zeros

    dev = device_type()
    mesh = init_device_mesh(dev, (world,))
    srv = MemFileServer().start() if rank == 0 else None
    box = [srv.address if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    addr = box[0]
    g = torch.Generator().manual_seed(0)
    full = {"w": torch.randn(12, 10, generator=g).to(dev), "b": torch.randn(7, generator=g).to(dev)}

    def make_state(scale):
        return {"w": distribute_tensor(full["w"] * scale, mesh, [RaggedShard((0,), (1, 5, 0, 6))]), "b": (full["b"] * scale).clone(), "step": int(scale)}

    for step in (1, 2):  # same structure twice: the second save re-uses the cached plan
        ckpt.save(f"mem://{addr}/run/step{step}", {"model": make_state(float(step))})
    dist.barrier()
    pl = _PLANNERS["model"]
    # the second save found its local plan in the cache on every rank and skipped the global planning step; the replicated "b"
    # (offered by all four ranks) was written by exactly one of them
    assert pl.cache.hits >= 1 and pl.global_plan_runs == (1 if dist.get_rank() == 0 else 0), (pl.cache.hits, pl.global_plan_runs)
    assert sum(1 for it in pl.plan.items if it.index.fqn == "b") in (0, 1)
    owners = [None] * world
    dist.all_gather_object(owners, sum(1 for it in pl.plan.items if it.index.fqn == "b"))
    assert sum(owners) == 1, owners
    client = MemFileClient(addr)
    names = client.listdir("run/step2/model")
    assert any(n.endswith(".metadata") for n in names) and sum(n.endswith(".distcp") for n in names) >= 1
    # reload from memory under another layout
    target = {"w": zeros(12, 10, device_mesh=mesh, placements=[Shard(1)]), "b": torch.zeros(7, device=dev), "step": 0}
    ckpt.load(f"mem://{addr}/run/step2", {"model": target})
    assert torch.equal(target["w"].full_tensor(), full["w"] * 2) and torch.equal(target["b"], full["b"] * 2)
    # background persistence, then an ordinary file-system load with broadcast of the replicated entries
    path = _shared_dir(rank)
    dist.barrier()
    if rank == 0:
        assert client.persist("run/step1", path) >= 2
    dist.barrier()
    target = {"w": zeros(12, 10, device_mesh=mesh, placements=[Replicate()]), "b": torch.zeros(7, device=dev), "step": 0}
    ckpt.load(path, {"model": target}, broadcast_checkpoint=True)
    assert torch.equal(target["w"].full_tensor(), full["w"]) and torch.equal(target["b"], full["b"])
    dist.barrier()
    if rank == 0:
        srv.stop()
        shutil.rmtree(path, ignore_errors=True)


def test_mem_checkpoint_plan_cache_broadcast_load():
    run_distributed(_mem_ckpt, 4)


def _save_ws(rank, world, path):
    """Phase 1 (4 ranks): train an FSDP model two steps, checkpoint model + optimizer, and record the full tensors."""
    import vescale_b200.checkpoint as ckpt
    from vescale_b200 import init_device_mesh
    from vescale_b200.models import LlamaConfig, LlamaModel
    from vescale_b200.optim import FSDPAdamW
    from vescale_b200.parallel.fsdp import MixedPrecisionPolicy, fully_shard

    dev = device_type()
    cfg = LlamaConfig.tiny()
    mesh = init_device_mesh(dev, (world,))
    mp = MixedPrecisionPolicy(param_dtype=torch.float32)
    m = LlamaModel(cfg).reset_parameters(seed=1).to(dev)
    for blk in m.layers:
        fully_shard(blk, mesh, mp_policy=mp)
    fully_shard(m, mesh, mp_policy=mp)
    o = FSDPAdamW(m, lr=1e-2)
    for s in range(2):
        tok = torch.randint(0, cfg.vocab_size, (2, 17), generator=torch.Generator().manual_seed(s)).to(dev)  # same batch on every rank
        m(tok[:, :-1], tok[:, 1:]).backward()
        o.step()
        o.zero_grad()
    ckpt.save(os.path.join(path, "ckpt"), {"model": m, "optimizer": o})
    full = {n: p.full_tensor().cpu() for n, p in m.named_parameters()}
    ostate = {k: {kk: vv.full_tensor().cpu() for kk, vv in v.items()} for k, v in o.state_dict()["state"].items()}
    tok = torch.randint(0, cfg.vocab_size, (2, 17), generator=torch.Generator().manual_seed(7)).to(dev)
    m(tok[:, :-1], tok[:, 1:]).backward()
    o.step()
    nxt = {n: p.full_tensor().cpu() for n, p in m.named_parameters()}
    if rank == 0:
        torch.save({"full": full, "opt": ostate, "next": nxt}, os.path.join(path, "golden.pt"))


def _load_ws(rank, world, path):
    """Phase 2 (2 ranks — a different world size and therefore different RaggedShard layouts): load, compare every parameter
    and optimizer moment with the golden full tensors, and check that the next optimizer step lands on the same weights."""
    import vescale_b200.checkpoint as ckpt
    from vescale_b200 import init_device_mesh
    from vescale_b200.models import LlamaConfig, LlamaModel
    from vescale_b200.optim import FSDPAdamW
    from vescale_b200.parallel.fsdp import MixedPrecisionPolicy, fully_shard

    dev = device_type()
    cfg = LlamaConfig.tiny()
    mesh = init_device_mesh(dev, (world,))
    mp = MixedPrecisionPolicy(param_dtype=torch.float32)
    m = LlamaModel(cfg).reset_parameters(seed=123).to(dev)
    for blk in m.layers:
        fully_shard(blk, mesh, mp_policy=mp)
    fully_shard(m, mesh, mp_policy=mp)
    o = FSDPAdamW(m, lr=1e-2)
    ckpt.load(os.path.join(path, "ckpt"), {"model": m, "optimizer": o})
    for u in m._fsdp_state.units:
        u.bf16_fresh = False
    gold = torch.load(os.path.join(path, "golden.pt"))
    for n, p in m.named_parameters():
        assert torch.equal(p.full_tensor().cpu(), gold["full"][n]), n
    for k, v in o.state_dict()["state"].items():
        for kk, vv in v.items():
            assert torch.equal(vv.full_tensor().cpu(), gold["opt"][k][kk]), (k, kk)
    assert o.step_count == 2
    tok = torch.randint(0, cfg.vocab_size, (2, 17), generator=torch.Generator().manual_seed(7)).to(dev)
    m(tok[:, :-1], tok[:, 1:]).backward()
    o.step()
    for n, p in m.named_parameters():
        torch.testing.assert_close(p.full_tensor().cpu(), gold["next"][n], rtol=1e-5, atol=1e-6, msg=n)


def test_checkpoint_reshards_across_world_sizes(tmp_path):
    """Save on 4 ranks, resume on 2 (reference ``checkpoint/open_llama/test_open_llama_dp_reshard.py`` strategy)."""
    run_distributed(_save_ws, 4, str(tmp_path))
    run_distributed(_load_ws, 2, str(tmp_path))


class _MLP(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a = torch.nn.Linear(16, 48)
        self.n = torch.nn.LayerNorm(48)
        self.b = torch.nn.Linear(48, 16)

    def forward(self, x):
        return self.b(self.n(torch.relu(self.a(x))))


def _dopt_batch(step):
    g = torch.Generator().manual_seed(step)
    return torch.randn(8, 16, generator=g), torch.randn(8, 16, generator=g)


def _dopt_build(dev, bucket):
    from vescale_b200.optim import DistributedOptimizer
    from vescale_b200.parallel.ddp import DistributedDataParallel as DDP

    torch.manual_seed(0)
    model = _MLP().to(dev)
    ddp = DDP(model, dist.group.WORLD, overlap_grad_reduce=False, use_distributed_optimizer=True, bucket_size=bucket)
    opt = DistributedOptimizer(torch.optim.AdamW(model.parameters(), lr=1e-2, weight_decay=0.05), [ddp])
    return model, ddp, opt


def _dopt_step(ddp, opt, step, dev):
    x, y = _dopt_batch(step)  # same batch on every rank: the trajectory does not depend on the DP size
    opt.zero_grad()
    torch.nn.functional.mse_loss(ddp(x.to(dev)), y.to(dev)).backward()
    opt.step()


def _dopt_save(rank, world, path):
    import vescale_b200.checkpoint as ckpt

    dev = device_type()
    model, ddp, opt = _dopt_build(dev, bucket=500)
    for s in range(2):
        _dopt_step(ddp, opt, s, dev)
    ckpt.save(os.path.join(path, "ckpt"), {"model": model, "optimizer": opt})
    _dopt_step(ddp, opt, 2, dev)
    if rank == 0:
        torch.save({n: p.detach().cpu() for n, p in model.named_parameters()}, os.path.join(path, "golden.pt"))


def _dopt_load(rank, world, path):
    """Different DP size *and* bucket size: every parameter's moments are split over the ranks differently than when saved."""
    import vescale_b200.checkpoint as ckpt

    dev = device_type()
    model, ddp, opt = _dopt_build(dev, bucket=1300)
    with torch.no_grad():
        for p in model.parameters():
            p.add_(1.0)  # make sure the load really restores the weights
    ckpt.load(os.path.join(path, "ckpt"), {"model": model, "optimizer": opt})
    _dopt_step(ddp, opt, 2, dev)
    gold = torch.load(os.path.join(path, "golden.pt"))
    for n, p in model.named_parameters():
        torch.testing.assert_close(p.detach().cpu(), gold[n], rtol=1e-5, atol=1e-6, msg=n)


def test_distributed_optimizer_state_reshards_across_dp_sizes(tmp_path):
    """ZeRO-2+ optimizer state saved on 4 DP ranks resumes on 2 (legacy ``test_open_llama_dp_reshard.py`` strategy)."""
    run_distributed(_dopt_save, 4, str(tmp_path))
    run_distributed(_dopt_load, 2, str(tmp_path))


class _GPTBlock(torch.nn.Module):
    def __init__(self, h=32):
        super().__init__()
        self.ln_1 = torch.nn.LayerNorm(h)
        self.c_fc = torch.nn.Linear(h, 4 * h)
        self.c_proj = torch.nn.Linear(4 * h, h)

    def forward(self, x):
        return x + self.c_proj(torch.nn.functional.gelu(self.c_fc(self.ln_1(x))))


def _tp_model(dev, seed):
    from vescale_b200 import Replicate, init_device_mesh
    from vescale_b200.parallel.dmp import auto_parallelize_module

    mesh = init_device_mesh(dev, (dist.get_world_size(),), mesh_dim_names=("TP",))
    torch.manual_seed(seed)
    model = torch.nn.Sequential(_GPTBlock(), _GPTBlock()).to(dev)
    for blk in model:
        auto_parallelize_module(blk, mesh, "MEGATRON", plan_override={"forward": {r"input": [[Replicate()]], r"c_proj\.output": [[Replicate()]]}})
    return model


def _tp_save(rank, world, path):
    import vescale_b200.checkpoint as ckpt

    dev = device_type()
    model = _tp_model(dev, seed=0)
    x = torch.randn(3, 5, 32, generator=torch.Generator().manual_seed(1)).to(dev)
    y = model(x)
    ckpt.save(os.path.join(path, "ckpt"), {"model": model})
    if rank == 0:
        torch.save(y.full_tensor().detach().cpu(), os.path.join(path, "y.pt"))


def _tp_load(rank, world, path):
    import vescale_b200.checkpoint as ckpt

    dev = device_type()
    model = _tp_model(dev, seed=123)  # different weights, different TP degree
    ckpt.load(os.path.join(path, "ckpt"), {"model": model})
    x = torch.randn(3, 5, 32, generator=torch.Generator().manual_seed(1)).to(dev)
    torch.testing.assert_close(model(x).full_tensor().detach().cpu(), torch.load(os.path.join(path, "y.pt")), rtol=1e-5, atol=1e-6)


def test_dmodule_checkpoint_reshards_across_tp_degrees(tmp_path):
    """A DModule (auto-planned TP/SP) saved at TP=4 reloads at TP=2 and computes the same function
    (legacy ``dmodule/test_saveload.py`` + ``checkpoint/open_llama/test_open_llama_tp_reshard.py``)."""
    run_distributed(_tp_save, 4, str(tmp_path))
    run_distributed(_tp_load, 2, str(tmp_path))


def _pp_layout_and_workers(rank, world, path):
    """Optimizer state of a pipelined model: one DCP checkpoint per stage under ``optimizer/pp_{rank}`` saved within the stage's
    process group (legacy layout, ``api/vescale_checkpointer.py:71-249``); files are serialised by worker processes
    (``storage.ProcessPoolWriter``, legacy ``storage/filesystem.py:401-460``) and reshard on load."""
    import vescale_b200.checkpoint as ckpt
    from vescale_b200 import Shard, distribute_tensor, init_device_mesh

    pp_rank = rank // 2
    groups = [dist.new_group([0, 1]), dist.new_group([2, 3])]
    from vescale_b200.mesh import DeviceMesh

    meshes = [DeviceMesh(device_type(), [0, 1]), DeviceMesh(device_type(), [2, 3])]  # every rank builds both (group creation is collective)
    mesh = meshes[pp_rank]
    w = torch.arange(48.0).view(6, 8) + 1000 * pp_rank  # every stage owns different optimizer state
    st = {"exp_avg": distribute_tensor(w.to(device_type()), mesh, [Shard(0)], src_data_rank=None), "step": 7 + pp_rank}
    # the background half of the save coordinates over the report service (no collective on any process group from that thread)
    from vescale_b200.checkpoint import server_lib

    box = [server_lib.start_server_in_new_process(world) if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ckpt.save(path, {"optimizer": st}, async_checkpoint=True, workers=2, pp_rank=pp_rank, pp_group=groups[pp_rank], coordinator_address=box[0])
    ckpt.wait_for_async()
    dist.barrier()
    if rank == 0:
        status = server_lib.get_server_status(server_lib.get_stub(box[0]))
        assert status["waiting"] == {} and status["completed"] >= 2 * 6, status  # both stages ran their rendezvous through it
        for p in server_lib.start_server_in_new_process.processes:
            p.terminate()
    d = os.path.join(path, "optimizer", f"pp_{pp_rank}")
    files = sorted(os.listdir(d))
    assert ".metadata" in files and sum(f.endswith(".distcp") for f in files) >= 2, files
    assert sorted(os.listdir(os.path.join(path, "optimizer"))) == ["pp_0", "pp_1"]
    tgt = {"exp_avg": distribute_tensor(torch.zeros(6, 8).to(device_type()), mesh, [Shard(1)], src_data_rank=None), "step": 0}
    ckpt.load(path, {"optimizer": tgt}, pp_rank=pp_rank, pp_group=groups[pp_rank])
    torch.testing.assert_close(tgt["exp_avg"].full_tensor().cpu(), w)


def test_pp_optimizer_layout_and_process_pool_writer(tmp_path):
    run_distributed(_pp_layout_and_workers, 4, str(tmp_path))


def test_bfile_schemes_recorder_and_logger(tmp_path, monkeypatch):
    """``bfile``: one file API over local paths, the in-memory file server (``mem://``) and registered schemes, atomic writes;
    ``TorchCheckpointRecorder`` turns ``torch.save`` calls asynchronous; checkpoint logger level from the environment
    (legacy ``checkpoint/utilities/{bfile,mem_checkpoint,logger}.py``)."""
    import logging

    from vescale_b200.checkpoint import MemFileServer, TorchCheckpointRecorder, bfile, get_vescale_checkpoint_logger

    # local
    p = str(tmp_path / "a" / "b" / "f.bin")
    bfile.safe_atomic_write(p, b"hello")
    assert bfile.exists(p) and bfile.read_bytes(p) == b"hello" and bfile.is_local_path(p) and bfile.get_schema(p) == bfile.FileType.LOCAL
    assert bfile.listdir(str(tmp_path / "a" / "b")) == ["f.bin"]  # no temporary file left behind
    bfile.rename(p, p + "2")
    assert not bfile.exists(p) and bfile.local_list_folder(str(tmp_path / "a"), recursive=True) == [p + "2"]
    bfile.remove(str(tmp_path / "a"))
    assert not bfile.exists(str(tmp_path / "a"))
    # in-memory file server
    srv = MemFileServer().start()
    try:
        root = f"mem://{srv.address}/ck"
        bfile.safe_atomic_write(root + "/x.bin", b"abc")
        with bfile.BFile(root + "/t.txt", "w") as f:
            f.write("text")
        assert bfile.get_schema(root) == bfile.FileType.LOCAL_MEM and sorted(bfile.listdir(root)) == ["t.txt", "x.bin"]
        assert bfile.read_bytes(root + "/x.bin") == b"abc"
        with bfile.BFile(root + "/t.txt", "r") as f:
            assert f.read() == "text"
        bfile.remove(root)
        assert not bfile.exists(root + "/x.bin")
    finally:
        srv.stop()
    # a deployment's own object store
    store = {}

    class Dict:
        def open(self, path, mode="r"):
            import io

            if "r" in mode:
                return io.BytesIO(store[path])
            b = io.BytesIO()
            close = b.close
            b.close = lambda: (store.__setitem__(path, b.getvalue()), close())[1]
            return b

        def exists(self, path):
            return path in store

        def listdir(self, path):
            return [k[len(path) + 1 :] for k in store if k.startswith(path + "/")]

        def remove(self, path):
            store.pop(path, None)

        def rename(self, src, dst, overwrite=False):
            store[dst] = store.pop(src)

        def makedirs(self, path):
            pass

    bfile.register_scheme("objstore", Dict())
    bfile.safe_atomic_write("objstore://bucket/ck/meta", b"m")
    assert store == {"objstore://bucket/ck/meta": b"m"} and bfile.get_schema("objstore://bucket") == bfile.FileType.REMOTE and not bfile.is_local_path("objstore://x")
    try:
        bfile.exists("nosuch://x")
        raise AssertionError("unknown scheme must raise")
    except ValueError:
        pass
    # torch.save recorder
    sd = {"w": torch.arange(12.0).view(3, 4), "step": 3, "nested": [torch.ones(2), ("a", torch.zeros(1))]}
    rec = TorchCheckpointRecorder()
    with rec:
        torch.save(sd, str(tmp_path / "r" / "model.pt"))
        sd["w"].mul_(0)  # the training loop moves on; the recorded copy must not see it
        torch.save({"x": torch.ones(1)}, str(tmp_path / "r" / "model.pt"))  # same path again: waits for the first write
        torch.save(sd, str(tmp_path / "r" / "second.pt"))
    assert torch.save.__module__.startswith("torch")  # restored
    written = rec.wait()
    assert set(written) == {str(tmp_path / "r" / "model.pt"), str(tmp_path / "r" / "second.pt")} and all(v > 0 for v in written.values())
    assert torch.equal(torch.load(str(tmp_path / "r" / "model.pt"))["x"], torch.ones(1))
    assert float(torch.load(str(tmp_path / "r" / "second.pt"))["w"].abs().sum()) == 0.0
    rec.close()
    # logger level from the environment
    monkeypatch.setenv("VESCALE_CHECKPOINT_LOGGING_LEVEL", "DEBUG")
    assert get_vescale_checkpoint_logger().level == logging.DEBUG
    monkeypatch.setenv("VESCALE_CHECKPOINT_LOGGING_LEVEL", "35")
    assert get_vescale_checkpoint_logger().level == 35
