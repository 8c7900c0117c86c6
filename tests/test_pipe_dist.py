"""Pipeline parallel: every schedule on 4 ranks must reproduce single-process loss and gradients
(``legacy/test/parallel/pipeline/e2e/test_pp_accuracy_alignment.py`` strategy); scheduler unit tests."""
import copy

import pytest
import torch
import torch.nn as nn

from common import device_type, run_distributed


class Blk(nn.Module):
    def __init__(self, h):
        super().__init__()
        self.l = nn.Linear(h, h)

    def forward(self, x):
        return x + torch.tanh(self.l(x))


def make_model(n=8, h=16):
    torch.manual_seed(0)
    return nn.Sequential(*[Blk(h) for _ in range(n)])


def _pp(rank, world, sched_name, tracer):
    from vescale_b200 import init_device_mesh
    from vescale_b200.parallel.pipe import PipeEngine, PipelineParallelPlan, PipelineScheduleType, TracerType, construct_pipeline_stage

    dev = device_type()
    ref = make_model().to(dev)
    model = copy.deepcopy(ref)
    mesh = init_device_mesh(dev, (world,), mesh_dim_names=("PP",))
    plan = PipelineParallelPlan(num_stages=world, schedule_type=PipelineScheduleType[sched_name], tracer_type=TracerType[tracer],
                                example_inputs=(torch.randn(3, 16).to(dev),) if tracer == "EXPORT" else None)
    pm = construct_pipeline_stage(model, plan, mesh)
    M = 8
    g = torch.Generator().manual_seed(5)
    xs = [torch.randn(3, 16, generator=g).to(dev) for _ in range(M)]
    ys = [torch.randn(3, 16, generator=g).to(dev) for _ in range(M)]
    loss_fn = lambda out, y: torch.nn.functional.mse_loss(out, y)
    engine = PipeEngine(pm, mesh, loss_fn, plan)
    for it in range(2):
        loss, _ = engine(xs, ys)
        ref_loss = sum(loss_fn(ref(x), y) / M for x, y in zip(xs, ys))
        ref_loss.backward()
        if engine.is_last_rank:
            torch.testing.assert_close(loss, ref_loss.detach(), rtol=1e-5, atol=1e-6)
    # gradients (accumulated over two identical passes) match the golden model's
    ref_params = dict(ref.named_parameters())
    units = [f"{i}" for i in range(8)]
    checked = 0
    for c in range(pm.num_chunks):
        stage = pm.chunk(c)
        names = getattr(stage, "names", None)
        for n, p in stage.named_parameters():
            if names is not None:
                _, idx, rest = n.split(".", 2)
                fq = f"{names[int(idx)]}.{rest}"
            else:
                fq = n.replace("_", ".", 1) if n[0].isdigit() is False else n
                continue
            torch.testing.assert_close(p.grad, ref_params[fq].grad, rtol=1e-4, atol=1e-6, msg=fq)
            checked += 1
    assert checked > 0 or tracer in ("FX", "EXPORT")


@pytest.mark.parametrize("sched", ["GPIPE", "SIMPLE_1F1B", "INTERLEAVED_1F1B", "ZERO_BUBBLE", "ZERO_BUBBLE_V"])
def test_pipeline_schedules_match_single_process(sched):
    run_distributed(_pp, 4, sched, "STRUCTURAL")


def _pp_overlap(rank, world, batch):
    """p2p is overlapped: receives are prefetched from the pair order, the shape record travels once, sends + following
    receives form one batch group when the plan asks for it, and every op shows up on the ndtimeline."""
    from vescale_b200 import init_device_mesh
    from vescale_b200 import profiler as ndt
    from vescale_b200.parallel.pipe import PipeEngine, PipelineParallelPlan, PipelineScheduleType, construct_pipeline_stage

    dev = device_type()
    seen = []

    class Capture(ndt.NDHandler):
        def __call__(self, records, rank_, step):
            seen.extend(r["metric"] for r in records)

    ndt.init_ndtimers(rank=rank, world_size=world, handlers=[Capture()])
    ref = make_model().to(dev)
    model = copy.deepcopy(ref)
    mesh = init_device_mesh(dev, (world,), mesh_dim_names=("PP",))
    plan = PipelineParallelPlan(num_stages=world, schedule_type=PipelineScheduleType.SIMPLE_1F1B, batch_p2p_comm=batch)
    pm = construct_pipeline_stage(model, plan, mesh)
    M = 8
    g = torch.Generator().manual_seed(5)
    xs = [torch.randn(3, 16, generator=g).to(dev) for _ in range(M)]
    ys = [torch.randn(3, 16, generator=g).to(dev) for _ in range(M)]
    loss_fn = lambda out, y: torch.nn.functional.mse_loss(out, y)
    engine = PipeEngine(pm, mesh, loss_fn, plan)
    engine(xs, ys)
    first = dict(engine.schedule_engine.last_p2p.stats)
    loss, _ = engine(xs, ys)
    st = engine.schedule_engine.last_p2p.stats
    assert first["meta_messages"] > 0 and st["meta_messages"] == 0, (first, st)  # shapes handshaken once, then reused
    n_recv = (M if rank > 0 else 0) + (M if rank < world - 1 else 0)
    assert st["prefetched_recvs"] >= n_recv - 2, (st, n_recv)  # all but the very first receive of each direction were posted early
    if batch and 0 < rank < world - 1:
        assert st["batched_groups"] > 0, st
    ref_loss = sum(loss_fn(ref(x), y) / M for x, y in zip(xs, ys))
    if engine.is_last_rank:
        torch.testing.assert_close(loss, ref_loss.detach(), rtol=1e-5, atol=1e-6)
    ndt.flush(asynchronous=False)
    assert "forward-compute" in seen and "backward-compute" in seen, set(seen)
    if rank < world - 1:
        assert "send-forward" in seen and "recv-backward" in seen, set(seen)
    if rank > 0:
        assert "recv-forward" in seen and "send-backward" in seen, set(seen)


@pytest.mark.parametrize("batch", [False, True])
def test_pipeline_p2p_overlap_and_timeline(batch):
    run_distributed(_pp_overlap, 4, batch)


def _pp_hf_graph(rank, world):
    """Graph tracer (torch.export + liveness pass, ``pipe/trace.py``) on an UNMODIFIED HuggingFace Llama: rotary tables and the
    mask skip over stages, the stages still form a chain; 1F1B over 2 ranks reproduces the single-process loss and gradients."""
    from transformers import LlamaConfig, LlamaForCausalLM

    from vescale_b200 import init_device_mesh
    from vescale_b200.parallel.pipe import PipeEngine, PipelineParallelPlan, PipelineScheduleType, TracerType, construct_pipeline_stage

    dev = device_type()
    cfg = LlamaConfig(vocab_size=128, hidden_size=32, intermediate_size=64, num_hidden_layers=4, num_attention_heads=4, num_key_value_heads=2,
                      max_position_embeddings=64, attn_implementation="sdpa", use_cache=False, tie_word_embeddings=False)
    torch.manual_seed(0)
    ref = LlamaForCausalLM(cfg).to(dev)
    model = copy.deepcopy(ref)
    g = torch.Generator().manual_seed(9)
    M = 4
    xs = [torch.randint(0, 128, (2, 16), generator=g).to(dev) for _ in range(M)]
    ys = [torch.randint(0, 128, (2, 16), generator=g).to(dev) for _ in range(M)]
    mesh = init_device_mesh(dev, (world,), mesh_dim_names=("PP",))
    plan = PipelineParallelPlan(num_stages=world, schedule_type=PipelineScheduleType.SIMPLE_1F1B, tracer_type=TracerType.GRAPH, example_inputs=(xs[0],))
    pm = construct_pipeline_stage(model, plan, mesh)
    loss_fn = lambda out, y: torch.nn.functional.cross_entropy((out[0] if isinstance(out, tuple) else out).reshape(-1, 128), y.reshape(-1))
    engine = PipeEngine(pm, mesh, loss_fn, plan)
    loss, _ = engine(xs, ys)
    ref_loss = sum(loss_fn(ref(x).logits, y) / M for x, y in zip(xs, ys))
    ref_loss.backward()
    if engine.is_last_rank:
        torch.testing.assert_close(loss, ref_loss.detach(), rtol=1e-5, atol=1e-6)
    ref_params = dict(ref.named_parameters())
    checked = 0
    for c in range(pm.num_chunks):
        st = pm.chunk(c)
        for key, p in st.weights.items():
            torch.testing.assert_close(p.grad, ref_params[st.names[key]].grad, rtol=1e-4, atol=1e-6, msg=st.names[key])
            checked += 1
    assert checked > 0


def test_pipeline_graph_tracer_hf_llama():
    run_distributed(_pp_hf_graph, 2)


def test_pipeline_fx_tracer():
    run_distributed(_pp, 2, "SIMPLE_1F1B", "FX")


def test_pipeline_export_tracer():
    run_distributed(_pp, 2, "SIMPLE_1F1B", "EXPORT")


def test_scheduler_properties():
    from vescale_b200.parallel.pipe import PipelineParallelPlan, PipelineScheduleType, build_schedule, bubble_fraction, split_units, PipelineSplitMethodType

    for st in PipelineScheduleType:
        plan = PipelineParallelPlan(num_stages=4, schedule_type=st)
        rows = build_schedule(plan, 8)
        nv = 4 * plan.virtual_chunks
        fs = sorted((i.microbatch, i.vstage) for r in rows for i in r if i.kind == "F")
        assert fs == sorted((m, v) for m in range(8) for v in range(nv))
        # dependencies respected in time
        end = {(i.kind, i.microbatch, i.vstage): i.end for r in rows for i in r}
        start = {(i.kind, i.microbatch, i.vstage): i.start for r in rows for i in r}
        for (k, m, v), s in start.items():
            if k == "F" and v > 0:
                assert end[("F", m, v - 1)] <= s + 1e-9
            if k == "B":
                assert end[("F", m, v)] <= s + 1e-9
                if v < nv - 1:
                    assert end[("B", m, v + 1)] <= s + 1e-9
    b1 = bubble_fraction(build_schedule(PipelineParallelPlan(num_stages=4, schedule_type=PipelineScheduleType.SIMPLE_1F1B), 8))
    bz = bubble_fraction(build_schedule(PipelineParallelPlan(num_stages=4, schedule_type=PipelineScheduleType.ZERO_BUBBLE), 8))
    bv = bubble_fraction(build_schedule(PipelineParallelPlan(num_stages=4, schedule_type=PipelineScheduleType.ZERO_BUBBLE_V), 8))
    assert bv < bz < b1
    # interleaved 1F1B: fixed Megatron operation order -> the textbook bubble (P-1)(tF+tB) / (M V (tF+tB) + (P-1)(tF+tB)) whenever
    # P divides M, and no deadlock for long runs (a greedy in-flight window deadlocks at M > 2P)
    for P, V, M in ((4, 2, 8), (4, 2, 16), (4, 4, 8), (8, 2, 16), (4, 3, 12), (2, 2, 4)):
        rows = build_schedule(PipelineParallelPlan(num_stages=P, virtual_chunks=V, schedule_type=PipelineScheduleType.INTERLEAVED_1F1B), M)
        assert all(len(r) == 2 * M * V for r in rows)
        assert abs(bubble_fraction(rows) - (P - 1) / (M * V + P - 1)) < 1e-9, (P, V, M)
    for P, V, M in ((4, 2, 6), (2, 2, 5), (2, 3, 7), (4, 2, 64)):
        rows = build_schedule(PipelineParallelPlan(num_stages=P, virtual_chunks=V, schedule_type=PipelineScheduleType.INTERLEAVED_1F1B), M)
        assert all(len(r) == 2 * M * V for r in rows)
    from vescale_b200.parallel.pipe import validate_pipeline_schedule
    from vescale_b200.parallel.pipe._schedules import InterleavedOneFOneBInstructionGenerator, OneFOneBInstrcutionGenerator, StageDeps

    gen = InterleavedOneFOneBInstructionGenerator(StageDeps(8), [None] * 4, 8)
    assert len(gen.schema.rows[3]) == 32 and sum(i.name in ('FWD', 'BWD') for i in gen.get_instruction_list(3)) == 32 and gen.bubble_fraction() < OneFOneBInstrcutionGenerator(StageDeps(4), [None] * 4, 8).bubble_fraction()
    try:
        validate_pipeline_schedule(PipelineParallelPlan(num_stages=2, virtual_chunks=2, schedule_type=PipelineScheduleType.SIMPLE_1F1B))
        raise AssertionError("SIMPLE_1F1B with two chunks must be rejected")
    except ValueError:
        pass
    m = make_model(10)
    units = list(m.named_children())
    g = split_units(units, PipelineParallelPlan(num_stages=4, split_method=PipelineSplitMethodType.UNIFORM))
    assert [len(x) for x in g] == [3, 3, 2, 2]
    g = split_units(units, PipelineParallelPlan(num_stages=3, split_method=PipelineSplitMethodType.MANUAL, split_points=["1", "5"]))
    assert [len(x) for x in g] == [2, 4, 4]
    g = split_units(units, PipelineParallelPlan(num_stages=5, split_method=PipelineSplitMethodType.PARAMETERS))
    assert sum(len(x) for x in g) == 10 and all(len(x) >= 1 for x in g)


def test_cost_driven_schedule_search():
    """The schedule search (legacy ``zero_bubble_v.py:198-600``: candidates under a memory limit, keep the fastest) over the list
    scheduler's free choices: never worse than the classic schedule, reaches the lower bound where the classic one does not (ZB-V /
    zero-bubble once p2p latency is non-zero), honours the activation-memory bound, rejects impossible bounds."""
    from vescale_b200.parallel.pipe import (PipelineParallelPlan, PipelineScheduleType, build_schedule, check_schedule, lower_bound, makespan, peak_memory,
                                            search_schedule)

    for st in PipelineScheduleType:
        for comm in (0.0, 0.1):
            plan = PipelineParallelPlan(num_stages=4, schedule_type=st, costs={"F": 1.0, "B": 1.2, "W": 0.8, "comm": comm})
            classic = build_schedule(plan, 8)
            check_schedule(classic, plan, 8)
            res = search_schedule(plan, 8)
            check_schedule(res.rows, plan, 8)
            assert res.makespan <= makespan(classic) + 1e-9 and res.makespan >= lower_bound(plan, 8) - 1e-9, (st, comm)
    for P, M in ((4, 8), (4, 12), (8, 16)):
        plan = PipelineParallelPlan(num_stages=P, schedule_type=PipelineScheduleType.ZERO_BUBBLE_V, costs={"F": 1.0, "B": 1.0, "W": 1.0, "comm": 0.1})
        res = search_schedule(plan, M)
        assert makespan(build_schedule(plan, M)) > res.makespan + 0.5 and abs(res.makespan - lower_bound(plan, M)) < 1e-6, (P, M, res.summary())
    # memory-bounded: every returned schedule stays under the bound, looser bounds are never slower
    prev = None
    for mm in (3, 4, 6, 8, 12):
        plan = PipelineParallelPlan(num_stages=4, schedule_type=PipelineScheduleType.ZERO_BUBBLE_V, costs={"F": 1.0, "B": 1.0, "W": 1.0, "comm": 0.1}, max_mem=mm)
        res = search_schedule(plan, 12)
        check_schedule(res.rows, plan, 12)
        assert max(peak_memory(res.rows, plan.mem_costs)) <= mm + 1e-9
        assert prev is None or res.makespan <= prev + 1e-9
        prev = res.makespan
        plan.auto_schedule = True  # build_schedule delegates to the search
        assert makespan(build_schedule(plan, 12)) == res.makespan
    with pytest.raises(RuntimeError):
        search_schedule(PipelineParallelPlan(num_stages=4, schedule_type=PipelineScheduleType.ZERO_BUBBLE_V, max_mem=0.5), 8)


def _pp_auto(rank, world, sched_name):
    """calibrate() measures F / B / W on the running pipeline, switches the plan to the searched schedule (here under a memory bound
    that the classic ZB-V window would exceed); loss and gradients still equal the single-process model."""
    from vescale_b200 import init_device_mesh
    from vescale_b200.parallel.pipe import PipeEngine, PipelineParallelPlan, PipelineScheduleType, construct_pipeline_stage, peak_memory

    dev = device_type()
    ref = make_model().to(dev)
    model = copy.deepcopy(ref)
    mesh = init_device_mesh(dev, (world,), mesh_dim_names=("PP",))
    plan = PipelineParallelPlan(num_stages=world, schedule_type=PipelineScheduleType[sched_name], max_mem=5.0)
    pm = construct_pipeline_stage(model, plan, mesh)
    M = 8
    g = torch.Generator().manual_seed(5)
    xs = [torch.randn(3, 16, generator=g).to(dev) for _ in range(M)]
    ys = [torch.randn(3, 16, generator=g).to(dev) for _ in range(M)]
    loss_fn = lambda out, y: torch.nn.functional.mse_loss(out, y)
    engine = PipeEngine(pm, mesh, loss_fn, plan)
    costs = engine.calibrate(xs, ys, comm=0.05)
    assert plan.auto_schedule and costs["F"] == 1.0 and costs["B"] > 0 and all(p.grad is None for p in pm.parameters())
    split = sched_name in ("ZERO_BUBBLE", "ZERO_BUBBLE_V")
    assert (costs["W"] > 0) == split
    # every rank measured the same (MAX-reduced) costs, so every rank searches the same schedule
    t = torch.tensor([costs["B"], costs["W"]], dtype=torch.float64)
    lst = [torch.zeros_like(t) for _ in range(world)]
    torch.distributed.all_gather(lst, t)
    assert all(torch.equal(x, t) for x in lst)
    loss, _ = engine(xs, ys)
    rows = engine.schedule_engine.schedule(M)
    eff = plan.mem_costs if split else {"F": 1.0, "B": -1.0, "W": 0.0}
    assert max(peak_memory(rows, eff)) <= 5.0 + 1e-9
    ref_loss = sum(loss_fn(ref(x), y) / M for x, y in zip(xs, ys))
    ref_loss.backward()
    if engine.is_last_rank:
        torch.testing.assert_close(loss, ref_loss.detach(), rtol=1e-5, atol=1e-6)
    ref_params = dict(ref.named_parameters())
    checked = 0
    for c in range(pm.num_chunks):
        stage = pm.chunk(c)
        for n, p in stage.named_parameters():
            _, idx, rest = n.split(".", 2)
            torch.testing.assert_close(p.grad, ref_params[f"{stage.names[int(idx)]}.{rest}"].grad, rtol=1e-4, atol=1e-6)
            checked += 1
    assert checked > 0


@pytest.mark.parametrize("sched", ["SIMPLE_1F1B", "ZERO_BUBBLE_V"])
def test_pipeline_calibrated_auto_schedule(sched):
    run_distributed(_pp_auto, 4, sched)


def _pp_graph_mode(rank, world, sched_name):
    """Graph mode (legacy compile mode, ``pp_collective_emitter.py``): the rank's program is ONE fx graph with the functional p2p ops
    in it; backward communication comes from autograd.  Loss and gradients equal the single-process model."""
    from vescale_b200 import init_device_mesh
    from vescale_b200.parallel.pipe import ModeType, PipeEngine, PipelineParallelPlan, PipelineScheduleType, construct_pipeline_stage

    dev = device_type()
    ref = make_model().to(dev)
    model = copy.deepcopy(ref)
    mesh = init_device_mesh(dev, (world,), mesh_dim_names=("PP",))
    plan = PipelineParallelPlan(num_stages=world, schedule_type=PipelineScheduleType[sched_name], mode=ModeType.GRAPH_EAGER)
    pm = construct_pipeline_stage(model, plan, mesh)
    M = 4
    g = torch.Generator().manual_seed(5)
    xs = [torch.randn(3, 16, generator=g).to(dev) for _ in range(M)]
    ys = [torch.randn(3, 16, generator=g).to(dev) for _ in range(M)]
    loss_fn = lambda out, y: torch.nn.functional.mse_loss(out, y)
    engine = PipeEngine(pm, mesh, loss_fn, plan)
    for it in range(2):
        loss, _ = engine(xs, ys)
        ref_loss = sum(loss_fn(ref(x), y) / M for x, y in zip(xs, ys))
        ref_loss.backward()
        if engine.is_last_rank:
            torch.testing.assert_close(loss, ref_loss.detach(), rtol=1e-5, atol=1e-6)
    prog = engine.graph_program(xs)
    V = plan.virtual_chunks
    sends = [n for n in prog.comm_nodes() if "send" in n.name]
    recvs = [n for n in prog.comm_nodes() if "recv" in n.name]
    # one receive per (micro-batch, chunk fed by another rank), one send per (micro-batch, chunk feeding another rank)
    topo = prog.emitter.gen_pp_collective_topo()
    assert len(recvs) == M * sum(1 for s in topo["fwd_recv_srcs"].values() if s is not None)
    assert len(sends) == M * sum(1 for d in topo["fwd_send_dsts"].values() if d is not None)
    assert topo["bwd_recv_srcs"] == topo["fwd_send_dsts"] and (("p2p_recv" in prog.gm.code) == bool(recvs))
    assert not [n for n, _ in prog.gm.named_parameters() if n.startswith("like_")]  # anchors are buffers, invisible to optimizers
    ref_params = dict(ref.named_parameters())
    checked = 0
    for c in range(pm.num_chunks):
        stage = pm.chunk(c)
        for n, p in stage.named_parameters():
            _, idx, rest = n.split(".", 2)
            torch.testing.assert_close(p.grad, ref_params[f"{stage.names[int(idx)]}.{rest}"].grad, rtol=1e-4, atol=1e-6)
            checked += 1
    assert checked > 0
    # forward-only evaluation through the same program
    _, outs = engine.forward_backward(xs, None, forward_only=True)
    if engine.is_last_rank:
        torch.testing.assert_close(outs[0], ref(xs[0]), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("sched", ["SIMPLE_1F1B", "INTERLEAVED_1F1B", "ZERO_BUBBLE_V"])
def test_pipeline_graph_mode_emitter(sched):
    run_distributed(_pp_graph_mode, 4, sched)


class _Emb(nn.Module):
    def __init__(self, v, h):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(v, h) * 0.1)

    def forward(self, ids):
        return torch.nn.functional.embedding(ids, self.weight)


class _Head(nn.Module):
    def __init__(self, v, h):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(v, h) * 0.1)

    def forward(self, x):
        return x @ self.weight.t()


def _pp_shared(rank, world):
    """Tied embedding / lm-head living on the first and last stage: values are synchronised once, gradients are summed over
    the owning ranks every step (legacy ``test_shared_params.py`` / ``pipe_stage.py:200-247``)."""
    from vescale_b200 import init_device_mesh
    from vescale_b200.parallel.pipe import PipeEngine, PipelineParallelPlan, PipelineScheduleType, construct_pipeline_stage

    dev = device_type()
    V, H = 11, 16
    torch.manual_seed(0)
    mods = [_Emb(V, H)] + [Blk(H) for _ in range(6)] + [_Head(V, H)]
    mods[-1].weight.data.copy_(mods[0].weight.data)
    ref = nn.Sequential(*mods).to(dev)
    ref[7].weight = ref[0].weight  # really tied in the golden model
    model = copy.deepcopy(nn.Sequential(*mods)).to(dev)
    mesh = init_device_mesh(dev, (world,), mesh_dim_names=("PP",))
    plan = PipelineParallelPlan(num_stages=world, schedule_type=PipelineScheduleType.SIMPLE_1F1B, shared_modules=[["0.weight", "7.weight"]])
    pm = construct_pipeline_stage(model, plan, mesh)
    assert (len(pm.shared_groups) == 1) == (rank in (0, world - 1))
    engine = PipeEngine(pm, mesh, lambda out, y: torch.nn.functional.cross_entropy(out, y), plan)
    engine.sync_shared_params(share_params=True)
    M = 4
    g = torch.Generator().manual_seed(5)
    xs = [torch.randint(0, V, (3,), generator=g).to(dev) for _ in range(M)]
    ys = [torch.randint(0, V, (3,), generator=g).to(dev) for _ in range(M)]
    loss, _ = engine(xs, ys)
    engine.sync_shared_params(share_params=False)  # sum the two partial gradients of the tied weight
    ref_loss = sum(torch.nn.functional.cross_entropy(ref(x), y) / M for x, y in zip(xs, ys))
    ref_loss.backward()
    if engine.is_last_rank:
        torch.testing.assert_close(loss, ref_loss.detach(), rtol=1e-5, atol=1e-6)
    if rank in (0, world - 1):
        (params, _grp), = pm.shared_groups
        assert len(params) == 1
        torch.testing.assert_close(params[0].grad, ref[0].weight.grad, rtol=1e-4, atol=1e-6)


def test_pipeline_shared_params():
    run_distributed(_pp_shared, 4)
