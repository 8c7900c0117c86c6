"""Emulator order/bit-exactness, auto-plan, VeDeviceMesh, debug logger, profiler, vescale alias (CPU)."""
import json
import os

import torch
import torch.nn as nn

from common import device_type, run_distributed


def test_emulator_ring_order_and_dtensor_api():
    from vescale_b200 import DeviceMesh, Partial, Replicate, Shard
    from vescale_b200.emulator import EmulatorProcessGroup, distribute_tensor, full_tensor, redistribute_dtensor, ring_all_reduce, tree_all_reduce

    torch.manual_seed(0)
    n = 4
    xs = [torch.randn(37) * 10 ** (i - 2) for i in range(n)]
    out = ring_all_reduce(xs)
    # chunk c is accumulated starting at rank c+1 and finishing at rank c
    from vescale_b200.emulator import nccl_chunking

    (_, off, cs), = nccl_chunking(37, n)
    for c in range(n):
        lo, hi = off + c * cs, min(off + (c + 1) * cs, 37)
        acc = xs[(c + 1) % n][lo:hi].clone()
        for j in range(2, n + 1):
            acc = acc + xs[(c + j) % n][lo:hi]
        assert torch.equal(out[0][lo:hi], acc)
    assert all(torch.equal(out[0], o) for o in out)
    # different association than a naive sum (that is the point of the emulator)
    torch.testing.assert_close(out[0], sum(xs), rtol=1e-5, atol=1e-3)
    t = tree_all_reduce(xs)
    assert torch.equal(t[0], (xs[0] + (xs[1] + xs[3])) + xs[2])  # a node's own buffer is source 0, then its children in order (NCCL reduceCopy)
    mesh = DeviceMesh("meta", torch.arange(4).reshape(2, 2), mesh_dim_names=("dp", "tp"), _rank=0)
    full = torch.randn(8, 6)
    shards = distribute_tensor(full, mesh, [Shard(0), Shard(1)])
    assert shards[3].shape == (4, 3)
    assert torch.equal(full_tensor(shards, (8, 6), mesh, [Shard(0), Shard(1)]), full)
    rs = redistribute_dtensor(shards, (8, 6), mesh, [Shard(0), Shard(1)], [Replicate(), Shard(0)])
    assert torch.equal(rs[1], full[4:8])
    parts = [torch.randn(8, 6) for _ in range(4)]
    tot = full_tensor(parts, (8, 6), mesh, [Partial(), Partial()])
    torch.testing.assert_close(tot, sum(parts), rtol=1e-5, atol=1e-5)


class GPTBlock(nn.Module):
    def __init__(self, h=32):
        super().__init__()
        self.ln_1 = nn.LayerNorm(h)
        self.c_fc = nn.Linear(h, 4 * h)
        self.c_proj = nn.Linear(4 * h, h)

    def forward(self, x):
        return x + self.c_proj(torch.nn.functional.gelu(self.c_fc(self.ln_1(x))))


def _auto_plan(rank, world):
    from vescale_b200 import Replicate, Shard
    from vescale_b200.devicemesh_api import VESCALE_DEVICE_MESH
    from vescale_b200.debug import DebugLogger, set_vescale_debug_mode
    from vescale_b200.parallel.dmp import auto_parallelize_module

    dev = device_type()
    mesh = VESCALE_DEVICE_MESH.init_device_mesh(dev, (2, 2), mesh_dim_names=("DP", "TP"))
    assert VESCALE_DEVICE_MESH.get_strategy_size("TP") == 2 and VESCALE_DEVICE_MESH.get_tensor_parallel_rank() == rank % 2
    assert VESCALE_DEVICE_MESH.is_first_stage() and VESCALE_DEVICE_MESH.get_data_parallel_rank() == rank // 2
    torch.manual_seed(0)
    ref = GPTBlock().to(dev)
    import copy

    model = copy.deepcopy(ref)
    set_vescale_debug_mode(True, rank_to_print=(0,), logger=None)
    DebugLogger.records.clear()
    auto_parallelize_module(model, VESCALE_DEVICE_MESH["TP"], "MEGATRON", plan_override={"forward": {r"input": [[Replicate()]], r"c_proj\.output": [[Replicate()]]}})
    assert model.c_fc.weight.placements == (Shard(0),) and model.c_proj.weight.placements == (Shard(1),)
    x = torch.randn(2, 4, 32, generator=torch.Generator().manual_seed(1)).to(dev)
    y = model(x)
    torch.testing.assert_close(y.full_tensor(), ref(x), rtol=1e-4, atol=1e-5)
    summary = DebugLogger.summary()
    set_vescale_debug_mode(False)
    if rank == 0:
        assert any("[comm]" in r for r in DebugLogger.records) and any("[op]" in r for r in DebugLogger.records)
        # ops are attributed to the module whose forward dispatched them, with operand / output specs spelled out
        assert any("[op] Linear forward()" in r and "in  = [" in r and "Shard(" in r for r in DebugLogger.records), DebugLogger.records[:5]
        # collectives name the user line that injected them (this file), and the summary aggregates bytes per site
        assert any("[comm]" in r and "test_aux_components.py:" in r for r in DebugLogger.records)
        assert "communication by call site" in summary and "MiB" in summary and "Linear.forward" in summary
    else:
        assert not DebugLogger.records  # rank filter
    DebugLogger.reset()
    # the manual / decorator forms and the environment switch (with a rank list after ':')
    import os

    os.environ["VESCALE_DEBUG_MODE"] = "1:-1"
    from vescale_b200.debug import update_vescale_debug_mode_from_env

    assert update_vescale_debug_mode_from_env() and DebugLogger.IS_DEBUG_MODE and DebugLogger.ranks == (-1,)

    @DebugLogger.log_communication_decorator()
    def my_all_reduce(t, tag=None):
        return t

    my_all_reduce(torch.ones(4, 2), tag="x")
    assert any("my_all_reduce(float32[4, 2], tag=x)" in r for r in DebugLogger.records), DebugLogger.records
    os.environ["VESCALE_DEBUG_MODE"] = "0"
    assert not update_vescale_debug_mode_from_env() and not DebugLogger.IS_DEBUG_MODE
    n = len(DebugLogger.records)
    (model(x)).full_tensor()
    assert len(DebugLogger.records) == n  # hooks are gone
    DebugLogger.reset()


def test_auto_plan_devicemesh_debuglog():
    run_distributed(_auto_plan, 4)


def test_profiler_chrome_trace(tmp_path):
    import vescale_b200.profiler as nd

    h = nd.ChromeTraceNDHandler(str(tmp_path))
    p = nd.ParserNDHandler()
    nd.init_ndtimers(0, 1, [h, p])

    @nd.ndtimer("decorated")
    def f():
        return sum(range(100))

    for step in range(2):
        with nd.ndtimeit(nd.predefined.FORWARD_COMPUTE, microbatch=step):
            f()
        nd.inc_step()
        nd.flush(asynchronous=True)
    nd.wait()
    ev = json.load(open(tmp_path / "ndtimeline_rank0.json"))["traceEvents"]
    slices = [e for e in ev if e["ph"] == "X"]
    assert {e["name"] for e in slices} == {"decorated", "forward-compute"} and all(e["dur"] >= 0 for e in slices)
    assert [e for e in ev if e["ph"] == "M"][0]["args"]["name"].startswith("rank 0")
    assert p.summary()["decorated"]["count"] == 2
    # typed records of the parser handler: ordered, on the global clock, tags preserved
    assert len(p.records) == 4 and all(isinstance(r, nd.DeviceTimerStreamRecord) for r in p.records) and p.records[0].end_ts >= p.records[0].ts
    assert {r.tags.get("microbatch") for r in p.records if r.metric == "forward-compute"} == {0, 1}


def test_chrome_trace_events_flows_and_host_timeline(tmp_path):
    """Typed trace events, stable thread ids, send -> recv flow arrows across ranks, merged host timeline (legacy
    ``handlers/chrome_trace_event.py``, ``local_timeline_handler.py``)."""
    import vescale_b200.profiler as nd
    from vescale_b200.profiler.chrome_trace_event import (BeginEvent, CombinedEvents, CompleteEvent, CounterEvent, DummyEvent, EndEvent, FlowEvent, ThreadMetadataEvent,
                                                          build_thread_index_table, link_p2p_flows)

    doc = CombinedEvents([CompleteEvent(name="k", ts=1.0, dur=2.0, pid=0, tid=1, args={"a": 1}), BeginEvent(name="b", ts=1.0), EndEvent(name="b", ts=2.0),
                          CounterEvent(name="bytes", ts=1.5, args={"in_flight": 3}), DummyEvent(), ThreadMetadataEvent(pid=0, tid=1, args={"name": "comm"})])
    d = json.loads(doc.to_json())
    assert [e["ph"] for e in d["traceEvents"]] == ["X", "B", "E", "C", "M"] and d["traceEvents"][0]["dur"] == 2.0 and "args" not in d["traceEvents"][1]
    table = build_thread_index_table([(0, 77), (0, 0), (1, 5), (0, 77), (1, 0)])
    assert table[(0, 0)] == 0 and table[(0, 77)] == 1 and table[(1, 0)] == 0 and table[(1, 5)] == 1
    # two ranks of a pipeline: rank 0 sends forward twice to rank 1, rank 1 sends one gradient back
    def rec(metric, start, peer, **tags):
        return {"metric": metric, "start_us": start, "duration_us": 10.0, "tags": {"peer": peer, **tags}, "stream": 0}
    h = nd.LocalTimelineNDHandler(str(tmp_path / "host.json"))
    h([rec("send-forward", 100.0, 1, microbatch=0), rec("send-forward", 200.0, 1, microbatch=1), rec("recv-backward", 400.0, 1, microbatch=0)], 0, 0)
    h([rec("recv-forward", 120.0, 0, microbatch=0), rec("recv-forward", 230.0, 0, microbatch=1), rec("send-backward", 380.0, 0, microbatch=0),
       {"metric": "grad-reduce-scatter", "start_us": 300.0, "duration_us": 50.0, "tags": {}, "stream": 9}], 1, 0)
    ev = json.load(open(tmp_path / "host.json"))["traceEvents"]
    flows = [e for e in ev if e["ph"] in ("s", "f")]
    assert len(flows) == 6 and {e["id"] for e in flows} == {1, 2, 3}
    starts = {e["id"]: e for e in flows if e["ph"] == "s"}
    ends = {e["id"]: e for e in flows if e["ph"] == "f"}
    assert all(ends[i]["bp"] == "e" and ends[i]["ts"] >= starts[i]["ts"] and ends[i]["pid"] != starts[i]["pid"] for i in starts)
    names = {(e["pid"], e["tid"]): e["args"]["name"] for e in ev if e["ph"] == "M" and e["name"] == "thread_name"}
    assert names[(1, 0)] == "compute" and names[(1, 1)] == "stream 9"
    # per-rank files merged after the fact get the same arrows
    for rk, recs in ((0, [rec("send-forward", 100.0, 1)]), (1, [rec("recv-forward", 130.0, 0)])):
        nd.ChromeTraceNDHandler(str(tmp_path), prefix="m")(recs, rk, 0)
    n = nd.ChromeTraceNDHandler.merge([str(tmp_path / "m_rank0.json"), str(tmp_path / "m_rank1.json")], str(tmp_path / "merged.json"))
    merged = json.load(open(tmp_path / "merged.json"))["traceEvents"]
    assert n == len(merged) and sum(1 for e in merged if e["ph"] in ("s", "f")) == 2
    assert isinstance(link_p2p_flows([])[0:0], list) and FlowEvent(ph="f", bp="e", id=3).to_dict()["bp"] == "e"


def test_vescale_alias_package():
    import vescale
    from vescale.dtensor import DTensor, RaggedShard, distribute_tensor  # noqa: F401
    from vescale.dmodule.api import parallelize_module  # noqa: F401
    import vescale.checkpoint as ck

    assert vescale.DTensor is DTensor and hasattr(ck, "save") and hasattr(vescale, "init_device_mesh")


def _make_stream_handlers():
    import os

    from vescale_b200.profiler import LocalTimelineNDHandler

    return [LocalTimelineNDHandler(os.environ["NDTL_TEST_OUT"])]


def test_profiler_socket_streamer(tmp_path):
    """Binary framing round-trip + records of two 'ranks' streamed over a unix socket into one per-host timeline."""
    import json
    import os

    from vescale_b200.profiler import NDtimelineStreamer, SockNDHandler, decode_frames, encode_frame

    recs = [{"metric": "forward-compute", "start_us": 10.0, "duration_us": 5.0, "stream": 7, "tags": {"mb": 1}}]
    buf = bytearray(encode_frame(recs, 3, 42) + encode_frame(recs, 4, 43)[:10])
    got = list(decode_frames(buf))
    assert got == [(0, 3, 42, recs)] and len(buf) == 10  # partial second frame stays buffered
    out = str(tmp_path / "host_timeline.json")
    os.environ["NDTL_TEST_OUT"] = out
    sock = str(tmp_path / "ndtl.sock")
    streamer = NDtimelineStreamer.start(sock, _make_stream_handlers, expected_clients=2)
    hs = [SockNDHandler(sock) for _ in range(2)]
    for step in range(3):
        for rank, h in enumerate(hs):
            h(recs, rank, step)
    for rank, h in enumerate(hs):
        h.close(rank)
    streamer.join(30)
    ev = json.load(open(out))["traceEvents"]
    spans = [e for e in ev if e["ph"] == "X"]
    assert len(spans) == 6 and {e["pid"] for e in spans} == {0, 1} and all(e["tid"] == 0 for e in spans)  # small stable thread ids ...
    assert {e["args"]["name"] for e in ev if e["ph"] == "M" and e["name"] == "thread_name"} == {"stream 7"}  # ... named after the CUDA stream
    # the receiving side on its own: frames off a byte stream through a bare recv(), carry-over between calls, validation
    import io

    import pytest

    from vescale_b200.profiler.binary_protocol import dumps, loads, read_or_recv, recv_and_validate
    from vescale_b200.profiler.exceptions import ProtocolValidationError

    stream = io.BytesIO(encode_frame(recs, 3, 42) + encode_frame([], 5, 1))
    recv = lambda n: stream.read(min(n, 7))  # noqa: E731  a peer that trickles 7 bytes at a time
    carry = bytearray()
    kind, rank, step, payload = recv_and_validate(recv, carry)
    assert (kind, rank, step, loads(payload)) == (0, 3, 42, recs)
    assert recv_and_validate(recv, carry)[1:3] == (5, 1) and loads(dumps({"a": [1, 2]})) == {"a": [1, 2]}
    with pytest.raises(EOFError):
        recv_and_validate(recv, carry)
    with pytest.raises(ProtocolValidationError, match="magic"):
        recv_and_validate(io.BytesIO(b"XXXX" + bytes(20)).read, bytearray())
    with pytest.raises(BrokenPipeError):
        recv_and_validate(io.BytesIO(encode_frame(recs, 0, 0)[:-3]).read, bytearray())
    pre = bytearray(b"abcdef")
    assert read_or_recv(4, None, pre) == b"abcd" and pre == bytearray(b"ef")


def _emulator_vs_real(rank, world):
    """The emulator's global-view collectives against the real process group on the same inputs, side by side
    (``legacy/test/emulator/test_distributed.py:72-101`` strategy).  Integer-valued floats make the sums order-independent,
    so the comparison is bitwise on gloo as well as NCCL."""
    import torch.distributed as dist

    from vescale_b200.emulator import EmulatorProcessGroup

    dev = device_type()
    pg = EmulatorProcessGroup(world, algo="ring")
    gens = [torch.Generator().manual_seed(7 + r) for r in range(world)]
    all_inputs = [torch.randint(-50, 50, (4 * world, 5), generator=g).float() for g in gens]  # every rank can build every input
    mine = all_inputs[rank].clone().to(dev)
    # all_reduce
    emu = pg.all_reduce(all_inputs)
    real = mine.clone()
    dist.all_reduce(real)
    assert torch.equal(real.cpu(), emu[rank])
    for algo in ("tree",):
        assert torch.equal(EmulatorProcessGroup(world, algo=algo).all_reduce(all_inputs)[rank], emu[rank])
    # all_gather
    emu = pg.all_gather(all_inputs)
    outs = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(outs, mine)
    assert torch.equal(torch.cat(outs).cpu(), emu[rank].reshape(-1, 5))
    # reduce_scatter (gloo has no native one: all_reduce + slice is the definition)
    emu = pg.reduce_scatter(all_inputs)
    want = sum(all_inputs).chunk(world, 0)[rank]
    assert torch.equal(emu[rank].reshape(want.shape), want)
    # all_to_all
    pieces = [list(t.chunk(world, 0)) for t in all_inputs]
    emu = pg.all_to_all(pieces)
    want = [all_inputs[src].chunk(world, 0)[rank] for src in range(world)]
    assert all(torch.equal(a, b) for a, b in zip(emu[rank], want))
    # broadcast
    emu = pg.broadcast(all_inputs, src=2)
    b = mine.clone()
    dist.broadcast(b, src=2)
    assert torch.equal(b.cpu(), emu[rank])


def test_emulator_matches_real_collectives():
    run_distributed(_emulator_vs_real, 4)


def test_emulator_topology_tuning_primitives():
    """Double binary tree structure, tuning-model monotonicity, chunk sizes, transition primitives and instrumentation."""
    from vescale_b200 import DeviceMesh
    from vescale_b200.emulator import (P2R, P2S, R2P, R2S, S2R, S2S, EmulatorInstrumentation, EmulatorProcessGroup, btree, calculate_chunk_size, double_tree,
                                       double_tree_all_reduce, parse_graph_dump, select_algorithm)

    for n in (2, 3, 4, 7, 8, 16):
        t0, t1 = double_tree(n).trees
        for t in (t0, t1):
            assert sorted(t.parent) == list(range(n)) and sum(1 for p in t.parent.values() if p == -1) == 1
            assert all(len(c) <= 2 for c in t.children.values())
            seen, stack = set(), [t.root]
            while stack:  # connected, acyclic
                r = stack.pop()
                assert r not in seen
                seen.add(r)
                stack += t.children[r]
            assert seen == set(range(n))
        if n > 2:  # leaves of one tree are inner nodes of the other (that is the point of the double tree)
            leaves0 = {r for r in range(n) if not t0.children[r]}
            leaves1 = {r for r in range(n) if not t1.children[r]}
            assert len(leaves0 & leaves1) <= 1
    xs = [torch.randn(50) * 10 ** (i - 3) for i in range(8)]
    out = double_tree_all_reduce(xs)
    torch.testing.assert_close(out[0], sum(xs), rtol=1e-5, atol=1e-3)
    assert all(torch.equal(out[0], o) for o in out)
    # tuning: small messages pick a low-latency protocol, large ones 'simple'; predicted time grows with size
    small, big = select_algorithm("all_reduce", 1024, 8), select_algorithm("all_reduce", 256 << 20, 8)
    assert small.proto in ("ll", "ll128") and big.proto == "simple" and big.time_us > small.time_us
    assert select_algorithm("all_gather", 64 << 20, 8).algo == "ring"
    assert calculate_chunk_size(256 << 20, 8, 16, "simple", "ring") >= calculate_chunk_size(64 << 10, 8, 16, "simple", "ring") >= 512
    assert calculate_chunk_size(1 << 20, 8, 4, "ll", "ring") % 16 == 0
    auto = EmulatorProcessGroup(8, algo="auto").all_reduce(xs)
    torch.testing.assert_close(auto[0], sum(xs), rtol=1e-5, atol=1e-3)
    xml = '<graphs version="1"><graph id="0" pattern="4" nchannels="2"><channel><net dev="0"/><gpu dev="0"/><gpu dev="2"/><gpu dev="1"/></channel><channel><gpu dev="1"/><gpu dev="0"/><gpu dev="2"/></channel></graph></graphs>'
    assert parse_graph_dump(xml)["ring"] == [[0, 2, 1], [1, 0, 2]]
    # primitives on a 2 x 2 mesh
    mesh = DeviceMesh("meta", torch.arange(4).reshape(2, 2), mesh_dim_names=("a", "b"), _rank=0)
    full = torch.arange(48.0).reshape(8, 6)
    rep = [full.clone() for _ in range(4)]
    sh = R2S(rep, mesh, 1, 0)
    assert torch.equal(sh[1], full[4:]) and torch.equal(sh[2], full[:4])
    assert all(torch.equal(t, full) for t in S2R(sh, mesh, 1, 0))
    s2 = S2S(sh, mesh, 1, 0, 1)
    assert torch.equal(s2[0], full[:, :3]) and torch.equal(s2[3], full[:, 3:])
    part = R2P(rep, mesh, 0)
    assert torch.equal(part[0], full) and torch.equal(part[2], torch.zeros_like(full))
    assert all(torch.equal(t, full) for t in P2R(part, mesh, 0))
    ps = P2S(part, mesh, 0, 1)
    assert torch.equal(ps[0], full[:, :3]) and torch.equal(ps[2], full[:, 3:])
    # instrumentation: ordinary torch code on per-rank lists
    with EmulatorInstrumentation(4, [(torch, "add"), (torch, "relu")]):
        y = torch.relu(torch.add(sh, 1.0))
    assert isinstance(y, list) and torch.equal(y[1], torch.relu(full[4:] + 1))
    assert torch.equal(torch.add(full, 1.0), full + 1)  # restored


def test_reference_import_paths_and_extension_points(tmp_path):
    """The module paths the reference's own examples / tests import from resolve to the SAME objects as the implementing
    modules; the class-name keyed auto-plan registry, the torch.distributed-shaped emulator front end, the emulated device
    mesh and the declared-timer metadata of ndtimeline behave as documented."""
    import importlib
    import warnings

    import torch.nn as nn

    import vescale
    import vescale_b200

    pairs = {
        "vescale.dtensor.api": ["from_local", "to_local", "distribute_tensor", "redistribute_dtensor", "vescale_all_gather", "vescale_all_reduce", "vescale_reduce_scatter", "normalize_placements"],
        "vescale.dtensor.placement_types": ["Shard", "Replicate", "Partial", "InterleavedShard", "RaggedShard", "DTensorSpec", "TensorMeta"],
        "vescale.dtensor.device_mesh": ["DeviceMesh", "init_device_mesh", "mesh_resources"],
        "vescale.dtensor.dtensor": ["DTensor", "make_dtensor"],
        "vescale.dtensor.random": ["manual_seed", "init_vescale_rng_tracker", "is_rng_supported_mesh", "OffsetBasedRNGTracker", "ThreadBasedRNGTracker"],
        "vescale.dtensor.vescale_utils": ["get_ragged_shard", "best_effort_reshape", "retrieve_flattened_index_before_ragged_shard", "cvt_inclusive_to_exclusive"],
        "vescale.dmodule.api": ["parallelize_module", "is_dmodule", "PlacementsInterface"],
        "vescale.ddp.distributed_data_parallel": ["DistributedDataParallel"],
        "vescale.optim.distributed_optimizer": ["DistributedOptimizer"],
        "vescale.optim.base_optimizer": ["BasicOptimizer", "BasicOptimizerHook"],
        "vescale.initialize.deferred_init": ["deferred_init", "is_deferred", "materialize_dtensor", "materialize_dparameter"],
        "vescale.plan": ["PipelineParallelPlan", "PipelineScheduleType", "PipelineSplitMethodType", "ModeType", "TracerType", "PipelineP2PSpec"],
        "vescale.plan.spec": ["PipelineScheduleType", "PipelineSplitMethodType"],
        "vescale.pipe": ["PipeModule", "construct_stage_modules", "construct_pipeline_stage", "build_shared_module_group", "build_stage_module_and_dependency", "PipeParser",
                         "parse_model_graph", "split_pipeline_point", "construct_pipeline_split_graph", "ScheduleEngine", "validate_pipeline_schedule"],
        "vescale.pipe.pipe_stage": ["PipeModule", "construct_pipeline_stage"],
        "vescale.pipe._schedules.instruction_base": ["StageDeps", "register_instruction", "Shape"],
        "vescale.pipe._schedules.pipedream_flush": ["OneFOneBInstrcutionGenerator"],
        "vescale.engine": ["PipeEngine"],
        "vescale.moe": ["parallelize_experts", "is_experts_parallized", "MoEOptimizer", "ExpertsAllocator", "TokenDispatcher"],
        "vescale.dmp": ["auto_parallelize_module", "set_plan_overriding_policy", "get_plan_overriding_policy"],
        "vescale.dmp.policies": ["REGISTRY"],
        "vescale.checkpoint": ["save", "load", "CheckpointState"],
        "vescale.devicemesh_api": ["VESCALE_DEVICE_MESH"],
        "vescale.model.patch": ["get_all_model_patch"],
        "vescale.emulator.distributed": ["ProcessGroup", "init_process_group", "new_group", "get_rank", "set_rank", "destroy_process_group"],
        "vescale.emulator.device_mesh": ["DeviceMesh", "init_device_mesh"],
        "vescale.emulator.reduce_kernel": ["ReduceOp"],
        "vescale.emulator.utils": ["flatten_tensors", "restore_tensors"],
        "vescale.emulator.mesh_collectives": ["mesh_all_gather", "mesh_all_reduce", "mesh_reduce_scatter", "mesh_all_to_all", "mesh_broadcast", "mesh_scatter"],
        "vescale.ndtimeline": ["init_ndtimers", "flush", "wait", "inc_step", "set_global_step", "ndtimeit", "ndtimer", "WorldInfo", "DeviceTimerMeta", "NDTimerManagerSingleton",
                               "CudaEventPool", "DefaultEventPool", "get_nccl_coll_stream", "get_nccl_p2p_stream", "encode_package", "serialize_to_package", "SOCK_PATH"],
        "vescale.ndtimeline.world_info": ["WorldInfo", "TopoInfo", "TrainingInfo"],
        "vescale.ndtimeline.handlers": ["ChromeTraceNDHandler", "LocalRawNDHandler", "LocalTimelineNDHandler", "LoggingNDHandler", "ParserNDHandler", "DoNothingNDHandler", "SockNDHandler"],
        "vescale.debug": ["DebugLogger"],
        "vescale": ["deprecated_function", "switch_dtensor_for_torch_export", "parallelize_module", "DistributedDataParallel", "DistributedOptimizer", "deferred_init"],
    }
    for mod, names in pairs.items():
        m = importlib.import_module(mod)
        for n in names:
            assert hasattr(m, n), f"{mod}.{n}"
    assert importlib.import_module("vescale.dtensor.api") is importlib.import_module("vescale_b200.dtensor.api")
    assert importlib.import_module("vescale.emulator.distributed") is vescale_b200.emulator.distributed

    # --- class-name keyed plan providers take precedence over the built-in policy
    from vescale.dmp.policies import REGISTRY
    from vescale_b200 import Replicate, Shard
    from vescale_b200.parallel.dmp.registry import get_policy

    register = REGISTRY.provide_register_for_policy("MEGATRON")

    class MyFancyFFN(nn.Module):
        def __init__(self):
            super().__init__()
            self.a, self.b = nn.Linear(4, 8), nn.Linear(8, 4)

    @register("FancyFFN")
    def provider(fqn, module):
        return {"a.weight": [Shard(0)], "b.weight": [Shard(1)]}, {"a.input": [[Replicate()]]}

    plan = get_policy("MEGATRON").provide("blocks.3.ffn", MyFancyFFN(), None)
    assert plan["parameter"] == {r"blocks\.3\.ffn\.a\.weight": [Shard(0)], r"blocks\.3\.ffn\.b\.weight": [Shard(1)]} and list(plan["forward"]) == [r"blocks\.3\.ffn\.a\.input"]
    assert REGISTRY.has_module("fancyffn") and REGISTRY.has_policy("megatron")
    try:
        register("FancyFFN")(provider)
        raise AssertionError("a second provider for the same (class, policy) must be rejected")
    except ValueError:
        pass

    # --- emulator front end: in-place list semantics, NCCL-ordered sums, emulated mesh
    import vescale.emulator.distributed as edist
    from vescale.emulator.device_mesh import init_device_mesh as emu_mesh

    mesh = emu_mesh("cpu", (2, 2), mesh_dim_names=("dp", "tp"))
    assert [g.ranks for g in mesh.get_dim_groups("tp")] == [[0, 1], [2, 3]] and edist.get_world_size() == 4
    xs = [torch.randn(33) for _ in range(4)]
    want = [xs[0] + xs[1], xs[0] + xs[1], xs[2] + xs[3], xs[2] + xs[3]]
    mesh.all_reduce(xs, "tp")
    for a, b in zip(xs, want):
        torch.testing.assert_close(a, b)
    pg = edist.new_group([0, 1, 2, 3])
    outs, tl = [None] * 4, [[torch.full((3,), float(10 * i + j)) for j in range(4)] for i in range(4)]
    pg.reduce_scatter(outs, tl)
    assert outs[1][0].item() == 1 + 11 + 21 + 31
    recv = [None] * 4
    pg.all_to_all(recv, tl)
    assert recv[1][2][0].item() == 21
    edist.set_rank(3)
    assert edist.get_rank(pg) == 3 and edist.get_group_rank(pg, 2) == 2
    edist.destroy_process_group()
    assert not edist.is_initialized()

    # --- ndtimeline: declared timers (level / legal tags / enabled), world info reaches the handlers, singleton accessor
    import vescale.ndtimeline as nd

    p = nd.ParserNDHandler()
    wi = nd.WorldInfo(rank=0, local_rank=0, tp_rank=1, tp_size=2, world_size=2, run_id=7, cluster="b200")
    mgr = nd.init_ndtimers(0, 1, [p], world_info=wi, metas=[nd.DeviceTimerMeta(name="declared", legal_tags=["mb"], level=nd.NDMetricLevel.INFO, common_extra={"k": 1}),
                                                          nd.DeviceTimerMeta(name="off", enabled=False, level=nd.NDMetricLevel.INFO)])
    assert nd.NDTimerManagerSingleton() is mgr and p.world_info is wi and wi["tp_rank"] == 1 and wi["cluster"] == "b200" and wi["run_id"] == 7
    with nd.ndtimeit("declared", mb=3):
        pass
    with nd.ndtimeit("off"):
        pass
    try:
        with nd.ndtimeit("declared", bogus=1):
            pass
        raise AssertionError("undeclared tag must be rejected")
    except ValueError:
        pass
    nd.flush(asynchronous=False)
    s = p.summary()
    assert s["declared"]["count"] == 1 and "off" not in s
    try:
        nd.TopoInfo(rank=-1)
        raise AssertionError
    except ValueError:
        pass
    pkg = nd.serialize_to_package([{"a": 1}], rank=2, step=5)
    assert list(nd.decode_frames(bytearray(pkg))) == [(0, 2, 5, [{"a": 1}])]

    @vescale.deprecated_function
    def old():
        return 1

    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        assert old() == 1 and len(w) == 1


def test_hierarchical_double_tree_and_rotating_raw_handler(tmp_path, monkeypatch):
    """NCCL's node-chain + inter-node double tree (legacy ``test/emulator/test_topo.py``) and the rotating raw-record file
    (``test/ndtimeline/test_local_raw_handler.py``)."""
    import os

    from vescale_b200.emulator.topo import DoubleTree

    table = [[n * 8 + i for i in range(8)] for n in range(4)]
    ranks = [0, 1, 2, 3, 8, 9, 10, 11, 16, 17, 18, 19, 24, 25, 26, 27]
    dt = DoubleTree(table, ranks, {r: i for i, r in enumerate(ranks)})
    t0, t1 = dt.tree
    assert str(t0[1]) == "[Rank 1] up: 0, down: [2, -1, 8].\n" and str(t0[9]) == "[Rank 9] up: 8, down: [10, 4, 12].\n"
    assert str(t1[5]) == "[Rank 5] up: 4, down: [6, 8, 0].\n" and str(t1[12]) == "[Rank 12] up: -1, down: [13, -1, -1].\n"
    for t in dt.tree:  # exactly one root; every other rank is reachable from it; chain links are mutual
        assert sum(n.up == -1 for n in t) == 1
        for n in t:
            for d in n.down:
                assert d == -1 or t[d].up == n.rank
    for nn_ in (3, 5):  # odd node counts use the shifted second tree; roots of the two trees sit on different nodes
        d = DoubleTree([[n * 2, n * 2 + 1] for n in range(nn_)], list(range(2 * nn_)))
        roots = [next(n.rank for n in t if n.up == -1) // 2 for t in d.tree]
        assert roots[0] != roots[1]

    import vescale_b200.profiler as prof
    from vescale_b200.profiler.handlers import LocalRawNDHandler, NDRecord

    monkeypatch.setattr(prof, "LOCAL_LOGGING_PATH", str(tmp_path))
    h = LocalRawNDHandler(run_id=7, chunk_sz=10, backup_cnt=3)
    for _ in range(6):
        out = h("m", 1.0, [1.0], [0.5], [{}], range(0, 1), None, {})
    assert isinstance(out[0], NDRecord) and out[0].metric == "m"
    base = os.path.join(str(tmp_path), "timeline_run7_raw.log")
    assert os.path.exists(base) and os.path.exists(base + ".3") and not os.path.exists(base + ".4")


def test_deferred_tensor_factories_and_torchdistx_names():
    """``deferred_init(torch.empty / full / randn, shape)`` yields one deferred tensor (legacy ``test/initialize/test_defer_init.py``);
    the ``torchdistx`` entry points the reference's users import forward to the same machinery."""
    import torch
    from torchdistx.deferred_init import deferred_init, is_deferred, materialize_module, materialize_tensor
    from torchdistx.fake import is_fake

    with torch.device("meta"):
        t = deferred_init(torch.empty, (4, 16, 16))
    assert is_deferred(t) and is_fake(t) and t.device == torch.device("meta") and tuple(t.shape) == (4, 16, 16)
    full = materialize_tensor(deferred_init(torch.full, (3, 2), 2.5, device="cpu"))
    assert full.device.type == "cpu" and torch.equal(full, torch.full((3, 2), 2.5))
    assert torch.equal(materialize_tensor(deferred_init(torch.ones, (5,))), torch.ones(5))
    r = materialize_tensor(deferred_init(torch.randn, (64, 64)))
    assert abs(float(r.mean())) < 0.1 and 0.8 < float(r.std()) < 1.2
    m = deferred_init(torch.nn.Linear, 8, 8)
    assert is_deferred(m) and all(is_fake(p) for p in m.parameters())
    materialize_module(m)
    assert not is_deferred(m) and float(m.weight.abs().sum()) > 0


def test_emulator_mesh_collectives_reference_call_forms():
    """``vescale.emulator.mesh_collectives`` argument order (legacy ``emulator/mesh_collectives.py``) over a 2 × 2 emulated mesh."""
    import torch

    import vescale  # noqa: F401
    import vescale.emulator.distributed as ed
    from vescale.emulator.device_mesh import DeviceMesh
    from vescale.emulator.mesh_collectives import mesh_all_gather, mesh_all_reduce, mesh_all_to_all, mesh_broadcast, mesh_reduce_scatter, mesh_scatter
    from vescale.emulator.reduce_kernel import ReduceOp

    ed.init_process_group(world_size=4, rank=0)
    try:
        mesh = DeviceMesh("cpu", torch.arange(4).view(2, 2))
        ts = [torch.full((4,), float(r)) for r in range(4)]
        assert [t[0].item() for t in mesh_all_reduce(ts, mesh, ReduceOp.SUM, 1)] == [1.0, 1.0, 5.0, 5.0]
        assert [t[0].item() for t in mesh_all_reduce(ts, mesh, ReduceOp.SUM, 0, tree_structure=[[0, 1], [2, 3]])] == [2.0, 4.0, 2.0, 4.0]
        assert mesh_all_gather(ts, mesh, 0, 1)[2].tolist() == [2.0] * 4 + [3.0] * 4
        assert mesh_reduce_scatter(ts, mesh, ReduceOp.SUM, 0, 0)[3].tolist() == [4.0, 4.0]
        outs = [[torch.zeros(4) for _ in range(2)] for _ in range(4)]
        keep = outs[2][1]
        mesh_all_to_all(outs, [[torch.full((4,), 10.0 * r + j) for j in range(2)] for r in range(4)], mesh, 1)
        assert [[o[0].item() for o in row] for row in outs] == [[0.0, 10.0], [1.0, 11.0], [20.0, 30.0], [21.0, 31.0]] and keep[0].item() == 30.0
        got = mesh_scatter([None] * 4, [[torch.tensor([r, j]) for j in range(2)] for r in range(4)], mesh, 0)
        assert [x.tolist() for x in got] == [[0, 0], [1, 0], [0, 1], [1, 1]]
        assert [t[0].item() for t in mesh_broadcast(ts, mesh, 1)] == [0.0, 0.0, 2.0, 2.0]
        assert list(ed._world.pg_group_ranks.values())[0] == {0: 0, 1: 1, 2: 2, 3: 3}
    finally:
        ed.destroy_process_group()


def test_emulator_dtensor_list_front_end():
    """List-of-DTensors front end (legacy ``test/emulator/test_dtensor.py``): one process holds every rank's DTensor, ops go through
    the real rules rank by rank, redistributions through the emulated collectives (ring order of the Partial reduction)."""
    import torch

    from vescale_b200 import Replicate, Shard
    from vescale_b200.emulator.collectives import ring_all_reduce
    from vescale_b200.emulator.dtensor_api import EmuMesh, distribute_tensor, emu_call, redistribute_dtensor

    mesh = EmuMesh("cpu", (4,))
    torch.manual_seed(0)
    t1, t2 = torch.randn(12, 8), torch.randn(8, 12)
    expect = {(Shard(0), Replicate()): Shard(0), (Shard(1), Shard(0)): None, (Replicate(), Shard(1)): Shard(1), (Replicate(), Replicate()): Replicate(), (Shard(0), Shard(0)): Shard(0)}
    for (p1, p2), want in expect.items():
        a, b = distribute_tensor([t1] * 4, mesh, [p1]), distribute_tensor([t2] * 4, mesh, [p2])
        y = emu_call(torch.mm, a, b)
        assert (y[0].placements[0].is_partial() if want is None else y[0].placements == (want,)), (p1, p2, y[0].placements)
        y = redistribute_dtensor(y, mesh, [Replicate()])
        for d in y:
            torch.testing.assert_close(d.to_local(), t1 @ t2, rtol=1e-5, atol=1e-5)
        if want is None:  # the Partial result was reduced in NCCL's ring association order, bit for bit
            parts = [t1[:, 2 * r : 2 * r + 2] @ t2[2 * r : 2 * r + 2] for r in range(4)]
            assert torch.equal(y[0].to_local(), ring_all_reduce(parts)[0])


def test_checkpoint_planner_dedup_balance_cache_and_single_process_io(tmp_path):
    """Load-balanced dedup, plan fingerprints / LRU cache, and the save / load drivers without a process group."""
    import pytest
    import torch
    from torch.distributed.checkpoint.metadata import ChunkStorageMetadata, MetadataIndex, TensorProperties
    from torch.distributed.checkpoint.planner import SavePlan, TensorWriteData, WriteItem, WriteItemType

    from vescale_b200.checkpoint.planner import PlanLRUCache, VeScaleLoadPlanner, VeScaleSavePlanner, custom_dedup_tensors, item_bytes, plan_fingerprint
    from vescale_b200.checkpoint.state_dict_io import CheckpointException, load_state_dict, save_state_dict

    def item(name, n, off=0):
        return WriteItem(MetadataIndex(name, torch.Size([off])), WriteItemType.TENSOR,
                         tensor_data=TensorWriteData(ChunkStorageMetadata(torch.Size([off]), torch.Size([n])), TensorProperties(dtype=torch.float32), torch.Size([n + off])))

    # four ranks all offer the replicated tensors r0..r5 (sizes 6..1 MB-ish); rank 2 also owns a big private shard
    sizes = [600, 500, 400, 300, 200, 100]
    plans = [SavePlan([item(f"r{i}", n) for i, n in enumerate(sizes)] + ([item("private", 900)] if r == 2 else [])) for r in range(4)]
    out = custom_dedup_tensors(plans)
    names = [sorted(it.index.fqn for it in p.items) for p in out]
    assert sorted(n for ns in names for n in ns) == sorted([f"r{i}" for i in range(6)] + ["private"])  # every piece exactly once
    load = [sum(item_bytes(it) for it in p.items) for p in out]
    assert max(load) - min(load) <= 4 * 600 and "private" in names[2] and len(names[2]) == 1  # the loaded rank gets nothing extra
    lowest = [sum(item_bytes(it) for it in p.items) for p in [SavePlan(plans[0].items)] + [SavePlan([i for i in p.items if i.index.fqn == "private"]) for p in plans[1:]]]
    assert max(load) < max(lowest)  # better than "lowest rank writes everything"
    assert plan_fingerprint(plans[0]) == plan_fingerprint(plans[1]) != plan_fingerprint(plans[2])

    cache = PlanLRUCache(capacity=2)
    for k in "abc":
        cache.put(k, plans[0], None)
    assert cache.get("a") is None and cache.get("c") is not None and len(cache) == 2 and (cache.hits, cache.misses) == (1, 1)

    sd = {"layer": {"w": torch.arange(12.0).reshape(3, 4), "b": torch.ones(4)}, "step": 7}
    pl = VeScaleSavePlanner()
    for k in range(2):
        save_state_dict(sd, str(tmp_path / f"ck{k}"), no_dist=True, planner=pl)
    assert pl.global_plan_runs == 1 and pl.cache.hits == 1  # second save: cached plan + metadata
    dst = {"layer": {"w": torch.zeros(3, 4), "b": torch.zeros(4)}, "step": 0}
    load_state_dict(dst, str(tmp_path / "ck1"), no_dist=True)
    assert torch.equal(dst["layer"]["w"], sd["layer"]["w"]) and dst["step"] == 7
    with pytest.raises(CheckpointException, match="rank"):
        load_state_dict({"missing": torch.zeros(2)}, str(tmp_path / "ck1"), no_dist=True)
    load_state_dict({"missing": torch.zeros(2), "step": 0}, str(tmp_path / "ck1"), no_dist=True, planner=VeScaleLoadPlanner(allow_partial_load=True))


def test_checkpoint_report_service_gather_broadcast_barrier():
    """The out-of-band coordination channel of asynchronous saves: rendezvous by tag, status of a stuck rendezvous, time-outs."""
    import threading

    import pytest

    from vescale_b200.checkpoint import server_lib as sl

    W = 4
    server, addr = sl.serve(sl.ReportServicer(W))
    try:
        results = [None] * W

        def rank_main(r):
            stub = sl.get_stub(addr)
            got = sl.gather(stub, 2, r, {"rank": r, "bytes": 10 * r}, tag="plans")
            cfg = sl.broadcast(stub, 1, r, obj=("final-plan", r) if r == 1 else None, tag="final")
            sl.barrier(stub, r, tag="done")
            results[r] = (got, cfg)
            stub.close()

        ts = [threading.Thread(target=rank_main, args=(r,)) for r in range(W)]
        [t.start() for t in ts]
        [t.join(30) for t in ts]
        assert all(not t.is_alive() for t in ts)
        assert results[2][0] == [{"rank": r, "bytes": 10 * r} for r in range(W)] and all(results[r][0] is None for r in (0, 1, 3))
        assert all(results[r][1] == ("final-plan", 1) for r in range(W))
        stub = sl.get_stub(addr)
        st = sl.get_server_status(stub)
        assert st["waiting"] == {} and st["completed"] == 3
        # a rendezvous nobody else joins: visible in the status while it waits, then a time-out naming the missing ranks
        box = {}
        t = threading.Thread(target=lambda: box.setdefault("err", _raises(lambda: sl.barrier(sl.get_stub(addr), 0, tag="lonely", timeout=1.0))))
        t.start()
        import time

        time.sleep(0.3)
        st = sl.get_server_status(stub)
        assert st["waiting"]["barrier/lonely"]["missing"] == [1, 2, 3]
        t.join(10)
        assert "ranks [1, 2, 3]" in box["err"]
        with pytest.raises(RuntimeError, match="twice"):
            s2 = sl.get_stub(addr)
            threading.Thread(target=lambda: _raises(lambda: sl.barrier(s2, 0, tag="dup", timeout=1.0)), daemon=True).start()
            time.sleep(0.2)
            sl.barrier(s2, 0, tag="dup", timeout=1.0)
    finally:
        server.stop(0)


def _raises(fn):
    try:
        fn()
    except Exception as e:  # noqa: BLE001
        return str(e)
    return ""


def test_named_mem_file_server_path_api_and_pinned_pool_functions():
    """``/local_mem/<name>/...`` file API over named servers (legacy ``mem_server_lib``), pool function forms (``mem_checkpoint``)."""
    import pytest
    import torch

    import vescale_b200.checkpoint.mem_server as m
    from vescale_b200.checkpoint.pinned_pool import GLOBAL_POOL, copy_gpu_tensor_to_cpu_pinned_mem_pool, deallocate_cpu_tensor_in_pinned_mem_pool

    name = f"pytest_{os.getpid()}"
    srv = m.start_server(name, force=True)
    try:
        assert m.wait_until_fs_ready(name, 5) and open(m.get_mem_server_sock_file(name)).read() == srv.address
        p = m.get_prefix(name)
        with m.open(p + "ck/step1/a.bin", "wb") as f:
            f.write(b"abc")
        with m.open(p + "ck/step1/a.bin", "ab") as f:
            f.write(b"def")
        assert m.open(p + "ck/step1/a.bin", "rb").read() == b"abcdef" and m.listdir(p + "ck") == ["step1"] and m.listdir(p + "ck/step1") == ["a.bin"]
        m.rename(p + "ck/step1/a.bin", p + "ck/step1/b.bin")
        with pytest.raises(FileExistsError):
            with m.open(p + "ck/step1/c.bin", "wb") as f:
                f.write(b"x")
            m.rename(p + "ck/step1/c.bin", p + "ck/step1/b.bin")
        m.remove(p + "ck/step1/b.bin")
        assert not m.exists(p + "ck/step1/b.bin") and m.exists(p + "ck/step1/c.bin")
        with pytest.raises(RuntimeError, match="already running"):
            m.start_server(name)
        with pytest.raises(ValueError):
            m.exists("/tmp/not_local_mem")
    finally:
        srv.stop(0)
        os.remove(m.get_mem_server_sock_file(name))
    t = torch.arange(6.0).reshape(2, 3)
    h = copy_gpu_tensor_to_cpu_pinned_mem_pool(t)
    assert torch.equal(h, t) and h.data_ptr() != t.data_ptr()
    deallocate_cpu_tensor_in_pinned_mem_pool(h)
    assert GLOBAL_POOL is not None


def test_profiler_device_timer_global_clock_and_singleton():
    import time

    import pytest

    from vescale_b200.profiler.timer import DeviceTimer, GlobalReferenceTime, Singleton

    class Once(metaclass=Singleton):
        def __init__(self, v=0):
            self.v = v

    assert Once(1) is Once(2) and Once().v == 1
    GlobalReferenceTime.sync_events()
    t0 = time.time_ns() / 1e3
    assert GlobalReferenceTime.initialized and abs(GlobalReferenceTime.host_time(time.perf_counter()) - t0) < 5e4
    assert GlobalReferenceTime.calibrate() == 1.0  # no GPU: nothing to drift against
    t = DeviceTimer("step", is_host=True)
    for _ in range(2):
        with t:
            time.sleep(0.01)
    with pytest.raises(RuntimeError, match="not started"):
        t.stop()
    iv = t.intervals(reset=False)
    assert len(iv) == 2 and iv[0][0] < iv[1][0] and all(9e3 < d < 2e5 for _, d in iv) and abs(iv[0][0] - t0) < 1e6
    assert 18 < t.elapsed() < 400 and t.elapsed() == 0.0
    t.disable()
    t.start(); t.stop()  # noqa: E702  a disabled timer ignores both
    assert t.intervals() == [] and not t.is_enabled()


def _emulator_dtensor_lists(rank, world):
    """The reference's list-of-DTensors front end, call for call (``legacy/test/emulator/test_dtensor.py:65-110``): per-rank
    full tensors -> ``comm_api.distribute_tensor`` -> instrumented ``torch.mm`` over the lists -> ``redistribute_dtensor`` on the
    emulated collectives, against the same computation on live DTensors."""
    import vescale
    import vescale.emulator.distributed as ed
    from vescale.dtensor.placement_types import Replicate, Shard
    from vescale.emulator.comm_api import distribute_tensor, redistribute_dtensor
    from vescale.emulator.device_mesh import DeviceMesh
    from vescale.emulator.emulator_instrumentation import EmulatorInstrumentation

    ed.init_process_group(backend="nccl", world_size=world, rank=0)
    ed.set_rank(0)
    emu_mesh = DeviceMesh(device_type(), list(range(world)))
    real_mesh = vescale.dtensor.device_mesh.DeviceMesh(device_type(), list(range(world)))
    torch.manual_seed(0)
    t1, t2 = torch.randn(12, 8, device=device_type()), torch.randn(8, 12, device=device_type())
    t1_list = [t1.clone().requires_grad_() for _ in range(world)]
    t2_list = [t2.clone().requires_grad_() for _ in range(world)]
    for p1, p2 in [(Shard(0), Replicate()), (Shard(1), Shard(0)), (Replicate(), Shard(1)), (Replicate(), Replicate())]:
        d1 = distribute_tensor(t1_list, emu_mesh, [p1])
        d2 = distribute_tensor(t2_list, emu_mesh, [p2])
        assert len(d1) == world and all(isinstance(d, vescale.dtensor.dtensor.DTensor) for d in d1)
        with EmulatorInstrumentation(torch, ["mm"], [(0, 1)]):
            res = torch.mm(d1, d2)
            res = redistribute_dtensor(res, emu_mesh, [Replicate()])
        real = torch.mm(vescale.distribute_tensor(t1, real_mesh, [p1]), vescale.distribute_tensor(t2, real_mesh, [p2])).redistribute(real_mesh, [Replicate()])
        for r_emu in res:
            assert r_emu.placements == (Replicate(),)
            if p1 == Shard(1) and torch.distributed.get_backend() != "nccl":
                # the Partial result is summed in NCCL ring order by the emulator; gloo associates differently
                assert torch.allclose(real.to_local(), r_emu.to_local(), atol=1e-5)
            else:
                assert torch.equal(real.to_local(), r_emu.to_local())
    ed.destroy_process_group()


def test_emulator_list_of_dtensors_front_end():
    run_distributed(_emulator_dtensor_lists, 4)


def _fsdp2_patch(rank, world):
    """``ndtimeline.fsdp_patch``: torch's own ``fully_shard`` emits UNSHARD_AG / GRAD_RS regions once patched; unpatching restores
    the originals; numerics are untouched."""
    from torch.distributed.fsdp import fully_shard

    import vescale_b200.profiler.fsdp_patch as fp
    from vescale.ndtimeline.fsdp_patch import is_fsdp_patched, patch_fsdp
    from vescale_b200.profiler import predefined

    seen = []
    real = fp.ndtimeit

    def spy(metric, *a, **kw):
        seen.append((metric, kw.get("unit")))
        return real(metric, *a, **kw)

    fp.ndtimeit = spy
    try:
        assert not is_fsdp_patched()
        patch_fsdp()
        patch_fsdp()  # idempotent
        assert is_fsdp_patched()
        torch.manual_seed(0)
        ref = nn.Sequential(nn.Linear(8, 8), nn.Linear(8, 4)).to(device_type())
        import copy

        m = copy.deepcopy(ref)
        for layer in m:
            fully_shard(layer)
        fully_shard(m)
        x = torch.randn(4, 8, device=device_type())
        m(x).sum().backward()
        ref(x).sum().backward()
        metrics = [s[0] for s in seen]
        assert metrics.count(predefined.UNSHARD_AG) >= 2 and metrics.count(predefined.GRAD_RS) >= 2, metrics
        assert torch.allclose(m[0].weight.grad.full_tensor(), ref[0].weight.grad, atol=1e-6)
        fp.unpatch_fsdp()
        assert not is_fsdp_patched()
    finally:
        fp.ndtimeit = real
        fp.unpatch_fsdp()


def test_ndtimeline_patch_for_torch_fsdp2():
    run_distributed(_fsdp2_patch, 2)
