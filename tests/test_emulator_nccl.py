"""NCCL host-side model (tuning, chunk geometry), step-level primitives, and their agreement with the closed-form emulator."""
import pytest
import torch

from vescale_b200.emulator import EmulatorProcessGroup
from vescale_b200.emulator.algorithms import (chunk_layout, run_all_to_all, run_broadcast, run_ring_all_gather, run_ring_all_reduce, run_ring_reduce_scatter,
                                              run_tree_all_reduce)
from vescale_b200.emulator.chunk_math import calcBytePerGrain, calcBytePerStep, compute_last_chunk_size, get_info_nchannels_nthreads_proto, get_loop_info, get_pattern_info
from vescale_b200.emulator.collectives import double_tree_all_reduce, ring_all_reduce, ring_reduce_scatter, tree_all_reduce
from vescale_b200.emulator.nccl import Algo, CollInfo, Func, Pattern, Proto, algo_time, get_algo_info, init_comm, parse_nccl_debug_log
from vescale_b200.emulator.primitives import Point2PointPrimitive, RingPrimitive, Traffic, TreePrimitive
from vescale_b200.emulator.topo import DoubleTree, btree, filter_tree_structure, global_rank_to_group_rank, tree_structure_from_graph_dump


@pytest.mark.parametrize("n", [2, 3, 4, 8])
def test_step_level_ring_equals_closed_form_bitwise(n):
    torch.manual_seed(n)
    xs = [torch.randn(1000).to(torch.bfloat16) for _ in range(n)]
    for nch, ce in ((1, None), (2, 32), (3, 16)):
        tr = Traffic()
        a, b = ring_all_reduce(xs, nchannels=nch, chunk_elems=ce), run_ring_all_reduce(xs, nchannels=nch, chunk_elems=ce, traffic=tr)
        assert all(torch.equal(x, y) for x, y in zip(a, b))
        # every element crosses n-1 links in the reduce-scatter phase and n-1 in the all-gather phase, whatever the chunking
        assert tr.total_bytes == 2 * (n - 1) * 1000 * 2 and abs(tr.sent_by(0) - tr.total_bytes // n) <= 64 * nch * 2 * 2 * (n - 1)
    ring = list(range(n))[::-1]
    assert all(torch.equal(x, y) for x, y in zip(ring_all_reduce(xs, ring=ring), run_ring_all_reduce(xs, ring=ring)))
    ys = [torch.randn(n * 64) for _ in range(n)]
    tr = Traffic()
    assert all(torch.equal(x, y) for x, y in zip(ring_reduce_scatter(ys, ring=ring), run_ring_reduce_scatter(ys, ring=ring, traffic=tr)))
    assert tr.steps == n - 1 and tr.total_bytes == n * (n - 1) * 64 * 4
    g = run_ring_all_gather([y[:64] for y in ys], ring=ring)
    assert all(torch.equal(gi, torch.cat([y[:64] for y in ys])) for gi in g)
    b = run_broadcast(ys, src=n - 1)
    assert all(torch.equal(bi, ys[n - 1]) for bi in b)


def test_tree_operand_order_is_local_first():
    """A node with two children accumulates (local + child0) + child1 — NCCL's reduceCopy source order — not (child0 + child1) + local.
    Values chosen so the two orders differ in bf16."""
    vals = [0.0, 1.0, 256.0, 1.0]  # 4 ranks: btree(4) = 0 -> 2 -> {1, 3}; bf16 ulp at 256 is 2: (256 + 1) + 1 = 256, 256 + (1 + 1) = 258
    xs = [torch.tensor([v], dtype=torch.bfloat16) for v in vals]
    t = btree(4)
    assert t.children[2] == [1, 3] and t.children[0] == [2]
    prim = TreePrimitive(xs, t)
    total = prim.reduce_up(0, 1)
    bf = lambda v: torch.tensor(v, dtype=torch.bfloat16)  # noqa: E731
    local_first = bf(vals[0]) + ((bf(vals[2]) + bf(vals[1])) + bf(vals[3]))
    children_first = bf(vals[0]) + ((bf(vals[1]) + bf(vals[3])) + bf(vals[2]))
    assert float(total) == float(local_first) and float(local_first) != float(children_first)
    # closed forms follow the same rule
    assert float(double_tree_all_reduce(xs, chunk_elems=1)[0]) == float(local_first)
    xs3 = [torch.tensor([v], dtype=torch.bfloat16) for v in (256.0, 1.0, 1.0)]  # heap tree: 0 -> {1, 2}
    assert float(tree_all_reduce(xs3)[0]) == float((bf(256.0) + bf(1.0)) + bf(1.0)) == 256.0


@pytest.mark.parametrize("n", [2, 4, 7, 8])
def test_tree_runner_equals_closed_form_and_hierarchical(n):
    torch.manual_seed(0)
    xs = [torch.randn(515).to(torch.bfloat16) for _ in range(n)]
    tr = Traffic()
    a, b = double_tree_all_reduce(xs, chunk_elems=64), run_tree_all_reduce(xs, chunk_elems=64, traffic=tr)
    assert all(torch.equal(x, y) for x, y in zip(a, b)) and all(torch.equal(b[0], bi) for bi in b)
    assert tr.total_bytes == 2 * (n - 1) * 515 * 2  # every element climbs n-1 edges and descends n-1 edges
    if n == 8:  # 2 nodes x 4 GPUs: chains inside a node, binary tree across
        structure = [[0, 1, 2, 3], [4, 5, 6, 7]]
        h = DoubleTree(structure, list(range(8)), {r: r for r in range(8)})
        out = run_tree_all_reduce(xs, trees=h, chunk_elems=128)
        ref = sum(x.float() for x in xs)
        assert all(torch.equal(out[0], o) for o in out) and torch.allclose(out[0].float(), ref, atol=0.25, rtol=0.05)
        assert filter_tree_structure(structure, [1, 2, 5]) == [[1, 2], [5]] and global_rank_to_group_rank([5, 1], {1: 0, 2: 1, 5: 2}) == [2, 0]


def test_protocol_chunk_layouts_partition_the_buffer_and_change_the_chains():
    torch.manual_seed(1)
    n, count = 4, 300001
    ys = [torch.randn(count).to(torch.bfloat16) for _ in range(n)]
    ref = sum(y.float() for y in ys)
    outs = {}
    for proto, nt, ce in ((int(Proto.LL), 512, 4096), (int(Proto.LL128), 640, 19200), (int(Proto.SIMPLE), 544, 4096)):
        layout = chunk_layout(count, n, 4, ce, proto, nt, 2)
        cover = torch.zeros(count, dtype=torch.int32)
        for lp in layout:
            for c in range(n):
                o, ne = lp.span(c)
                cover[o:o + ne] += 1
        assert bool((cover == 1).all())
        r = run_ring_all_reduce(ys, layout=layout)
        assert all(torch.equal(r[0], ri) for ri in r) and torch.allclose(r[0].float(), ref, atol=0.25, rtol=0.05)
        outs[proto] = r[0]
    # same data, same ring, different protocol => different element -> chunk assignment => (some) different bf16 roundings
    assert not torch.equal(outs[int(Proto.LL)], outs[int(Proto.SIMPLE)])
    assert calcBytePerStep(int(Proto.SIMPLE), 1 << 22) == 1 << 19 and calcBytePerStep(int(Proto.LL), 1 << 19) == 1 << 15 and calcBytePerGrain(int(Proto.LL128)) == 60


def test_tuning_model_shape_and_enqueue_decisions():
    comm = init_comm(8, nvls=False)
    # small messages take the low-latency protocol, large ones Simple; time grows with size; LL never wins at 256 MB
    small = get_algo_info(comm, CollInfo(int(Func.ALL_REDUCE), 256, 4))
    big = get_algo_info(comm, CollInfo(int(Func.ALL_REDUCE), 1 << 26, 4))
    assert small.proto == int(Proto.LL) and big.proto == int(Proto.SIMPLE) and big.algo == int(Algo.RING) and small.time_us < big.time_us
    assert small.n_channels == 1 and small.n_threads < 512 and big.n_channels == comm.n_channels and big.n_threads == 512 + 32
    assert big.chunk_steps == 4 and big.chunk_size == (1 << 22) // 8 * 4 and big.nsteps_per_loop == 14 and big.nchunks_per_loop == 8
    assert big.n_loops == -(-(1 << 28) // (big.n_channels * 8 * big.chunk_size))
    times = [get_algo_info(comm, CollInfo(int(Func.ALL_REDUCE), 1 << k, 1)).time_us for k in range(8, 30, 3)]
    assert times == sorted(times)
    # all-gather / reduce-scatter never use the tree; forcing (algo, proto) is honoured
    for f in (Func.ALL_GATHER, Func.REDUCE_SCATTER, Func.BROADCAST):
        assert get_algo_info(comm, CollInfo(int(f), 1 << 20, 2)).algo == int(Algo.RING)
    forced = get_algo_info(comm, CollInfo(int(Func.ALL_REDUCE), 1 << 20, 2), force=(int(Algo.TREE), int(Proto.SIMPLE)))
    assert forced.algo == int(Algo.TREE) and forced.pattern == int(Pattern.TREE_UP_DOWN) and forced.last_chunk_size > 0 and forced.n_threads == 512 + 4 * 32
    assert algo_time(comm, Func.ALL_REDUCE, Algo.NVLS, Proto.SIMPLE, 1 << 20) == -1.0  # disabled
    # with NVLS available, the switch reduction wins the mid-size range on Hopper+
    nv = init_comm(8, compcap=100)
    assert nv.nvls and get_algo_info(nv, CollInfo(int(Func.ALL_REDUCE), 1 << 22, 2)).algo == int(Algo.NVLS)
    # two nodes: tree beats ring for mid-size all-reduce (log-depth latency)
    two = init_comm(16, n_nodes=2, compcap=90, nvls=False)
    assert get_algo_info(two, CollInfo(int(Func.ALL_REDUCE), 1 << 18, 2)).algo == int(Algo.TREE)
    assert get_pattern_info(int(Func.ALL_REDUCE), int(Algo.RING)) == int(Pattern.RING_TWICE) and get_loop_info(int(Pattern.RING), 8) == (7, 8)
    a, p, nc, nt = get_info_nchannels_nthreads_proto(int(Func.ALL_REDUCE), 1 << 20, 2, 8)
    assert (a, p) == (int(Algo.RING), int(Proto.LL)) and compute_last_chunk_size(comm, int(Func.ALL_REDUCE), 1 << 20, 2, a, p, nc, nt) > 0


def test_model_driven_process_group_and_graph_dump():
    pg = EmulatorProcessGroup(4, algo="nccl")
    torch.manual_seed(3)
    for cnt in (100, 70000):
        xs = [torch.randn(cnt).to(torch.bfloat16) for _ in range(4)]
        out = pg.all_reduce(xs)
        assert all(torch.equal(out[0], o) for o in out) and torch.allclose(out[0].float(), sum(x.float() for x in xs), atol=0.2, rtol=0.05)
        assert pg.last_traffic.total_bytes == 2 * 3 * cnt * 2 and pg.last_traffic.estimate_us() > 0
    pg.force_algo_proto = (int(Algo.TREE), int(Proto.SIMPLE))
    out = pg.all_reduce(xs)
    assert pg.last_info.algo == int(Algo.TREE) and all(torch.equal(out[0], o) for o in out)
    xml = """<graphs version="1"><graph id="0" pattern="4" crossnic="0" nchannels="2" speedintra="40" speedinter="40" latencyinter="0" typeintra="NVL" typeinter="PIX" samechannels="1">
      <channel><gpu dev="0"/><gpu dev="2"/><gpu dev="3"/><gpu dev="1"/></channel><channel><gpu dev="0"/><gpu dev="1"/><gpu dev="3"/><gpu dev="2"/></channel></graph>
      <graph id="1" pattern="1" nchannels="2" speedintra="40" speedinter="40" typeintra="NVL" typeinter="PIX" samechannels="1"><channel><gpu dev="0"/><gpu dev="2"/><gpu dev="3"/><gpu dev="1"/></channel></graph></graphs>"""
    comm = init_comm(4, xml=xml)
    assert comm.graphs[int(Algo.RING)].channels == [[0, 2, 3, 1], [0, 1, 3, 2]] and comm.n_channels == 2 and comm.graphs[int(Algo.RING)].bw_intra == 40.0
    assert comm.bandwidths[int(Func.ALL_REDUCE)][int(Algo.RING)][int(Proto.SIMPLE)] == pytest.approx(2 * 40.0 * 4 / 6)
    assert tree_structure_from_graph_dump(xml, n_nodes=2) == [[0, 2, 3, 1], [4, 6, 7, 5]]


def test_debug_log_parser_and_p2p():
    log = """host:1:2 [0] NCCL INFO Channel 00/02 :    0   1   2   3
host:1:2 [0] NCCL INFO Channel 01/02 :    0   3   2   1
host:1:2 [0] NCCL INFO Trees [0] 1/-1/-1->0->-1 [1] 3/-1/-1->0->-1
host:1:2 [1] NCCL INFO Trees [0] 2/-1/-1->1->0 [1] -1/-1/-1->1->2
host:1:2 [0] NCCL INFO 2 coll channels, 0 collnet channels, 0 nvls channels, 2 p2p channels, 2 p2p channels per peer
host:1:2 [0] NCCL INFO AllReduce: opCount 1a sendbuff 0x7 recvbuff 0x7 count 1048576 datatype 9 op 0 root 0 comm 0x5 [nranks=4] stream 0x1
host:1:2 [0] NCCL INFO 2097152 Bytes -> Algo 1 proto 2 time 41.5
"""
    res = parse_nccl_debug_log(log)
    assert res.ring_orders() == [[0, 1, 2, 3], [0, 3, 2, 1]] and res.n_channels == 2
    assert res.trees[0][0] == ([1, -1, -1], -1) and res.trees[1][1] == ([-1, -1, -1], 2)
    c = res.collectives[0]
    assert (c.func, c.op_count, c.count, c.nranks, c.algo, c.proto, c.time_us) == ("AllReduce", 0x1A, 1048576, 4, 1, 2, 41.5) and res.choice_for(2097152) == (1, 2)
    assert parse_nccl_debug_log(log, rank=1).rings == {} and 0 in parse_nccl_debug_log(log, rank=1).trees
    ins = [[torch.full((2,), 10.0 * s + d) for d in range(3)] for s in range(3)]
    tr = Traffic()
    out = run_all_to_all(ins, traffic=tr)
    assert all(float(out[d][s][0]) == 10.0 * s + d for s in range(3) for d in range(3)) and tr.steps == 2 and tr.total_bytes == 6 * 8
    p2p = Point2PointPrimitive(2)
    p2p.send(0, 1, torch.ones(1))
    with pytest.raises(AssertionError):
        p2p.assert_drained()
    with pytest.raises(AssertionError):
        RingPrimitive([torch.zeros(4), torch.zeros(4)]).recv(0, 0, 4)  # nothing was sent


def test_emulator_per_dim_redistribute_and_nccl_named_helpers():
    """Redistribution as the real planner does it (one collective per differing mesh dim), transition objects, NCCL-named helpers."""
    import torch

    from vescale_b200.emulator import comm_api as ca
    from vescale_b200.emulator.algorithms import contract_tensor_list, expand_tensor_list
    from vescale_b200.emulator.comm_primitive import BaseRedistributeFunc
    from vescale_b200.emulator.nccl import comm, constants as c, profiler_result as pr, tuning
    from vescale_b200.mesh import DeviceMesh
    from vescale_b200.placement import Partial, Replicate, Shard

    mesh = DeviceMesh("cpu", torch.arange(4).reshape(2, 2), _rank=0)
    full = torch.arange(48.0).reshape(8, 6)
    cases = [([Shard(0), Shard(1)], [Replicate(), Shard(0)], ["all_gather", "all_to_all"]), ([Partial(), Shard(1)], [Shard(0), Shard(1)], ["reduce_scatter"]),
             ([Partial(), Partial()], [Replicate(), Shard(0)], ["all_reduce", "reduce_scatter"]), ([Shard(0), Shard(0)], [Replicate(), Replicate()], None)]
    for src, dst, want in cases:
        loc = ca.distribute_tensor(full, mesh, src)
        tr = []
        out = ca.DTensorRedistribute.apply(loc, full.shape, mesh, src, dst, trace=tr)
        assert torch.equal(ca.full_tensor(out, full.shape, mesh, dst), ca.full_tensor(loc, full.shape, mesh, src))
        if want is not None:
            assert [t[2] for t in tr] == want
        else:
            assert tr == []  # two mesh dims on one tensor dim: no single-step path, full-tensor route
    f = BaseRedistributeFunc.of(Partial("sum"), Shard(1))
    assert (f.name, f.collective, f.bound["shard_dim"]) == ("P2S", "reduce_scatter", 1) and "P2S" in repr(f)
    e = expand_tensor_list([torch.full((2,), float(i)) for i in range(3)])
    assert e[1].tolist() == [0, 0, 1, 1, 0, 0] and contract_tensor_list(e)[2].tolist() == [2, 2]

    assert (c.div_up(7, 2), c.round_up(7, 4), c.align_up(5000, 4096), c.log2i(9), c.NcclFunc.ALL_REDUCE) == (4, 8, 8192, 3, c.Func.ALL_REDUCE)
    cm = comm.init(8)
    assert comm.compute_buff_sizes(cm, {"NCCL_BUFFSIZE": "8388608"})[int(c.Proto.SIMPLE)] == 8388608
    assert comm.compute_buff_sizes(cm, {})[int(c.Proto.SIMPLE)] == c.DEFAULT_BUFFSIZE[c.Proto.SIMPLE]
    assert (tuning.get_nthreads("t", 100, 64, 512, 256), tuning.get_nthreads("t", -2, 64, 512, 256), tuning.get_nthreads("t", 32, 64, 512, 256)) == (512, 256, 64)
    assert tuning.nccl_topo_get_algo_time is tuning.algo_time and tuning.DIVUP(9, 4) == 3 and tuning.get_cpu_info()[0] in (1, 2, 3)
    topo = ('<system version="1"><cpu numaid="0" arch="x86_64" vendor="AuthenticAMD"><pci><gpu dev="0" sm="100" rank="0"><nvlink target="x" count="18" tclass="0x068000"/></gpu>'
            '<gpu dev="1" sm="90" rank="1"/></pci><nic><net name="mlx5_0" speed="400000" gdr="1"/></nic></cpu></system>')
    t = pr.parse_nccl_topo(topo)
    assert t["cpu_arch_amd"] and t["nvswitch"] and pr.get_default_min_max_compcap(t) == (90, 100) and pr.get_default_min_max_compcap() == (100, 100)
    g = pr.parse_graph_xml('<graphs version="1"><graph id="0" pattern="4" nchannels="2" speedintra="40" speedinter="40" typeintra="NVL" typeinter="PIX">'
                           '<channel><gpu dev="0"/><gpu dev="1"/></channel><channel><gpu dev="1"/><gpu dev="0"/></channel></graph></graphs>')
    assert g[0].channels == [[0, 1], [1, 0]] and g[0].bw_intra == 40.0
    assert comm.nccl_info_set_derived(comm.CollInfo(func=int(c.Func.ALL_GATHER), count=1024), 8).total_bytes == 8 * 4096
