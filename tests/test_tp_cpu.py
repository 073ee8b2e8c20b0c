"""N > 1 host logic on CPU (gloo, world_size 2): Megatron sharding arithmetic of the target weights and the
driver / follower control protocol of sequoia_b200.tp (the CUDA kernels themselves need a GPU and are stubbed)."""
import os
import socket
import sys
import types

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _init(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)


def _shard_worker(rank, world, port, q):
    import cases
    from sequoia_b200.model import _DictSource, config_from, load_sharded_weights
    _init(rank, world, port)
    cfg_o, w = cases.model_weights("target_gqa")                # H=4, Hkv=2 -> 2 ranks: 2 q heads + 1 kv head each
    cfg = config_from(cfg_o)
    src = _DictSource({k: v.float() for k, v in w.items()}, "cpu")
    src.get = lambda name, shape, _s=src: _s.sd[name]            # keep fp32 on CPU for an exact comparison
    full = load_sharded_weights(cfg, src, 0, 1)["layers"][0]
    mine = load_sharded_weights(cfg, src, rank, world)["layers"][0]
    g = torch.Generator().manual_seed(3)
    x = torch.randn(5, cfg.hidden_size, generator=g)
    D, H, Hkv, I = cfg.head_dim, cfg.num_attention_heads, cfg.num_key_value_heads, cfg.intermediate_size
    # column-parallel qkv: my rows are the matching slices of the full projection
    qkv_full = x @ full["wqkv"].t()
    qkv_mine = x @ mine["wqkv"].t()
    h2, k2 = H // world, Hkv // world
    exp = torch.cat([qkv_full[:, rank * h2 * D:(rank + 1) * h2 * D],
                     qkv_full[:, H * D + rank * k2 * D: H * D + (rank + 1) * k2 * D],
                     qkv_full[:, (H + Hkv) * D + rank * k2 * D:(H + Hkv) * D + (rank + 1) * k2 * D]], dim=1)
    ok = torch.allclose(qkv_mine, exp, atol=1e-5)
    # row-parallel o_proj / down_proj: partial products summed by the allreduce equal the full product
    a_full = torch.randn(5, H * D, generator=g)
    part = a_full[:, rank * h2 * D:(rank + 1) * h2 * D] @ mine["wo"].t()
    dist.all_reduce(part)
    ok &= torch.allclose(part, a_full @ full["wo"].t(), atol=1e-4)
    gu_full = x @ full["wgu"].t()
    act_full = torch.nn.functional.silu(gu_full[:, :I]) * gu_full[:, I:]
    gu = x @ mine["wgu"].t()
    Ir = I // world
    act = torch.nn.functional.silu(gu[:, :Ir]) * gu[:, Ir:]
    ok &= torch.allclose(act, act_full[:, rank * Ir:(rank + 1) * Ir], atol=1e-5)
    down = act @ mine["wd"].t()
    dist.all_reduce(down)
    ok &= torch.allclose(down, act_full @ full["wd"].t(), atol=1e-4)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_megatron_sharding_math_gloo():
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(60) for p in procs]
    assert res == [(0, True), (1, True)]


class _StubEngine:
    """Records what a follower rank would run on its target shard."""

    def __init__(self, log):
        self.log = log
        outer = self

        class _Runner:
            def forward(self, n, tokens, position_ids, storage_ids, **kw):
                outer.log.append(("forward", n, int(tokens[:4].sum()), kw.get("n0"), kw.get("kv_end"),
                                  kw.get("prefix_len", 0), kw.get("skip_lm_head")))

        class _KV:
            def gather_from_state(self, idx, state, max_n, zero_tail=False):
                outer.log.append(("gather", int(state[3]), int(state[4]), idx[:2].tolist()))

        self.engine = types.SimpleNamespace(runner=_Runner(), kv_cache=_KV())

    def clear_kv(self):
        self.log.append(("clear",))


def _proto_worker(rank, world, port, q):
    import cases
    from sequoia_b200 import tp
    _init(rank, world, port)
    gm = cases.load_growmap("L40_growmaps/8x8-tree.pt")
    S, M = gm["size"], 256
    if rank == 0:
        drv = tp.TPDriver(dist.group.WORLD, "cpu")
        rt = types.SimpleNamespace(tokens=torch.arange(M), position_ids=torch.arange(M), state=torch.zeros(16, dtype=torch.int32),
                                   accept_idx=torch.zeros(S, dtype=torch.int32))
        drv.send_ctrl(tp.OP_CLEAR)
        rt.state[0] = 100
        drv.send_ctrl(tp.OP_FIRST, 0, 100)
        drv.bcast_inputs(rt)
        rt.state[3], rt.state[4] = 2, 100
        rt.accept_idx[:2] = torch.tensor([101, 109], dtype=torch.int32)
        drv.bcast_accept(rt)
        for step in range(2):
            rt.tokens += 1
            drv.send_ctrl(tp.OP_STEADY)
            drv.bcast_inputs(rt)
            rt.state[3], rt.state[4] = step, 103 + step
            drv.bcast_accept(rt)
        drv.send_ctrl(tp.OP_STOP)
        q.put((0, "done"))
    else:
        log = []
        f = tp.TPFollower(_StubEngine(log), gm, False, M, "cpu", dist.group.WORLD)
        f.use_graphs = False
        f.serve()
        q.put((1, log))
    dist.destroy_process_group()


def test_tp_driver_follower_protocol_gloo():
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_proto_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = dict(q.get(timeout=120) for _ in range(2))
    [p.join(60) for p in procs]
    log = res[1]
    S = 65
    assert log[0] == ("clear",)
    assert log[1] == ("forward", 100 + S - 1, 0 + 1 + 2 + 3, 0, 100 + S - 1, 100, True)      # OP_FIRST: rows [0, P+S-1)
    assert log[2] == ("gather", 2, 100, [101, 109])
    assert log[3] == ("forward", S, 1 + 2 + 3 + 4, 0, S, 0, True) and log[4] == ("gather", 0, 103, [101, 109])
    assert log[5] == ("forward", S, 2 + 3 + 4 + 5, 0, S, 0, True) and log[6] == ("gather", 1, 104, [101, 109])
    assert len(log) == 7
