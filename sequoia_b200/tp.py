"""Target tensor parallelism across the GPUs of one box (the reference has no multi-GPU code).

The target's layers are sharded Megatron-style inside LlamaRunner (column-parallel qkv / gate_up, row-parallel
o_proj / down_proj + NCCL sum-allreduce over NVLink, KV cache sharded by kv head).  The draft model, the tree state,
sampling and the accept walk live on rank 0 only ("driver").  Every other rank ("follower") owns a target shard and
mirrors the driver's target-side work:

    driver                                   follower
    ctrl  = [op, a, b, mode] - eager bcast -> blocks on it (the only host sync besides the driver's own per verify)
    tokens / position_ids / state --bcast->  (inside the captured graphs on both sides)
    target forward (allreduce x 2L)  <---->  target forward shard, lm_head skipped
    accept walk
    accept_idx / state   ------ bcast --->   KV compaction of the shard

The collectives are issued in the same order on every rank; in the steady state they are all inside one CUDA graph
per rank.  `mode` tells the follower HOW MANY times and how the driver executes the steady sequence for this control
word, so both sides always issue the same number of collectives: MODE_REPLAY = one graph replay, MODE_EAGER = one eager
execution (benchmark=True / use_graphs=False on the driver), MODE_CAPTURE = the driver is about to capture its graph
(one eager warm-up execution + one replay; the capture pass itself executes nothing).  A follower serves ONE growmap
(the one it was constructed with); driving a second growmap through the same followers is not supported."""
from __future__ import annotations

import torch
import torch.distributed as dist

from .tree import _Static

OP_STOP, OP_STEADY, OP_FIRST, OP_CLEAR, OP_BARRIER, OP_MICRO = 0, 1, 2, 3, 4, 5
MODE_REPLAY, MODE_EAGER, MODE_CAPTURE = 0, 1, 2


class TPDriver:
    """Rank-0 side: control messages + the two in-graph broadcast points."""

    def __init__(self, group, device, peer=None):
        self.group, self.device = group, torch.device(device)
        self.ctrl = torch.zeros(4, dtype=torch.int64, device=self.device)
        self.src = dist.get_global_rank(group, 0) if group is not dist.group.WORLD else 0
        # peer-memory mailboxes (peer.PeerBuffers.publish / consume) instead of NCCL broadcasts for the in-graph messages
        self.peer = peer if (peer is not None and peer.msg_on) else None

    def send_ctrl(self, op: int, a: int = 0, b: int = 0, mode: int = 0):
        self.ctrl.copy_(torch.tensor([op, a, b, mode], dtype=torch.int64), non_blocking=False)
        dist.broadcast(self.ctrl, self.src, group=self.group)

    def barrier(self):
        """Device-synchronise every rank and meet in a collective barrier (bench.py brackets its timed region with it)."""
        self.send_ctrl(OP_BARRIER)
        if self.device.type == "cuda":
            torch.cuda.synchronize()
        dist.barrier(group=self.group)

    def bcast_inputs(self, rt):
        if self.peer is not None and self.peer.fits([rt.tokens, rt.position_ids, rt.state]):
            self.peer.publish(0, [rt.tokens, rt.position_ids, rt.state])
            return
        dist.broadcast(rt.tokens, self.src, group=self.group)
        dist.broadcast(rt.position_ids, self.src, group=self.group)
        dist.broadcast(rt.state, self.src, group=self.group)

    def bcast_accept(self, rt):
        if self.peer is not None and self.peer.fits([rt.accept_idx, rt.state]):
            self.peer.publish(1, [rt.accept_idx, rt.state])
            return
        dist.broadcast(rt.accept_idx, self.src, group=self.group)
        dist.broadcast(rt.state, self.src, group=self.group)


def attach_tp(draft_engine, target_engine, group):
    """Mark the (rank-0) target engine as tensor-parallel so Tree runtimes broadcast to the follower ranks."""
    target_engine._tp_driver = TPDriver(group, target_engine.device, getattr(target_engine.engine.runner, "peer", None))
    return target_engine._tp_driver


def stop_followers(group, device):
    d = TPDriver(group, device)
    d.send_ctrl(OP_STOP)
    if torch.device(device).type == "cuda":
        torch.cuda.synchronize()


def micro_allreduce(runner, k: int, n: int):
    """k back-to-back fused all-reduce + residual + RMSNorm launches of n rows on this rank (every rank must issue the
    same k: each launch handshakes with its peers).  Used by bench.py to time the kernel on the device."""
    for i in range(k):
        runner.peer.allreduce_add_rmsnorm(i & 1, runner.hidden, runner.norm, runner.normed, n, runner.eps)


class TPFollower:
    """Non-zero ranks: serve target-shard forwards / KV compactions until OP_STOP."""

    def __init__(self, target_engine, grow_map: dict, greedy: bool, M: int, device, group):
        self.target, self.group, self.device = target_engine, group, torch.device(device)
        self.st = _Static(grow_map, self.device)
        S = self.st.S
        dev = self.device
        self.M = M
        self.tokens = torch.zeros(M, dtype=torch.int64, device=dev)
        self.position_ids = torch.zeros(M, dtype=torch.int64, device=dev)
        self.storage_ids = torch.arange(M, dtype=torch.int64, device=dev)
        self.accept_idx = torch.zeros(max(S, 8), dtype=torch.int32, device=dev)
        self.state = torch.zeros(16, dtype=torch.int32, device=dev)
        self.ctrl = torch.zeros(4, dtype=torch.int64, device=dev)
        self.src = 0
        self.graph = None
        self.use_graphs = True
        peer = getattr(target_engine.engine.runner, "peer", None)
        self.peer = peer if (peer is not None and peer.msg_on) else None

    def _mask_kw(self):
        return dict(tree_bits=self.st.tree_bits, tree_words=self.st.tree_words, tree_size=self.st.S)

    def _recv_inputs(self):
        if self.peer is not None and self.peer.fits([self.tokens, self.position_ids, self.state]):
            self.peer.consume(0, [self.tokens, self.position_ids, self.state])
            return
        dist.broadcast(self.tokens, self.src, group=self.group)
        dist.broadcast(self.position_ids, self.src, group=self.group)
        dist.broadcast(self.state, self.src, group=self.group)

    def _recv_accept_and_gather(self):
        if self.peer is not None and self.peer.fits([self.accept_idx, self.state]):
            self.peer.consume(1, [self.accept_idx, self.state])
        else:
            dist.broadcast(self.accept_idx, self.src, group=self.group)
            dist.broadcast(self.state, self.src, group=self.group)
        self.target.engine.kv_cache.gather_from_state(self.accept_idx, self.state, max(self.st.max_depth, 1))

    def _steady(self):
        S = self.st.S
        self._recv_inputs()
        self.target.engine.runner.forward(S, self.tokens, self.position_ids, self.storage_ids, state=self.state, n0=0,
                                          kv_end=S, skip_lm_head=True, **self._mask_kw())
        self._recv_accept_and_gather()

    def _first(self, start: int, P: int):
        S = self.st.S
        end = P + S - 1
        self._recv_inputs()
        self.target.engine.runner.forward(end - start, self.tokens, self.position_ids, self.storage_ids, state=None,
                                          n0=start, kv_end=end, prefix_len=P, skip_lm_head=True, **self._mask_kw())
        self._recv_accept_and_gather()

    @torch.inference_mode()
    def serve(self):
        while True:
            dist.broadcast(self.ctrl, self.src, group=self.group)
            op, a, b, mode = self.ctrl.tolist()
            if op == OP_STOP:
                if self.device.type == "cuda":
                    torch.cuda.synchronize()
                return
            if op == OP_BARRIER:
                if self.device.type == "cuda":
                    torch.cuda.synchronize()
                dist.barrier(group=self.group)
            elif op == OP_CLEAR:
                self.target.clear_kv()
            elif op == OP_MICRO:                                # bench.py: `a` fused all-reduce launches of `b` rows
                micro_allreduce(self.target.engine.runner, int(a), int(b))
            elif op == OP_FIRST:
                self._first(a, b)
            elif op == OP_STEADY:
                if mode == MODE_EAGER or not self.use_graphs:
                    self._steady()                              # one execution on the driver, one here
                    if mode == MODE_CAPTURE:
                        self._steady()                          # (graphs disabled here: mirror the replay eagerly)
                    continue
                if mode == MODE_CAPTURE:
                    # the driver runs a warm-up execution, a capture pass (executes nothing) and the first replay
                    self._warm_steady()
                    if self.graph is None:
                        self._capture_steady()
                elif self.graph is None:                        # driver replays a graph this rank never captured
                    self._steady()
                    continue
                self.graph.replay()

    def _warm_steady(self):
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self._steady()                      # pairs with the driver's warm-up execution
            s.synchronize()
        torch.cuda.current_stream().wait_stream(s)

    def _capture_steady(self):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._steady()
        self.graph = g
