#!/usr/bin/env python
"""Benchmark of the Sequoia hot path on B200 (BASELINE.json metric: decoded tokens/s + mean accepted length/step).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config c2]

One "step" = one construct_grow_map() + verify() iteration of tests/testbed.py's simulation_fast loop (:80-87).
Workload (config c2, BASELINE.json configs[1]): random-init Llama-68m draft -> random-init Llama-2-7B target, growmap
A100_growmaps/68m_7b/growmaps/A100-CNN-68m-7b-stochastic.pt (128-node tree), T=0.6, P=1.0, M=384, synthetic prompts
torch.randint(3, 32000, (128,)) (seed 17), decode until 256 tokens.  With N > 1 ranks the target is tensor-sharded
over the N GPUs (NCCL allreduce, 2 per layer) while the draft stays on rank 0: total work is fixed => "strong".

Prints ONE JSON line on rank 0 (see README / DESIGN.md for every field).  `value` is timed with CUDA events around
the decode loops with everything already resident in HBM; `e2e` goes through the public API from pinned HOST buffers
(prompt H2D, Tree construction incl. its CPU-drawn random numbers, prefill, decode, D2H of the result).
`--impl reference` times the reference's own algorithm (the torch-CPU oracle port, oracle/) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

CONFIGS = {
    # name: (draft, target, growmap, greedy, T, top_p, M, prefix, max_len)
    "c1": ("llama-68m", "llama-160m", "L40_growmaps/2-chain.pt", True, 0.6, 1.0, 256, 128, 256),
    "c2": ("llama-68m", "llama-2-7b", "A100_growmaps/68m_7b/growmaps/A100-CNN-68m-7b-stochastic.pt", False, 0.6, 1.0, 384,
           128, 256),
    # c2 with the growmap tree_search.py derives from this GPU's measured draft/verify times (B200_growmaps/)
    "c2b": ("llama-68m", "llama-2-7b", "B200_growmaps/68m_7b-demo_acceptance.pt", False, 0.6, 1.0, 384, 128, 256),
    "c3": ("llama-68m", "llama-2-13b", "L40_growmaps/8x8-tree.pt", False, 0.6, 1.0, 384, 128, 256),
    "c4": ("llama-2-7b", "llama-2-70b", "L40_growmaps/L40-CNN-7b-70b-stochastic.pt", False, 0.6, 1.0, 1024, 128, 256),
}
METRIC = "decoded tokens/sec (mean accepted len/step in config.accepted_tokens_per_step), Sequoia tree speculative decoding, 68m->7B Llama (config c2 unless --config says otherwise)"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------------
def synthetic_prompts(n, length, seed=17):
    from data_converter import synthetic_prompts as sp
    return sp(n, length, 32000, seed)


def run_b200(args):
    import torch.distributed as dist
    from sequoia_b200 import _lib
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    tp_group = None
    if world > 1:
        # NCCL prints its version banner on stdout when the first communicator is created; the contract is ONE JSON line,
        # so create the communicator with fd 1 pointed at /dev/null
        sys.stdout.flush()
        saved_fd, null_fd = os.dup(1), os.open(os.devnull, os.O_WRONLY)
        os.dup2(null_fd, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device(dev))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            os.dup2(saved_fd, 1)
            os.close(null_fd)
            os.close(saved_fd)
        tp_group = dist.group.WORLD
    dname, tname, gm_path, greedy, T, top_p, M, prefix, max_len = CONFIGS[args.config]
    grow_map = torch.load(os.path.join(ROOT, gm_path))
    S = grow_map["size"]
    torch.manual_seed(17)
    from sequoia_b200.engine import GraphInferenceEngine, GraphInferenceEngineTG
    from sequoia_b200.tp import TPFollower, attach_tp
    from sequoia_b200.tree import GreedyTree, SpecTree
    target = GraphInferenceEngineTG(M, f"random-init:{tname}:2", device=dev, tp_group=tp_group)
    n_prompts_max = 4096
    prompts = synthetic_prompts(64, prefix)

    def barrier():
        if world > 1:
            target._tp_driver.barrier()      # follower ranks are slaved to rank 0: sync + barrier through the control op
        torch.cuda.synchronize()

    def finish():
        sys.stdout.flush()
        sys.stderr.flush()
        if world > 1:
            os._exit(0)                      # NCCL communicators captured in CUDA graphs: skip the slow teardown

    if rank != 0:
        # follower ranks: target shard only, driven by rank 0's broadcasts
        TPFollower(target, grow_map, greedy, M, dev, tp_group).serve()
        finish()
        return
    draft = GraphInferenceEngine(M, f"random-init:{dname}:1", device=dev)
    if world > 1:
        attach_tp(draft, target, tp_group)
    buf = dict(attn_mask=torch.full((M, M), torch.finfo(torch.float16).min, dtype=torch.float16, device=dev),
               sequence=torch.arange(M, device=dev).unsqueeze(-1), new_tokens_buffer=torch.zeros(M, device=dev).long(),
               parents_buffer=torch.zeros(M, device=dev).long(), position_ids=torch.zeros(M, device=dev).long())
    cls = GreedyTree if greedy else SpecTree

    def new_tree(prompt_dev):
        return cls(prefix=prompt_dev, device=dev, temperature=T, top_p=top_p, draft_kv_len=0, target_kv_len=0,
                   draft_model_engine=draft, target_model_engine=target, max_length=M, max_target_seq=M,
                   grow_map=grow_map, **buf)

    class Loop:
        """tests/testbed.py:45-95 (simulation_fast) with a step budget; timing by CUDA events around each prompt's
        while-loop (construction / prefill excluded exactly as :78-91 does)."""

        def __init__(self):
            self.pi = 0
            self.tree = None
            self.len = 0
            self.terminate = True

        def next_prompt(self, host=False):
            if self.tree is not None:
                draft.clear_kv()
                target.clear_kv()
            p = prompts[self.pi % len(prompts)]
            self.pi += 1
            if host:
                self.h2d += p.numel() * 8 + M * 2 + S * 32000 * 2 + M * 8 + 64
                p = pinned_prompts[(self.pi - 1) % len(prompts)].to(dev, non_blocking=True)
            else:
                p = p.to(dev)
            self.tree = new_tree(p)
            self.len = prefix
            self.terminate = False

        h2d = 0
        d2h = 0

        def run_steps(self, k, timed, host=False):
            done = tokens = 0
            ms = 0.0
            while done < k:
                if self.terminate or self.len >= max_len:
                    if host and self.tree is not None:
                        _ = self.tree.tokens[:self.len].to("cpu")        # D2H of the finished sequence
                        self.d2h += self.len * 8
                    self.next_prompt(host)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                while done < k and self.len < max_len and not self.terminate:
                    self.tree.construct_grow_map()
                    valid, _, _, self.terminate = self.tree.verify()
                    tokens += valid.shape[0] - self.len
                    self.len = valid.shape[0]
                    if int(self.tree.rt.host_state[5]) in (0, 2):       # bonus token is EOS / pad (testbed.py:87)
                        self.terminate = True
                    done += 1
                    self.d2h += 64
                e1.record()
                e1.synchronize()
                ms += e0.elapsed_time(e1)
            return tokens, ms

    loop = Loop()
    # warm-up: captures the graphs (first prompt) and W untimed steps
    loop.run_steps(max(args.warmup, 3), timed=False)
    lc0 = loop.tree.rt.kernel_launches()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    t_wall0 = time.time()
    tokens, ms = loop.run_steps(args.steps, timed=True)
    barrier()
    wall = time.time() - t_wall0
    clocks = sampler.stop()
    launches = loop.tree.rt.kernel_launches() - lc0
    value = tokens / (ms / 1e3)
    acc_per_step = tokens / args.steps

    # ---- e2e: same metric through the public API from pinned host buffers -------------------------------------------
    pinned_prompts = [p.pin_memory() for p in prompts]
    loop2 = Loop()
    loop2.pi = 1000
    loop2.tree = loop.tree
    loop2.h2d = loop2.d2h = 0
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    tok2, _ = loop2.run_steps(args.steps, timed=True, host=True)
    _ = loop2.tree.tokens[:loop2.len].to("cpu")
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    e2e = {"value": tok2 / (e2e_ms / 1e3), "unit": "tokens/s", "h2d_bytes_per_step": loop2.h2d // args.steps,
           "d2h_bytes_per_step": loop2.d2h // args.steps,
           "note": "includes per-prompt Tree construction (CPU-drawn r/rand as in the reference), prefill, decode"}

    # ---- roofline of the verify tree-attention kernel, measured live (CUDA events on the launching stream) ----------
    roof = extra = None
    if not args.no_micro:
        roof = attention_roofline(target, grow_map, prefix, M)
        extra = micro_kernels(draft, target, loop.tree, grow_map)
    draft.clear_kv()
    target.clear_kv()
    if world > 1:
        from sequoia_b200.tp import stop_followers
        stop_followers(tp_group, dev)
    out = {
        "metric": METRIC, "value": round(value, 2), "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f16", "data": "synthetic (random-init weights, random prompts)",
        "config": {"workload": f"{args.config}: {dname}->{tname}, {os.path.basename(gm_path)} (tree {S}), "
                               f"{'greedy' if greedy else 'stochastic'} T={T} P={top_p} M={M}, prefix {prefix}->{max_len} tokens",
                   "accepted_tokens_per_step": round(acc_per_step, 4), "parallelism": f"target tp{world}, draft on rank 0",
                   "l2": "inputs larger than L2: each step streams the target's %.1f GB of weights" % (target.engine.runner.weight_bytes() / 1e9)},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "kernels": extra,
        "wall_s_timed_region": round(wall, 3),
    }
    if not args.no_cpu_baseline and world == 1:
        out["cpu_baseline"] = cpu_reference(args.config, max_seconds=25.0, max_iters=3)
    print(json.dumps(out))
    finish()


def _timeit(fn, iters=20, warm=3, reps=5):
    """Average device time (us) of one fn() call: `iters` calls are captured into a CUDA graph (so host launch overhead
    is excluded, as inside the real decode graphs) and the graph is replayed `reps` times between CUDA events."""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(warm):
            fn()
        s.synchronize()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / (iters * reps) * 1e3   # us


def attention_roofline(target, grow_map, prefix, M):
    """Verify attention (Engine/Llama_modules.py:220-248) of the steady-state shape: q = S tree rows, kv = P-1+S with
    P = (prefix + max_len)/2-ish mid-decode value; cycles through all layers so K/V come from HBM (cache >> L2 at 7B)."""
    from sequoia_b200 import ops
    from sequoia_b200.tree import pack_tree_mask
    rn = target.engine.runner
    S = grow_map["size"]
    P = 193                                    # kv = 320 as in SURVEY.md 8(d) for config 2
    kv = P - 1 + S
    bits = pack_tree_mask(grow_map["mask"]).to(rn.device)
    state = torch.zeros(16, dtype=torch.int32, device=rn.device)
    state[0] = P
    rn.qkv.normal_(0, 1)
    rn.k_cache.normal_(0, 1)
    rn.v_cache.normal_(0, 1)
    layer = [0]

    def call():
        ops.tree_attn(rn.plan, layer[0] % rn.L, S, state=state, n0=0, kv_end=S, tree_bits=bits, tree_words=bits.shape[1],
                      tree_size=S, impl=0)
        layer[0] += 1

    us = _timeit(call, iters=4 * rn.L, warm=rn.L)
    D = rn.D
    alg_bytes = 2 * D * 2 * (rn.Hkv * kv + rn.H * S)           # K+V read once, Q read + O write
    flops = 4 * rn.H * S * kv * D
    peak, how = load_peaks()
    traffic = None
    tp = os.path.join(ROOT, "profiles", "attn_traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    rn.k_cache.zero_()
    rn.v_cache.zero_()
    return {"kernel": "tree_attn_tc_kernel<128,false> (verify attention incl. its in-cluster split-KV reduction, q=%d kv=%d H=%d)" % (S, kv, rn.H),
            "bound": "hbm", "achieved": round(alg_bytes / (us * 1e-6) / 1e9, 1), "peak": peak, "unit": "GB/s",
            "frac": round(alg_bytes / (us * 1e-6) / 1e9 / peak, 4), "traffic": traffic, "peak_source": how,
            "algorithmic_bytes": alg_bytes, "us_per_launch": round(us, 3), "tflops": round(flops / (us * 1e-6) / 1e12, 2)}


def micro_kernels(draft, target, tree, grow_map):
    """Per-kernel device times (us) of the other hot-path kernels at this config's shapes, for DESIGN.md's table."""
    from sequoia_b200 import ops
    rt = tree.rt
    S = grow_map["size"]
    out = {}
    lv = max(range(len(rt.st.levels)), key=lambda i: rt.st.levels[i]["n_parents"])
    rows = rt.st.levels[lv]["n_parents"]
    us = _timeit(lambda: rt.op_sample(lv))
    out["sample_level"] = {"us": round(us, 2), "rows": rows, "GBps": round(rows * 32000 * 4 / us / 1e3, 1)}
    snap_t, snap_p, snap_s = rt.tokens.clone(), rt.position_ids.clone(), rt.state.clone()

    def acc():
        rt.state.copy_(snap_s)
        rt.op_accept()
    us = _timeit(acc)
    out["accept_walk(+state copy)"] = {"us": round(us, 2)}
    rt.tokens.copy_(snap_t); rt.position_ids.copy_(snap_p); rt.state.copy_(snap_s)
    st = torch.zeros(16, dtype=torch.int32, device=rt.device)
    st[3], st[4] = 5, 150
    idx = torch.tensor([160, 170, 180, 190, 200, 0, 0, 0], dtype=torch.int32, device=rt.device)
    kvc = target.engine.kv_cache
    us = _timeit(lambda: kvc.gather_from_state(idx, st, 8))
    L, _, Hkv, _, D = kvc.k_cache.shape
    out["kv_gather(target,n=5)"] = {"us": round(us, 2), "GBps": round(2 * 2 * L * Hkv * 5 * D * 2 / us / 1e3, 1)}
    rn = target.engine.runner
    us = _timeit(lambda: ops.add_rmsnorm(rn.hidden, rn.proj, rn.norm, rn.normed, S, rn.eps))
    out["add_rmsnorm(S rows)"] = {"us": round(us, 2), "GBps": round(S * rn.h * 2 * 4 / us / 1e3, 1)}
    return out


# ----------------------------------------------------------------------------------------------------------------------
def pick_cpu_threads():
    """torch's fp16 CPU GEMM does not scale to every core count (128 threads were 10x slower than 8 on the build box):
    try powers of two up to os.cpu_count() on one 7B-shaped linear and keep the fastest, i.e. all the threads the
    reference's CPU path can actually use."""
    n = os.cpu_count() or 1
    x = torch.randn(128, 4096).half()
    w = torch.randn(4096, 4096).half()
    best, best_t = 1, float("inf")
    c = 1
    cands = []
    while c < n:
        cands.append(c)
        c *= 2
    cands.append(n)
    for c in cands[-5:]:
        torch.set_num_threads(c)
        torch.nn.functional.linear(x, w)
        t0 = time.time()
        for _ in range(3):
            torch.nn.functional.linear(x, w)
        dt = time.time() - t0
        if dt < best_t:
            best, best_t = c, dt
    return best


def cpu_reference(config, max_seconds, max_iters, warm_iters=1):
    """The reference's algorithm on the host cores: the torch-CPU oracle port (oracle/sequoia_oracle.py) on the same
    shapes.  To keep host init bounded, all target layers alias ONE layer's random weights (identical FLOPs/bytes per
    layer; the working set still exceeds the caches).  Returns tokens/s over the timed iterations."""
    from oracle import sequoia_oracle as O
    from sequoia_b200.model import NAMED_CONFIGS
    dname, tname, gm_path, greedy, T, top_p, M, prefix, max_len = CONFIGS[config]
    ncores = pick_cpu_threads()
    torch.set_num_threads(ncores)
    grow_map = torch.load(os.path.join(ROOT, gm_path))

    def shared_weights(name, seed):
        c = NAMED_CONFIGS[name]
        cfg = O.LlamaCfg(c.hidden_size, c.intermediate_size, c.num_hidden_layers, c.num_attention_heads,
                         c.num_key_value_heads, c.vocab_size, c.rms_norm_eps, c.rope_theta, c.max_position_embeddings)
        one = O.LlamaCfg(c.hidden_size, c.intermediate_size, 1, c.num_attention_heads, c.num_key_value_heads,
                         c.vocab_size, c.rms_norm_eps, c.rope_theta, c.max_position_embeddings)
        w1 = O.init_llama_weights(one, seed)
        w = dict(w1)
        for l in range(1, c.num_hidden_layers):
            for k, v in w1.items():
                if k.startswith("model.layers.0."):
                    w[k.replace("model.layers.0.", f"model.layers.{l}.")] = v
        return cfg, w

    t0 = time.time()
    dcfg, dw = shared_weights(dname, 1)
    tcfg, tw = shared_weights(tname, 2)
    draft = O.EngineOracle(O.LlamaOracle(dcfg, dw, M, "FI"))
    target = O.EngineOracle(O.LlamaOracle(tcfg, tw, M, "TG"))
    torch.manual_seed(17)
    prompt = synthetic_prompts(1, prefix)[0]
    tree = (O.GreedyTreeOracle(draft, target, prompt, grow_map, max_length=M) if greedy else
            O.SpecTreeOracle(draft, target, prompt, grow_map, temperature=T, top_p=top_p, max_length=M))
    length = prefix
    for _ in range(warm_iters):                      # first iteration = prefill of the target (untimed, like the GPU arm)
        tree.construct_grow_map()
        valid, _, _, term = tree.verify()
        length = valid.shape[0]
    init_s = time.time() - t0
    iters = tokens = 0
    t1 = time.time()
    while iters < max_iters and (time.time() - t1) < max_seconds and not term and length < max_len:
        tree.construct_grow_map()
        valid, _, _, term = tree.verify()
        tokens += valid.shape[0] - length
        length = valid.shape[0]
        iters += 1
    dt = time.time() - t1
    return {"value": round(tokens / dt, 4) if dt > 0 and iters else None, "unit": "tokens/s", "cores": ncores,
            "kind": "port", "ms_per_step": round(dt / max(iters, 1) * 1e3, 1), "steps_timed": iters,
            "accepted_tokens_per_step": round(tokens / max(iters, 1), 3),
            "sample": f"{iters} steady decode iteration(s) of the torch-CPU oracle (fp16, {ncores} threads) on the same "
                      f"shapes/growmap after 1 untimed prefill iteration; target layers alias one layer's weights "
                      f"(init {init_s:.0f}s)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    dname, tname, gm_path, greedy, T, top_p, M, prefix, max_len = CONFIGS[args.config]
    S = torch.load(os.path.join(ROOT, gm_path))["size"]
    cb = cpu_reference(args.config, max_seconds=150.0, max_iters=max(args.steps, 1), warm_iters=1)
    out = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "tokens/s",
           "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "f16", "data": "synthetic (random-init weights, random prompts)",
           "config": {"workload": f"{args.config}: {dname}->{tname}, {os.path.basename(gm_path)} (tree {S}), "
                                  f"{'greedy' if greedy else 'stochastic'} T={T} P={top_p} M={M}, prefix {prefix}->{max_len} tokens",
                      "accepted_tokens_per_step": cb["accepted_tokens_per_step"], "parallelism": "host CPU"},
           "cpu_baseline": cb, "gpu_launches": 0,
           "e2e": {"value": cb["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=list(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-micro", action="store_true", help="skip the per-kernel micro timings (for ncu launch lists)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
