"""Target tensor parallelism on real GPUs (needs >= 2 devices; skipped otherwise): a TP-2 decode driven through
sequoia_b200.tp (NCCL broadcasts + 2 allreduces per layer, captured in CUDA graphs) must reproduce the single-GPU
decode of the same models: target logits within fp16 allreduce-order noise, identical accept lengths / tokens on the
first iterations."""
import os
import socket
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q, variant="tiny"):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import cases
    from sequoia_b200.engine import GraphInferenceEngine, GraphInferenceEngineTG
    from sequoia_b200.tp import TPFollower, attach_tp, stop_followers
    from sequoia_b200.tree import SpecTree
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = f"cuda:{rank}"
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(dev))
    grp = dist.group.WORLD
    dcfg, dw = cases.model_weights("draft")
    if variant == "tiny":
        gm = cases.load_growmap("L40_growmaps/8x8-tree.pt")
        M, plen, iters = 256, 100, 4
        tcfg, tw = cases.model_weights("target_gqa")               # H=4, Hkv=2 -> shards of 2 q heads + 1 kv head
        tspec = {"config": tcfg, "state_dict": tw}
    else:
        # config 4's REAL target shapes (h=8192, I=28672, 64 q heads on 8 kv heads, 8 layers' worth) under the 768-node
        # tree at M=1024: per rank 32 q heads / 4 kv heads, 6 query tiles x 8 KV tiles in the attention kernel
        gm = cases.load_growmap("L40_growmaps/L40-CNN-7b-70b-stochastic.pt")
        M, plen, iters = 1024, 128, 3
        tspec = "random-init:llama-2-70b-8l:2"
    target_tp = GraphInferenceEngineTG(M, tspec, device=dev, tp_group=grp)
    if rank != 0:
        TPFollower(target_tp, gm, False, M, dev, grp).serve()
        os._exit(0)
    try:
        draft = GraphInferenceEngine(M, {"config": dcfg, "state_dict": dw}, device=dev)
        draft2 = GraphInferenceEngine(M, {"config": dcfg, "state_dict": dw}, device=dev)
        target_1 = GraphInferenceEngineTG(M, tspec, device=dev)
        attach_tp(draft, target_tp, grp)
        prompt = cases.make_prompt(25, plen).to(dev)
        buf = lambda: dict(attn_mask=torch.full((M, M), torch.finfo(torch.float16).min, dtype=torch.float16, device=dev),
                           sequence=torch.arange(M, device=dev).unsqueeze(-1), new_tokens_buffer=torch.zeros(M, device=dev).long(),
                           parents_buffer=torch.zeros(M, device=dev).long(), position_ids=torch.zeros(M, device=dev).long())
        noise = torch.empty(iters, cases.V, dtype=torch.float16).exponential_(1.0, generator=torch.Generator().manual_seed(5)).to(dev)
        trees = []
        for d, t in ((draft, target_tp), (draft2, target_1)):
            torch.manual_seed(17)
            tr = SpecTree(prefix=prompt, device=dev, temperature=0.6, top_p=1.0, draft_model_engine=d, target_model_engine=t,
                          max_length=M, max_target_seq=M, grow_map=gm, **buf())
            tr.rt.external_noise = noise
            trees.append(tr)
        res = []
        for it in range(iters):
            outs = []
            for tr in trees:
                tr.construct_grow_map()
                v, a, _, term = tr.verify()
                outs.append((v.clone(), a, term, tr.rt.target_logits.float().clone()))
            (v0, a0, t0, l0), (v1, a1, t1, l1) = outs
            rel = ((l0 - l1).abs().max() / l1.abs().max()).item()
            res.append((it, a0 == a1 and t0 == t1 and torch.equal(v0, v1), rel))
            if not res[-1][1]:
                break
        res.append(("peer_error", target_tp.engine.runner.peer.error() if target_tp.engine.runner.peer else 0, 0.0))
        q.put(res)
        q.close()
        q.join_thread()
    finally:
        stop_followers(grp, dev)
        os._exit(0)                                  # skip the slow NCCL / CUDA-graph teardown


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("variant,shot", [("tiny", "1"), ("tiny", "2"), ("tiny", "3"), ("tiny", "4"), ("70b-8l", "2"), ("70b-8l", "4")])
def test_tp2_decode_matches_single_gpu(variant, shot):
    """shot = fused all-reduce flavour (csrc/sq_tp.cu): 1 = one-shot pull, 2 = two-shot reduce-scatter + all-gather in one
    kernel, 4 = LL two-shot (data + epoch in one 8-byte store, readers poll; the default for small payloads),
    3 = one-shot push for small payloads (opt-in; larger ones fall back to pull / two-shot, as the 768-row
    verify of the 70B-shaped variant does) -- all must work at any N."""
    import torch.multiprocessing as mp
    os.environ["SQ_TP_SHOT"] = shot                   # inherited by the spawned ranks
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, variant)) for r in range(2)]
    [p.start() for p in procs]
    res = q.get(timeout=600)
    [p.join(60) for p in procs]
    [p.kill() for p in procs if p.is_alive()]
    assert len(res) >= 2
    tag, err, _ = res.pop()
    assert tag == "peer_error" and err == 0, "fused all-reduce handshake timed out"
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "tp_parity_test.log"), "a") as f:
        f.write(f"{variant} shot={shot}: peer_error={err} " + " ".join(f"[iter {it} same={same} rel={rel:.3e}]" for it, same, rel in res) + "\n")
    # logits are comparable while both runs hold the same token tree (up to and including the first differing iteration)
    # The row-parallel partial products cross NVLink as fp16 (what an NCCL fp16 all-reduce does too), one extra rounding
    # per reduction that the unsharded model does not have: measured 1.1e-3 on the 3-layer toy, 3.7e-3 on the 8-layer
    # 70B-shaped model (h=8192) -- with IDENTICAL accept sequences in both.
    tol = 2e-3 if variant == "tiny" else 6e-3
    for it, same, rel in res:
        assert rel < tol, f"iter {it}: TP-2 target logits differ from TP-1 by {rel} (relative to max |logit|)"
        if not same:
            break
    assert res[0][1], "first iteration: TP-2 accept length / tokens differ from single GPU"
    assert sum(1 for r in res if r[1]) >= 2, f"TP-2 decode forked too early: {res}"
