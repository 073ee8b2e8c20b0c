"""Measure the two inputs of the growmap search on this GPU: the draft model's time per tree level and the target's
verify time per tree budget (reference `tree_search.py` config keys `draft_time`, `valid_budget`, `target_time`).

    python tools/measure_tree_times.py --draft random-init:llama-68m --target random-init:llama-2-7b \
        --acceptance acceptance-rate-vector.pt --out gpurun_out/b200_68m_7b.json
    python tree_search.py --config gpurun_out/b200_68m_7b.json

Times are device times of the captured forward (CUDA graph replay between CUDA events), prefix 128 tokens resident.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def time_forward(engine, n, prefix, tg, iters=8, reps=4):
    rn = engine.runner
    dev = rn.device
    ids = torch.randint(0, rn.cfg.vocab_size, (n,), device=dev)
    pos = torch.arange(prefix, prefix + n, device=dev)
    sto = torch.arange(prefix, prefix + n, device=dev)
    kv_end = prefix + n if tg else engine.max_length
    mask = torch.zeros(n, kv_end, dtype=torch.float16, device=dev)
    mask[:, prefix:prefix + n] = torch.triu(torch.full((n, n), torch.finfo(torch.float16).min, dtype=torch.float16, device=dev), 1)
    if not tg:
        mask[:, prefix + n:] = torch.finfo(torch.float16).min

    def fn():
        rn.forward(n, ids, pos, sto, kv_end=kv_end, dense_mask=mask, mask_ld=mask.stride(0))
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s), torch.inference_mode():
        for _ in range(3):
            fn()
        s.synchronize()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.inference_mode(), torch.cuda.graph(g):
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
    return e0.elapsed_time(e1) / (iters * reps)          # ms


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--draft", default="random-init:llama-68m")
    ap.add_argument("--target", default="random-init:llama-2-7b")
    ap.add_argument("--budgets", default="1,2,4,8,16,32,64,128,256,512,768")
    ap.add_argument("--draft-width", type=int, default=32, help="rows per draft level used for draft_time")
    ap.add_argument("--prefix", type=int, default=128)
    ap.add_argument("--max-depth", type=int, default=24)
    ap.add_argument("--acceptance", default="acceptance-rate-vector.pt")
    ap.add_argument("--dst", default="B200_growmap.pt")
    ap.add_argument("--out", default="gpurun_out/tree_times.json")
    a = ap.parse_args()
    from sequoia_b200.engine import InferenceEngine, InferenceEngineTG
    budgets = [int(x) for x in a.budgets.split(",")]
    M = a.prefix + max(budgets) + 8
    draft = InferenceEngine(M, a.draft, device="cuda:0")
    widths = sorted({1, 8, a.draft_width, 64, 128})
    dt = {w: time_forward(draft, w, a.prefix, tg=False) for w in widths}
    del draft
    target = InferenceEngineTG(M, a.target, device="cuda:0")
    tt = [time_forward(target, b, a.prefix, tg=True) for b in budgets]
    cfg = {"acceptance_rate_vector": a.acceptance, "max_depth": a.max_depth, "max_budget": max(budgets),
           "draft_time": round(dt[a.draft_width], 4), "valid_budget": budgets, "target_time": [round(t, 4) for t in tt],
           "dst": a.dst,
           "_measured": {"gpu": torch.cuda.get_device_name(0), "draft": a.draft, "target": a.target, "prefix": a.prefix,
                         "unit": "ms", "draft_time_by_width": {str(k): round(v, 4) for k, v in dt.items()}}}
    os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(cfg, f, indent=1)
    print(json.dumps(cfg))


if __name__ == "__main__":
    main()
