"""Acceptance-rate vector for the growmap search (reference tests/test_accept.py:36-140, `--dst`): per decode step draft
`W` children of the last token, verify with the target, histogram WHICH child rank was accepted.  Output format is the
reference's: a (W+2,) float tensor [0, p_rank1 .. p_rankW, p_none] (tree_search.py drops the last entry).

    python tools/measure_acceptance.py --model <draft dir|random-init:llama-68m> --target <...> --W 32 --dst acc.pt

Prompts: `--dataset` = a JSON list of token-id lists (e.g. the reference's dataset/c4_small.json) or synthetic random ids.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def simulate(draft, target, prompts, T, top_p, W, M, greedy, max_new_len=256):
    from Tree.GreedyTree import GreedyTreeTest
    from Tree.SpecTree import SpecTreeTest
    dev = "cuda:0"
    dtype = torch.float16
    attn_mask = torch.full((M, M), torch.finfo(dtype).min, dtype=dtype, device=dev)
    sequence = torch.arange(M, device=dev).long().unsqueeze(-1)
    new_tokens_buffer = torch.zeros(M, device=dev).long()
    parents_buffer = torch.zeros(M, device=dev).long()
    position_ids = torch.zeros(M, device=dev).long()
    branch = torch.zeros(W + 1)
    steps = decoded = 0
    cls = GreedyTreeTest if greedy else SpecTreeTest
    for prompt in prompts:
        input_ids = prompt.view(1, -1).to(dev)
        dkv = tkv = 0
        terminate = False
        while input_ids.shape[1] < max_new_len and not terminate:
            tree = cls(prefix=input_ids.squeeze(0), device=dev, temperature=T, top_p=top_p, draft_kv_len=dkv,
                       target_kv_len=tkv, draft_model_engine=draft, target_model_engine=target, max_length=M,
                       attn_mask=attn_mask, sequence=sequence, new_tokens_buffer=new_tokens_buffer,
                       parents_buffer=parents_buffer, position_ids=position_ids, max_width=W)
            valid, dkv, tkv, b, terminate = tree.verify(benchmark=True)
            n0 = input_ids.shape[1]
            input_ids = valid.clone().unsqueeze(0)
            if bool(((input_ids[0] == 2) | (input_ids[0] == 0)).any()):
                terminate = True
            if greedy and terminate:                   # tests/test_accept.py:118-121 counts only non-terminal greedy steps
                continue
            branch[b] += 1                             # b = -1 -> last slot ("none accepted")
            decoded += valid.shape[0] - n0
            steps += 1
        draft.clear_kv()
        target.clear_kv()
    out = torch.zeros(W + 2)
    out[1:] = branch / branch.sum().clamp(min=1)
    return out, decoded / max(steps, 1), steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="random-init:llama-68m")
    ap.add_argument("--target", default="random-init:llama-68m")
    ap.add_argument("--dataset", default=None)
    ap.add_argument("--start", type=int, default=0)
    ap.add_argument("--end", type=int, default=20)
    ap.add_argument("--T", type=float, default=0.6)
    ap.add_argument("--P", type=float, default=1.0)
    ap.add_argument("--W", type=int, default=32)
    ap.add_argument("--M", type=int, default=384)
    ap.add_argument("--Mode", default="stochastic", choices=["stochastic", "greedy"])
    ap.add_argument("--dst", default="gpurun_out/acceptance-rate-vector.pt")
    a = ap.parse_args()
    from Engine.Engine import GraphInferenceEngine, GraphInferenceEngineTG
    torch.manual_seed(17)
    draft = GraphInferenceEngine(a.M, a.model, device="cuda:0")
    target = GraphInferenceEngineTG(a.M, a.target, device="cuda:0")
    if a.dataset:
        with open(a.dataset) as f:
            rows = json.load(f)
        prompts = [torch.tensor(r[:128], dtype=torch.long) for r in rows[a.start:a.end]]
    else:
        g = torch.Generator().manual_seed(17)
        prompts = [torch.randint(3, 32000, (128,), generator=g) for _ in range(a.end - a.start)]
    vec, tok_per_step, steps = simulate(draft, target, prompts, a.T, a.P, a.W, a.M, a.Mode == "greedy")
    os.makedirs(os.path.dirname(os.path.abspath(a.dst)), exist_ok=True)
    torch.save(vec, a.dst)
    print(json.dumps({"steps": steps, "tokens_per_step": round(tok_per_step, 4), "vector": [round(float(x), 4) for x in vec]}))


if __name__ == "__main__":
    main()
