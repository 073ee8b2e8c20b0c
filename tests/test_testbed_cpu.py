"""Host-side checks of the reference-CLI driver (testbed.py): flags of tests/testbed.py:21-33, prompt loading."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import testbed  # noqa: E402


def test_reference_flags_are_accepted():
    a = testbed.build_parser().parse_args(["--model", "d", "--target", "t", "--dataset", "x.json", "--growmap", "g.pt",
                                           "--start", "3", "--end", "9", "--T", "0.7", "--P", "0.95", "--M", "512",
                                           "--seed", "5", "--Mode", "benchmark", "--offloading"])
    assert (a.model, a.target, a.dataset, a.growmap, a.start, a.end) == ("d", "t", "x.json", "g.pt", 3, 9)
    assert (a.T, a.P, a.M, a.seed, a.Mode, a.offloading) == (0.7, 0.95, 512, 5, "benchmark", True)
    for mode in ("greedy", "benchmark", "baseline"):
        assert testbed.build_parser().parse_args(["--Mode", mode]).Mode == mode


def test_prompt_sources(tmp_path):
    syn = testbed.load_prompts("synthetic", 2, 6, seed=17)
    assert len(syn) == 4 and all(p.shape == (128,) and p.dtype == torch.long and int(p.min()) >= 3 for p in syn)
    again = testbed.load_prompts("synthetic", 0, 6, seed=17)
    assert torch.equal(again[2], syn[0])                          # deterministic, sliced like the reference's select()
    rows = [{"input_tokens": list(range(5, 205))}, {"input_tokens": list(range(7, 60))}, {"input_ids": list(range(9, 300))}]
    p1 = tmp_path / "a.jsonl"
    p1.write_text("\n".join(json.dumps(r) for r in rows))
    got = testbed.load_prompts(str(p1), 0, 3, seed=0)
    assert len(got) == 2                                          # the short (padded) row is skipped
    assert got[0].tolist() == list(range(5, 133)) and got[1].tolist() == list(range(9, 137))
    p2 = tmp_path / "b.json"
    p2.write_text(json.dumps([list(range(3, 140)), list(range(4, 150))]))
    got = testbed.load_prompts(str(p2), 1, 2, seed=0)
    assert len(got) == 1 and got[0][0] == 4
