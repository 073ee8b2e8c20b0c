"""Generate tests/golden/tree_search_golden.pt by running the UNMODIFIED reference script
(/root/reference/tree_search.py) as a subprocess on a few timing configs.

Run in the build container only:   python tests/golden/make_tree_search_golden.py
The reference's acceptance-rate vector was saved from a CUDA tensor and its script loads it without
map_location, so the script is handed a CPU re-save of the same values (input data, not code).
"""
import json
import os
import subprocess
import sys
import tempfile

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"

CONFIGS = {
    # the reference's own demo (demo-config.json) — its committed demo_tree.pt is the expected output
    "demo": dict(max_depth=10, max_budget=128, draft_time=0.3, valid_budget=[1, 2, 4, 8, 16, 32],
                 target_time=[10, 10, 10, 12, 14, 18]),
    # cheap draft, flat target cost: the search should spend the whole budget
    "flat64": dict(max_depth=8, max_budget=64, draft_time=0.05, valid_budget=[16, 32, 64], target_time=[10, 10, 10]),
    "deep48": dict(max_depth=12, max_budget=48, draft_time=0.01, valid_budget=[8, 24, 48], target_time=[5, 5.5, 6]),
    "wide96": dict(max_depth=5, max_budget=96, draft_time=0.2, valid_budget=[32, 96], target_time=[20, 21]),
}


def main():
    p = torch.load(os.path.join(REF, "acceptance-rate-vector.pt"), map_location="cpu").float().cpu()
    out = {"acceptance_rate_vector": p.clone(), "cases": {}}
    with tempfile.TemporaryDirectory() as tmp:
        torch.save(p, os.path.join(tmp, "p.pt"))
        for name, cfg in CONFIGS.items():
            cfg = dict(cfg, acceptance_rate_vector=os.path.join(tmp, "p.pt"), dst=os.path.join(tmp, name + ".pt"))
            with open(os.path.join(tmp, name + ".json"), "w") as f:
                json.dump(cfg, f)
            r = subprocess.run([sys.executable, os.path.join(REF, "tree_search.py"), "--config",
                                os.path.join(tmp, name + ".json")], cwd=tmp, capture_output=True, text=True)
            assert r.returncode == 0, r.stderr[-2000:]
            g = torch.load(cfg["dst"], map_location="cpu")
            keep = {k: v for k, v in cfg.items() if k not in ("acceptance_rate_vector", "dst")}
            out["cases"][name] = {"config": keep, "grow_map": g}
            print(name, "size", g["size"], "levels", len(g["roots"]), "root branches", g["branches"][0])
    demo = torch.load(os.path.join(REF, "demo_tree.pt"), map_location="cpu")
    got = out["cases"]["demo"]["grow_map"]
    assert got["size"] == demo["size"] and got["branches"] == demo["branches"] and torch.equal(got["mask"], demo["mask"])
    torch.save(out, os.path.join(HERE, "tree_search_golden.pt"))


if __name__ == "__main__":
    main()
