"""Growmap search (sequoia_b200/tree_search.py) against growmaps produced by the reference's own
tree_search.py (tests/golden/make_tree_search_golden.py).  CPU only."""
import json
import os
import subprocess
import sys
import time

import numpy as np
import pytest
import torch

from sequoia_b200 import tree_search as ts

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = torch.load(os.path.join(HERE, "golden", "tree_search_golden.pt"), map_location="cpu")
P = GOLD["acceptance_rate_vector"][:-1].numpy()


def _same_growmap(a, b):
    assert a["size"] == b["size"]
    assert a["roots"] == b["roots"]
    assert a["branches"] == b["branches"]
    assert a["Successors"] == b["Successors"]
    assert torch.equal(a["mask"], b["mask"]) and a["mask"].dtype == b["mask"].dtype
    assert torch.equal(a["depth"], b["depth"]) and a["depth"].dtype == b["depth"].dtype


@pytest.mark.parametrize("name", sorted(GOLD["cases"]))
def test_matches_reference_growmap(name):
    case = GOLD["cases"][name]
    cfg = case["config"]
    table = ts.search(P, cfg["max_depth"], cfg["max_budget"])
    _, pair = ts.choose_budget_depth(table, cfg["draft_time"], cfg["target_time"], cfg["valid_budget"])
    _same_growmap(ts.build_grow_map(table, *pair), case["grow_map"])


def test_value_table_against_scalar_recurrence():
    """Vectorised DP == a direct scalar float32 evaluation of the recurrence (small sizes)."""
    p = P[:7].copy()
    M, L, B = 14, 5, 6
    F = np.full((M + 1, L + 1, B + 1), -np.inf, dtype=np.float32)
    F[1, 1:, 0] = 1
    for m in range(2, M + 1):
        for l in range(2, L + 1):
            F[m, l, 1] = np.float32(1) + p[1] * F[m - 1, l - 1].max()
            for b in range(2, B + 1):
                best = np.float32(-np.inf)
                for y in range(1, m):
                    with np.errstate(invalid="ignore"):
                        v = F[y, l, b - 1] + p[b] * F[m - y, l - 1].max()
                    if v > best:
                        best = v
                F[m, l, b] = best
    got = ts.search(p, L, M).F
    assert np.array_equal(got, F)


def test_growmap_invariants_and_expected_value():
    table = ts.search(P, 8, 64)
    g = ts.build_grow_map(table, 64, 8)
    n = g["size"]
    assert n == 64 and sum(len(r) for r in g["roots"]) == n and g["mask"].sum(1).tolist() == (g["depth"] + 1).tolist()
    # expected accepted tokens recomputed from the tree == the DP value
    def ev(i):
        return 1.0 + sum(float(P[j + 1]) * ev(c) for j, c in enumerate(g["Successors"][i]))
    assert abs(ev(0) - float(table.best[64, 8])) < 1e-4
    # monotone in budget and depth
    best = table.best
    assert np.all(np.diff(best[1:, 8]) >= -1e-6)


def test_cli_writes_reference_format(tmp_path):
    p_path = tmp_path / "p.pt"
    torch.save(GOLD["acceptance_rate_vector"], p_path)
    cfg = dict(GOLD["cases"]["flat64"]["config"], acceptance_rate_vector=str(p_path), dst=str(tmp_path / "t.pt"))
    (tmp_path / "c.json").write_text(json.dumps(cfg))
    root = os.path.dirname(HERE)
    r = subprocess.run([sys.executable, os.path.join(root, "tree_search.py"), "--config", str(tmp_path / "c.json")],
                       capture_output=True, text=True, cwd=root)
    assert r.returncode == 0, r.stderr
    info = json.loads(r.stdout.strip().splitlines()[-1])
    assert info["budget"] == 64
    _same_growmap(torch.load(cfg["dst"]), GOLD["cases"]["flat64"]["grow_map"])


def test_large_search_is_fast():
    t0 = time.time()
    table = ts.search(P, 16, 256)
    assert time.time() - t0 < 60
    g = ts.build_grow_map(table, 256, 16)
    assert g["size"] == 256


def test_searched_growmaps_load_into_the_runtime_tables():
    """What tree_search produces is what the decode runtime consumes: `_Static` (CSR successors, per-level parents /
    first-child / branch tables, packed ancestor bits) accepts every searched growmap, the shipped B200 one and the
    star growmap of the acceptance-rate trees; the packed bits reproduce the mask."""
    from sequoia_b200.tree import _Static, pack_tree_mask, star_grow_map
    root = os.path.dirname(HERE)
    maps = [c["grow_map"] for c in GOLD["cases"].values()]
    maps.append(torch.load(os.path.join(root, "B200_growmaps", "68m_7b-demo_acceptance.pt")))
    maps.append(star_grow_map(32))
    for gm in maps:
        st = _Static(gm, "cpu")
        S = gm["size"]
        assert st.S == S and st.succ_off.numel() == S + 1 and int(st.succ_off[-1]) == S - 1
        assert sum(lv["tb"] for lv in st.levels) == S - 1
        bits = pack_tree_mask(gm["mask"])
        for r in range(S):
            row = [(int(bits[r, c // 32]) >> (c % 32)) & 1 for c in range(S)]
            assert row == gm["mask"][r].tolist()
        assert st.max_depth == int(gm["depth"].max())
