"""Offline growmap search: the dynamic programme that picks the tree shape SpecTree/GreedyTree replay.

Host-side, CPU-only tooling (not on the decode hot path).  Behaviour follows the reference's
`tree_search.py` (whole file; DP at :21-52, budget/depth choice at :56-74, level-order expansion at
:76-120) so that a growmap produced here is the growmap the reference would have produced from the
same acceptance-rate vector and timing config.

State: F[m, l, b] = expected number of accepted tokens of the best tree with exactly m nodes, at most
l levels, whose root has exactly b children (children tried in draft-rank order, child j accepted with
probability p[j]).  Recurrence (reference :33-50):

    F[1, l, 0] = 1
    F[m, l, 1] = 1 + p[1] * G[m-1, l-1]
    F[m, l, b] = max_y  F[y, l, b-1] + p[b] * G[m-y, l-1]          (1 <= y < m, first maximiser wins)
    G[m, l]    = max_b F[m, l, b]

All arithmetic is float32, one rounding per multiply and per add, and ties break toward the smallest
index, which is what the reference's torch scalar code does; the tests compare against growmaps the
reference script itself generated (tests/golden/tree_search_golden.pt).

Implementation: instead of a dict of deep-copied child lists per state (reference :25,48-50) the DP
keeps two back-pointer arrays (split point y and the branch count of the last child) and the m-loop is
vectorised over (y, l, b), so a 768-node / depth-30 / 32-branch search is seconds, not hours.
"""
from __future__ import annotations

import json
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

NEG = np.float32(-np.inf)


class SearchTable:
    """Result of the DP: value table + back-pointers."""

    def __init__(self, F: np.ndarray, split: np.ndarray, last_branch: np.ndarray):
        self.F = F                        # (m+1, l+1, b+1) float32
        self.split = split                # y chosen for state (m,l,b), b >= 2
        self.last_branch = last_branch    # branch count of the LAST child's subtree for state (m,l,b), b >= 1

    @property
    def best(self) -> np.ndarray:
        """G[m, l] (the reference's `results`, tree_search.py:56)."""
        return self.F.max(axis=2)

    def children(self, m: int, l: int, b: int) -> List[Tuple[int, int, int]]:
        """Sub-states of the b children of the root of state (m,l,b), in draft-rank order."""
        out: List[Tuple[int, int, int]] = []
        while b >= 1:
            if b == 1:
                out.append((m - 1, l - 1, int(self.last_branch[m, l, 1])))
                break
            y = int(self.split[m, l, b])
            out.append((m - y, l - 1, int(self.last_branch[m, l, b])))
            m, b = y, b - 1
        out.reverse()
        return out


def search(p: Sequence[float], max_depth: int, max_budget: int) -> SearchTable:
    """p[j] = probability that the j-th ranked draft child is the accepted one (p[0] unused)."""
    p = np.asarray(p, dtype=np.float32)
    B = p.shape[0] - 1
    M, L = max_budget, max_depth
    F = np.full((M + 1, L + 1, B + 1), NEG, dtype=np.float32)
    G = np.full((M + 1, L + 1), NEG, dtype=np.float32)          # running max over b
    Garg = np.zeros((M + 1, L + 1), dtype=np.int64)
    split = np.zeros((M + 1, L + 1, B + 1), dtype=np.int64)
    last_branch = np.zeros((M + 1, L + 1, B + 1), dtype=np.int64)
    F[1, 1:, 0] = 1.0
    G[1, 1:] = 1.0
    G[1, 0] = NEG

    one = np.float32(1.0)
    for m in range(2, M + 1):
        # b == 1: the only child takes all remaining m-1 nodes
        if B >= 1:
            with np.errstate(invalid="ignore"):
                F[m, 2:, 1] = one + p[1] * G[m - 1, 1:L]
            last_branch[m, 2:, 1] = Garg[m - 1, 1:L]
        if B >= 2 and m >= 3:
            # cand[y-1, l-2, b-2] = F[y, l, b-1] + p[b] * G[m-y, l-1]   for y in 1..m-1, l in 2..L, b in 2..B
            left = F[1:m, 2:, 1:B]                               # (m-1, L-1, B-1)
            right = G[m - 1:0:-1, 1:L]                           # (m-1, L-1): G[m-y, l-1]
            with np.errstate(invalid="ignore"):
                prod = right[:, :, None] * p[None, None, 2:]    # one fp32 rounding
                cand = left + prod                               # second fp32 rounding
            cand = np.where(np.isnan(cand), NEG, cand)           # (-inf) + (+/-0 * -inf) never beats a finite value
            ybest = cand.argmax(axis=0)                          # first maximiser == reference's strict '>' scan
            val = np.take_along_axis(cand, ybest[None], axis=0)[0]
            F[m, 2:, 2:] = val
            split[m, 2:, 2:] = ybest + 1
            ll = np.arange(2, L + 1)[:, None]
            last_branch[m, 2:, 2:] = Garg[m - (ybest + 1), ll - 1]
        G[m] = F[m].max(axis=1)
        Garg[m] = F[m].argmax(axis=1)
    return SearchTable(F, split, last_branch)


def choose_budget_depth(table: SearchTable, draft_time: float, target_time: Sequence[float],
                        valid_budget: Sequence[int]) -> Tuple[float, Tuple[int, int]]:
    """Minimise (depth * draft_time + target_time[budget]) / expected_accept  (reference :62-74)."""
    res = table.best
    best_t, pair = np.float32(np.inf), None
    for i, b in enumerate(valid_budget):
        for d in range(res.shape[1]):
            ac = res[b, d]
            if ac < 0:
                continue
            x = np.float32(d * draft_time + target_time[i]) / ac
            if x < best_t:
                best_t, pair = x, (b, d)
    return float(best_t), pair


def build_grow_map(table: SearchTable, m: int, l: int) -> Dict:
    """Expand state (m, l, argmax_b) level by level into the growmap dict SpecTree consumes
    (roots / branches / Successors / mask / depth / size; reference :76-131)."""
    b = int(table.F[m, l].argmax())
    states = [(m, l, b)]
    parents = [-1]
    depth = [0]
    successors: List[List[int]] = [[]]
    mask = torch.zeros(m, m, dtype=torch.long)
    roots: List[List[int]] = []
    branches: List[List[int]] = []
    frontier = [0]
    while frontier:
        level_roots, level_branches, nxt = [], [], []
        for i in frontier:
            if parents[i] >= 0:
                mask[i] = mask[parents[i]]
            mask[i, i] = 1
            sm, sl, sb = states[i]
            level_roots.append(i)
            level_branches.append(sb)
            kids = table.children(sm, sl, sb) if sb else []
            assert len(kids) == sb
            first = len(states)
            ids = list(range(first, first + sb))
            successors[i].extend(ids)
            successors.extend([] for _ in ids)
            parents.extend(i for _ in ids)
            depth.extend(depth[i] + 1 for _ in ids)
            states.extend(kids)
            nxt.extend(ids)
        roots.append(level_roots)
        branches.append(level_branches)
        frontier = nxt
    n = len(states)
    assert n == m, (n, m)
    return {"roots": roots, "branches": branches, "Successors": successors, "mask": mask,
            "depth": torch.LongTensor(depth), "size": n}


def run_config(config: Dict) -> Tuple[Dict, Dict]:
    """Whole pipeline for one reference-style config dict (demo-config.json keys)."""
    p = torch.load(config["acceptance_rate_vector"], map_location="cpu").float().cpu()[:-1].numpy()
    table = search(p, config["max_depth"], config["max_budget"])
    dec_time, pair = choose_budget_depth(table, config["draft_time"], config["target_time"], config["valid_budget"])
    grow_map = build_grow_map(table, *pair)
    info = {"dec_time": dec_time, "speedup": config["target_time"][0] / dec_time, "budget": pair[0], "depth": pair[1],
            "expected_accept": float(table.best[pair[0], pair[1]])}
    return grow_map, info


def main(argv=None):
    import argparse
    ap = argparse.ArgumentParser(description="growmap search (same config keys as the reference's tree_search.py)")
    ap.add_argument("--config", type=str, default="demo-config.json")
    args = ap.parse_args(argv)
    with open(args.config) as f:
        config = json.load(f)
    grow_map, info = run_config(config)
    print(json.dumps(info))
    torch.save(grow_map, config["dst"])
    return grow_map, info
