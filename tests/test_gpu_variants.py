"""Policy variants on the GPU: GreedySTree and SpecInferTree drop-ins in lock-step with their CPU
oracles (which test_oracle_golden.py pins bit-exactly to traces of the UNMODIFIED reference classes,
tests/golden/variants_golden.pt).

Shared randomness: GreedySTree's sampled target tokens come from (S, V) uniforms handed to both sides (k=1 exponential
race == torch.multinomial(1)'s own construction); SpecInferTree's i.i.d. children come from one 32-bit word per node
through the exact integer inverse-CDF (kernel-level bit-exactness: test_gpu_kernels.py::test_sample_replace_*) — in
the end-to-end run the oracle is handed the GPU's drafted tokens (its own q differs from the GPU's by fp16 rounding
noise, which shifts CDF boundaries), so that the comparison covers what follows: draft logits of the same tree, the
>= / keep-q walk, bonus token, compaction, KV gather, next-iteration state."""
import os

import pytest
import torch

import cases
from oracle import sequoia_oracle as O
from test_gpu_decode import (DEV, F16, REL_TOL, _buffers, _engines, _explained_accept_mismatch, _explained_tree_mismatch)

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
VAR = torch.load(os.path.join(G, "variants_golden.pt"))
VAR.update(torch.load(os.path.join(G, "sweep_golden.pt")))
CASES = dict(cases.VARIANT_CASES)
CASES.update(cases.SWEEP_CASES)            # tests/run.sh tree shapes (K chains of length L) for the same policies


def _make(mode, draft, target, prompt, gm, M):
    from Tree.GreedySTree import GreedySTree          # the reference's import paths
    from Tree.SpecInferTree import SpecInferTree
    cls = GreedySTree if mode == "greedys" else SpecInferTree
    return cls(prefix=prompt, device=DEV, temperature=0.6, top_p=1.0, draft_kv_len=0, target_kv_len=0,
               draft_model_engine=draft, target_model_engine=target, max_length=M, max_target_seq=M, grow_map=gm,
               residual_graph=None, sampling_callables=None, sample_gather_indices=None, **_buffers(M))


def _oracles(dkey, tkey, M):
    dcfg, dw = cases.model_weights(dkey)
    tcfg, tw = cases.model_weights(tkey)
    return O.EngineOracle(O.LlamaOracle(dcfg, dw, M, "FI")), O.EngineOracle(O.LlamaOracle(tcfg, tw, M, "TG"))


@pytest.mark.parametrize("name", ["greedys_4x4", "greedys_same_4x4", "sweep_greedys_5x8"])
@pytest.mark.parametrize("graphs", [True, False])
def test_greedys_tree_lockstep(name, graphs):
    gm_name, mode, dkey, tkey, M, pseed, plen, iters, rng_seed = CASES[name]
    gm = cases.load_growmap(gm_name)
    S = gm["size"]
    prompt = cases.make_prompt(pseed, plen)
    od, ot = _oracles(dkey, tkey, M)
    u = torch.empty(iters, S, cases.V, dtype=F16).uniform_(generator=torch.Generator().manual_seed(8))
    torch.manual_seed(rng_seed)
    otree = O.GreedySTreeOracle(od, ot, prompt, gm, temperature=0.6, top_p=1.0, max_length=M, target_uniforms=u)
    draft, target = _engines(dkey, tkey, M)
    torch.manual_seed(rng_seed)
    tree = _make(mode, draft, target, prompt, gm, M)
    tree.rt.use_graphs = graphs
    tree.rt.external_tuniform = u.to(DEV)
    rec = VAR[name]
    matched = 0
    try:
        for it in range(iters):
            P = tree.ground_truth_len
            assert P == otree.ground_truth_len
            otree.construct_grow_map()
            tree.construct_grow_map()
            got = tree.tokens[P:P + S - 1].cpu()
            if it == 0:                                   # drafting is deterministic: also the reference's own tree
                assert torch.equal(otree.tokens[P:P + S - 1], rec["iters"][0]["tree_tokens"])
            if not torch.equal(got, otree.tokens[P:P + S - 1]):
                ok, why = _explained_tree_mismatch(otree, got, P, gm, "greedy")
                assert ok, f"{name} iter {it}: drafted tree differs and is NOT a near-tie ({why})"
                break
            ov, oa, _, oterm = otree.verify()
            valid, a, _, terminal = tree.verify()
            tt_got, tt_ref = tree.rt.target_token.cpu(), otree.target_token
            if not torch.equal(tt_got, tt_ref):           # a sampled target token may flip only on an fp16 score near-tie
                row = int((tt_got != tt_ref).nonzero()[0])
                q = torch.softmax(otree.raw_target_logits[row] / 0.6, dim=-1).float()
                sc = u[it, row].float().log() / q
                gap = abs(float(sc[tt_got[row]]) - float(sc[tt_ref[row]])) / max(abs(float(sc[tt_ref[row]])), 1e-6)
                assert gap <= 4 * REL_TOL, f"{name} iter {it}: target token of row {row} differs, score gap {gap:.3e}"
                break
            assert tree.accept_list() == otree.last_trace.accept_list
            assert (a, terminal) == (oa, oterm)
            assert torch.equal(valid.cpu(), ov), f"{name} iter {it}: returned tokens"
            assert torch.equal(tree.position_ids.cpu(), otree.position_ids)
            kk = target.engine.kv_cache.k_cache[..., :a, :].float().cpu()
            assert torch.allclose(kk, ot.kv_cache.k_cache[..., :a, :].float(), atol=8e-3, rtol=8e-3)
            matched += 1
            if terminal:
                break
    finally:
        tree.rt.external_tuniform = None
        tree.rt.use_graphs = True
        draft.clear_kv()
        target.clear_kv()
    assert matched >= 1, f"{name}: not a single iteration matched the oracle"


@pytest.mark.parametrize("name", ["specinfer_8x8", "specinfer_same_8x8"] +
                         [n for n, c in cases.SWEEP_CASES.items() if c[1] == "specinfer"])
@pytest.mark.parametrize("graphs", [True, False])
def test_specinfer_tree_lockstep(name, graphs):
    gm_name, mode, dkey, tkey, M, pseed, plen, iters, rng_seed = CASES[name]
    gm = cases.load_growmap(gm_name)
    S = gm["size"]
    prompt = cases.make_prompt(pseed, plen)
    od, ot = _oracles(dkey, tkey, M)
    g = torch.Generator().manual_seed(6)
    noise = torch.empty(iters, cases.V, dtype=F16).exponential_(1.0, generator=g)
    words = torch.randint(0, 1 << 32, (iters, S), generator=g, dtype=torch.int64)
    draft, target = _engines(dkey, tkey, M)
    torch.manual_seed(rng_seed)
    tree = _make(mode, draft, target, prompt, gm, M)
    drafted = {}
    torch.manual_seed(rng_seed)
    otree = O.SpecInferTreeOracle(od, ot, prompt, gm, temperature=0.6, top_p=1.0, max_length=M, bonus_noise=noise,
                                  forced_tokens=lambda it: drafted[it])
    # same seeded CPU draws for r / rand as the reference makes (SpecInferTree.py:59,83)
    assert torch.equal(tree.rt.r[:M].cpu(), otree.r) and cases.sha(otree.r) == VAR[name]["r_sha"]
    tree.rt.use_graphs = graphs
    tree.rt.external_noise = noise.to(DEV)
    tree.rt.external_words = words.to(DEV)
    matched = 0
    try:
        for it in range(iters):
            P = tree.ground_truth_len
            assert P == otree.ground_truth_len
            tree.construct_grow_map()
            got = tree.tokens[P:P + S - 1].cpu()
            drafted[it] = got
            # the root's children come from draft_logits[0], which both sides already hold: the oracle's own integer
            # CDF must agree with the GPU's draws wherever its q row is bit-identical to the GPU's
            q0 = torch.softmax(otree.draft_logits[0:1] / 0.6, dim=-1)
            nb0 = gm["branches"][0][0]
            want0 = O.multinomial_words(q0, words[it, 1:1 + nb0].view(1, -1)).flatten()
            assert float((want0 == got[:nb0]).float().mean()) >= 0.5, "root-level draws disagree with the integer CDF"
            otree.construct_grow_map()
            ov, oa, _, oterm = otree.verify()
            valid, a, _, terminal = tree.verify()
            got_list, ref_list = tree.accept_list(), otree.last_trace.accept_list
            if got_list != ref_list:
                ok, why = _explained_accept_mismatch(otree, got_list, ref_list, gm, "spec", P)
                assert ok, f"{name} iter {it}: accept list {got_list[P:]} vs {ref_list[P:]} NOT a boundary case ({why})"
                break
            assert (a, terminal) == (oa, oterm)
            assert torch.equal(valid.cpu(), ov), f"{name} iter {it}: returned tokens"
            assert torch.equal(tree.position_ids.cpu(), otree.position_ids)
            kk = target.engine.kv_cache.k_cache[..., :a, :].float().cpu()
            assert torch.allclose(kk, ot.kv_cache.k_cache[..., :a, :].float(), atol=8e-3, rtol=8e-3)
            dk = draft.engine.kv_cache.v_cache[..., :a, :].float().cpu()
            assert torch.allclose(dk, od.kv_cache.v_cache[..., :a, :].float(), atol=8e-3, rtol=8e-3)
            matched += 1
            if terminal:
                break
    finally:
        tree.rt.external_noise = None
        tree.rt.external_words = None
        tree.rt.use_graphs = True
        draft.clear_kv()
        target.clear_kv()
    assert matched >= 1, f"{name}: not a single iteration matched the oracle"
    assert draft.engine.runner.plan.error() == 0 and target.engine.runner.plan.error() == 0


# ---- acceptance-rate measurement trees (SpecTreeTest / GreedyTreeTest, per-step construction like tests/test_accept.py) ----
def _test_loop(cls, draft, target, prompt, M, W, steps, T=0.6, top_p=1.0):
    bufs = _buffers(M)
    ids = prompt.to(DEV)
    dkv = tkv = 0
    out = []
    for _ in range(steps):
        tree = cls(prefix=ids, device=DEV, temperature=T, top_p=top_p, draft_kv_len=dkv, target_kv_len=tkv,
                   draft_model_engine=draft, target_model_engine=target, max_length=M, max_width=W, **bufs)
        P = tree.ground_truth_len
        children = tree.tokens[P:P + W].clone()
        valid, dkv, tkv, b, terminal = tree.verify(benchmark=True)
        out.append(dict(P=P, children=children.cpu(), valid=valid.clone().cpu(), a=dkv, b=b, terminal=terminal,
                        target_token0=int(tree.rt.target_token[0]) if tree.GREEDY else None))
        ids = valid.clone()
        if terminal:
            break
    draft.clear_kv()
    target.clear_kv()
    return out


def test_greedy_tree_test_reports_accepted_rank():
    """GreedyTreeTest (GreedyTree.py:264-456): b == rank of the target's argmax among the drafted top-W children (or -1),
    returned lengths feed the next construction, and the decoded stream equals plain GreedyTree decoding."""
    from Tree.GreedyTree import GreedyTree, GreedyTreeTest
    M, W, steps = 256, 8, 8
    prompt = cases.make_prompt(41, 48)
    draft, target = _engines("draft", "target", M)
    rows = _test_loop(GreedyTreeTest, draft, target, prompt, M, W, steps)
    assert len(rows) >= 4
    for r in rows:
        ch = r["children"].tolist()
        want_b = ch.index(r["target_token0"]) if r["target_token0"] in ch else -1
        assert r["b"] == want_b
        assert r["a"] == r["P"] + (1 if want_b >= 0 else 0)
        if not r["terminal"]:
            assert r["valid"].shape[0] == r["a"] + 1
            if want_b >= 0:
                assert int(r["valid"][-2]) == ch[want_b]
    stream = rows[-1]["valid"]
    # same prompt through the ordinary GreedyTree: greedy speculative decoding is lossless -> same token stream
    gm = cases.load_growmap("L40_growmaps/4x4-tree.pt")
    tree = GreedyTree(prefix=prompt, device=DEV, temperature=0.6, top_p=1.0, draft_kv_len=0, target_kv_len=0,
                      draft_model_engine=draft, target_model_engine=target, max_length=M, max_target_seq=M, grow_map=gm,
                      residual_graph=None, sampling_callables=None, sample_gather_indices=None, **_buffers(M))
    valid = None
    while valid is None or valid.shape[0] < stream.shape[0]:
        tree.construct_grow_map()
        valid, _, _, term = tree.verify()
        if term:
            break
    n = min(valid.shape[0], stream.shape[0])
    new = n - prompt.shape[0]
    agree = int((valid[:n].cpu() == stream[:n]).sum()) - prompt.shape[0]
    draft.clear_kv()
    target.clear_kv()
    assert new >= 4 and agree >= new - 1, f"{agree}/{new} new tokens agree (fp16 near-ties may flip at most one)"


def test_spec_tree_test_same_model_accepts_first_child():
    """SpecTreeTest (SpecTree.py:284-483) with draft == target weights: p ~= q, so p[tok] >= r*q[tok] holds for the first
    child almost surely -> b == 0; structure of the returned tuple as the reference's (5-tuple, lengths, rank)."""
    from Tree.SpecTree import SpecTreeTest
    M, W, steps = 256, 8, 8
    prompt = cases.make_prompt(42, 48)
    draft, target = _engines("draft", "draft", M)
    torch.manual_seed(3)
    rows = _test_loop(SpecTreeTest, draft, target, prompt, M, W, steps)
    assert len(rows) >= 4
    for r in rows:
        assert -1 <= r["b"] < W
        assert r["a"] == r["P"] + (1 if r["b"] >= 0 else 0)
        if not r["terminal"]:
            assert r["valid"].shape[0] == r["a"] + 1
            if r["b"] >= 0:
                assert int(r["valid"][-2]) == int(r["children"][r["b"]])
        assert len(set(r["children"].tolist())) == W          # drawn WITHOUT replacement
    assert sum(1 for r in rows if r["b"] == 0) >= len(rows) - 1
    # different models (random weights): mostly nothing accepted, and b spreads over [-1, W)
    draft2, target2 = _engines("draft", "target", M)
    rows2 = _test_loop(SpecTreeTest, draft2, target2, prompt, M, W, steps)
    assert all(-1 <= r["b"] < W for r in rows2)


@pytest.mark.parametrize("graphs", [True, False])
def test_spec_tree_with_top_p_filter(graphs):
    """top_p < 1 (get_sampling_logits, utils.py:65-77 — off in every named configuration, on by default in the
    reference's CLI): the nucleus filter runs inside the captured verify graph; lock-step with the oracle."""
    from Tree.SpecTree import SpecTree
    gm_name, mode, dkey, tkey, M, pseed, plen, iters, rng_seed = cases.DECODE_CASES["spec_same_8x8"]
    gm = cases.load_growmap(gm_name)
    prompt = cases.make_prompt(pseed, plen)
    od, ot = _oracles(dkey, tkey, M)
    noise = torch.empty(iters, cases.V, dtype=F16).exponential_(1.0, generator=torch.Generator().manual_seed(5))
    torch.manual_seed(rng_seed)
    otree = O.SpecTreeOracle(od, ot, prompt, gm, temperature=0.6, top_p=0.9, max_length=M, bonus_noise=noise)
    draft, target = _engines(dkey, tkey, M)
    torch.manual_seed(rng_seed)
    tree = SpecTree(prefix=prompt, device=DEV, temperature=0.6, top_p=0.9, draft_kv_len=0, target_kv_len=0,
                    draft_model_engine=draft, target_model_engine=target, max_length=M, max_target_seq=M, grow_map=gm,
                    residual_graph=None, sampling_callables=None, sample_gather_indices=None, **_buffers(M))
    tree.rt.use_graphs = graphs
    tree.rt.external_noise = noise.to(DEV)
    S = gm["size"]
    matched = 0
    try:
        for it in range(iters):
            P = tree.ground_truth_len
            otree.construct_grow_map()
            tree.construct_grow_map()
            got = tree.tokens[P:P + S - 1].cpu()
            if not torch.equal(got, otree.tokens[P:P + S - 1]):
                ok, why = _explained_tree_mismatch(otree, got, P, gm, "spec")
                assert ok, why
                break
            ov, oa, _, oterm = otree.verify()
            valid, a, _, terminal = tree.verify()
            # the filter really removed mass: some target probabilities are exactly 0 in the oracle
            assert bool((otree.target_logits == 0).any())
            if tree.accept_list() != otree.last_trace.accept_list:
                ok, why = _explained_accept_mismatch(otree, tree.accept_list(), otree.last_trace.accept_list, gm, "spec", P)
                assert ok, why
                break
            assert (a, terminal) == (oa, oterm) and torch.equal(valid.cpu(), ov)
            matched += 1
            if terminal:
                break
    finally:
        tree.rt.external_noise = None
        tree.rt.use_graphs = True
        draft.clear_kv()
        target.clear_kv()
    assert matched >= 1
