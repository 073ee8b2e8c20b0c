"""End-to-end parity on the GPU: the drop-in SpecTree / GreedyTree + engines (CUDA graphs, device-side walk) against
(a) the decode traces recorded from the UNMODIFIED reference (tests/golden/decode_golden.pt) and (b) the CPU oracle
run side by side with shared random numbers.  Same seed -> identical accepted token sequence."""
import os

import pytest
import torch

import cases
from oracle import sequoia_oracle as O

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DEC = torch.load(os.path.join(G, "decode_golden.pt"))
DEV = "cuda:0"
F16 = torch.float16


def _engines(dkey, tkey, M):
    from Engine.Engine import GraphInferenceEngine, GraphInferenceEngineTG   # the reference's import paths
    dcfg, dw = cases.model_weights(dkey)
    tcfg, tw = cases.model_weights(tkey)
    return (GraphInferenceEngine(M, {"config": dcfg, "state_dict": dw}, device=DEV),
            GraphInferenceEngineTG(M, {"config": tcfg, "state_dict": tw}, device=DEV))


def _buffers(M):
    dtype = F16
    return dict(attn_mask=torch.full((M, M), torch.finfo(dtype).min, dtype=dtype, device=DEV),
                sequence=torch.arange(M, device=DEV).long().unsqueeze(-1),
                new_tokens_buffer=torch.zeros(M, device=DEV).long(), parents_buffer=torch.zeros(M, device=DEV).long(),
                position_ids=torch.zeros(M, device=DEV).long())


def _make_tree(mode, draft, target, prompt, gm, M):
    from Tree.GreedyTree import GreedyTree
    from Tree.SpecTree import SpecTree
    cls = SpecTree if mode == "spec" else GreedyTree
    return cls(prefix=prompt, device=DEV, temperature=0.6, top_p=1.0, draft_kv_len=0, target_kv_len=0,
               draft_model_engine=draft, target_model_engine=target, max_length=M, max_target_seq=M, grow_map=gm,
               residual_graph=None, sampling_callables=None, sample_gather_indices=None, **_buffers(M))


@pytest.mark.parametrize("name", [k for k in cases.DECODE_CASES if k.startswith("greedy")])
def test_greedy_decode_vs_reference_golden(name):
    """GreedyTree is deterministic: tree tokens, accept lengths and returned tokens must equal the trace recorded
    from the reference for every iteration (bit-exact indices)."""
    gm_name, mode, dkey, tkey, M, pseed, plen, iters, rng_seed = cases.DECODE_CASES[name]
    gm = cases.load_growmap(gm_name)
    S = gm["size"]
    draft, target = _engines(dkey, tkey, M)
    tree = _make_tree(mode, draft, target, cases.make_prompt(pseed, plen), gm, M)
    rec = DEC[name]
    assert torch.equal((tree.attn_mask[:plen + S - 1, :plen + S - 1] == 0).cpu(), rec["mask_visible0"])
    for it, g in enumerate(rec["iters"]):
        P = tree.ground_truth_len
        assert P == g["P"]
        tree.construct_grow_map()
        assert torch.equal(tree.tokens[P:P + S - 1].cpu(), g["tree_tokens"]), f"{name} iter {it}: drafted tree differs"
        valid, a, _, terminal = tree.verify()
        assert a == g["accept_len"] and terminal == g["terminal"], f"{name} iter {it}: accept {a} vs {g['accept_len']}"
        assert torch.equal(valid.cpu(), g["valid_tokens"])
        assert torch.equal(tree.position_ids.cpu(), g["position_ids"])
        if not terminal:
            n = tree.ground_truth_len
            assert torch.equal((tree.attn_mask[:n + S - 1, :n + S - 1] == 0).cpu(), g["mask_visible_next"])
    draft.clear_kv()
    target.clear_kv()


@pytest.mark.parametrize("name", [k for k in cases.DECODE_CASES if k.startswith("spec")])
def test_spec_first_iteration_vs_reference_golden(name):
    """Stochastic tree: r / rand come from the CPU generator (same seed => same draws as the reference run), so the
    drafted tree and the accept walk of iteration 0 must match the reference trace.  (Later iterations depend on
    the reference's CPU multinomial stream for the bonus token, which no GPU run can share -- see the oracle test.)"""
    gm_name, mode, dkey, tkey, M, pseed, plen, iters, rng_seed = cases.DECODE_CASES[name]
    gm = cases.load_growmap(gm_name)
    S = gm["size"]
    draft, target = _engines(dkey, tkey, M)
    torch.manual_seed(rng_seed)
    tree = _make_tree(mode, draft, target, cases.make_prompt(pseed, plen), gm, M)
    g = DEC[name]["iters"][0]
    tree.construct_grow_map()
    assert torch.equal(tree.tokens[plen:plen + S - 1].cpu(), g["tree_tokens"]), f"{name}: drafted tree differs"
    valid, a, _, terminal = tree.verify()
    assert a == g["accept_len"] and terminal == g["terminal"]
    assert torch.equal(valid[:a].cpu(), g["valid_tokens"][:a])
    draft.clear_kv()
    target.clear_kv()


@pytest.mark.parametrize("name", list(cases.DECODE_CASES))
@pytest.mark.parametrize("graphs", [True, False])
def test_decode_vs_oracle_side_by_side(name, graphs):
    """Full multi-iteration decode against the CPU oracle with shared r / rand (CPU generator) and shared Exp(1)
    noise for the bonus token: identical tree tokens, accept lists, bonus tokens and returned sequences."""
    gm_name, mode, dkey, tkey, M, pseed, plen, iters, rng_seed = cases.DECODE_CASES[name]
    gm = cases.load_growmap(gm_name)
    S = gm["size"]
    prompt = cases.make_prompt(pseed, plen)
    dcfg, dw = cases.model_weights(dkey)
    tcfg, tw = cases.model_weights(tkey)
    od, ot = O.EngineOracle(O.LlamaOracle(dcfg, dw, M, "FI")), O.EngineOracle(O.LlamaOracle(tcfg, tw, M, "TG"))
    noise = torch.empty(iters, cases.V, dtype=F16).exponential_(1.0, generator=torch.Generator().manual_seed(5))
    torch.manual_seed(rng_seed)
    if mode == "spec":
        otree = O.SpecTreeOracle(od, ot, prompt, gm, temperature=0.6, top_p=1.0, max_length=M, bonus_noise=noise)
    else:
        otree = O.GreedyTreeOracle(od, ot, prompt, gm, max_length=M)
    draft, target = _engines(dkey, tkey, M)
    torch.manual_seed(rng_seed)
    tree = _make_tree(mode, draft, target, prompt, gm, M)
    tree.rt.use_graphs = graphs
    tree.rt.external_noise = noise.to(DEV) if mode == "spec" else None
    try:
        for it in range(iters):
            P = tree.ground_truth_len
            assert P == otree.ground_truth_len
            otree.construct_grow_map()
            tree.construct_grow_map()
            assert torch.equal(tree.tokens[P:P + S - 1].cpu(), otree.tokens[P:P + S - 1]), f"{name} iter {it}: tree"
            ov, oa, _, oterm = otree.verify()
            valid, a, _, terminal = tree.verify()
            assert tree.accept_list() == otree.last_trace.accept_list, f"{name} iter {it}: accept list"
            assert (a, terminal) == (oa, oterm)
            assert torch.equal(valid.cpu(), ov), f"{name} iter {it}: returned tokens"
            # KV rows below kv_len must hold the same accepted path (values within fp16 GEMM noise)
            kk = target.engine.kv_cache.k_cache[..., :a, :].float().cpu()
            assert torch.allclose(kk, ot.kv_cache.k_cache[..., :a, :].float(), atol=5e-3, rtol=5e-3)
            if terminal:
                break
    finally:
        tree.rt.external_noise = None
        tree.rt.use_graphs = True
        draft.clear_kv()
        target.clear_kv()


def test_benchmark_mode_tuple_arity():
    """benchmark=True keeps the reference's return arity (SpecTree.py:234-242: 7-tuple; construct: 2-tuple)."""
    gm_name, mode, dkey, tkey, M, pseed, plen, iters, rng_seed = cases.DECODE_CASES["spec_8x8"]
    gm = cases.load_growmap(gm_name)
    draft, target = _engines(dkey, tkey, M)
    torch.manual_seed(rng_seed)
    tree = _make_tree(mode, draft, target, cases.make_prompt(pseed, plen), gm, M)
    out = tree.construct_grow_map(benchmark=True)
    assert isinstance(out, tuple) and len(out) == 2
    res = tree.verify(benchmark=True)
    assert len(res) == 7 and isinstance(res[-1], bool)
    tree.construct_grow_map()
    assert len(tree.verify()) == 4
