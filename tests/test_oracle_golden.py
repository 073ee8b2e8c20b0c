"""Pin the CPU oracle (oracle/sequoia_oracle.py) against the golden vectors produced by the
UNMODIFIED reference (tests/golden/make_golden.py).  Bit-exact: same torch build, same ops."""
import os

import pytest
import torch

import cases
from oracle import sequoia_oracle as O

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
UT = torch.load(os.path.join(G, "utils_golden.pt"))
DEC = torch.load(os.path.join(G, "decode_golden.pt"))
GM = torch.load(os.path.join(G, "growmaps_golden.pt"))


@pytest.mark.parametrize("name", [k for k in UT if k.startswith("swor")])
def test_sampling_without_replacement(name):
    g = UT[name]
    logits, rand = cases.sampling_case(g["seed"], g["rows"], g["peaked"])
    pos = O.sampling_without_replacement(logits, rand, g["k"], g["T"])
    assert torch.equal(pos, g["positions"])
    q = torch.softmax(logits / g["T"], dim=-1)
    assert cases.sha(q) == g["q_sha"]


@pytest.mark.parametrize("name", [k for k in UT if k.startswith("argmax")])
def test_sampling_argmax(name):
    g = UT[name]
    logits, _ = cases.sampling_case(g["seed"], g["rows"], g["peaked"])
    assert torch.equal(O.sampling_argmax(logits, g["k"]), g["positions"])


def test_residual_and_masks():
    for seed in (7, 8):
        p, q = cases.residual_case(seed)
        assert torch.equal(O.get_residual(p.clone(), q.clone()), UT[f"residual_{seed}"]["residual"])
    p, _ = cases.residual_case(9)
    assert torch.isnan(O.get_residual(p.clone(), p.clone())).all() and UT["residual_nan"]["residual_isnan_all"]
    assert torch.equal(O.make_causal_mask(8), UT["causal_8"])
    g = torch.Generator().manual_seed(21)
    lg = (torch.randn(4, 1000, generator=g) * 3).to(torch.float16)
    assert torch.equal(O.get_sampling_logits(lg.clone(), 0.9, 0.6), UT["top_p_0.9"]["out"])


def test_growmap_files_match_reference():
    """Every shipped growmap is byte-identical in structure to the reference's file."""
    assert len(GM) >= 100
    for rel, g in GM.items():
        m = cases.load_growmap(rel)
        assert m["size"] == g["size"]
        assert cases.sha(m["mask"]) == g["mask_sha"] and cases.sha(m["depth"]) == g["depth_sha"]
        assert [len(r) for r in m["roots"]] == g["levels"]
        assert [sum(b) for b in m["branches"]] == g["n_children"]


def _oracle_engines(dkey, tkey, M):
    dcfg, dw = cases.model_weights(dkey)
    tcfg, tw = cases.model_weights(tkey)
    return (O.EngineOracle(O.LlamaOracle(dcfg, dw, M, "FI")), O.EngineOracle(O.LlamaOracle(tcfg, tw, M, "TG")))


VAR = torch.load(os.path.join(G, "variants_golden.pt"))
VAR.update(torch.load(os.path.join(G, "sweep_golden.pt")))
ALL_CASES = dict(cases.DECODE_CASES, **cases.VARIANT_CASES, **cases.SWEEP_CASES)


@pytest.mark.parametrize("name", list(ALL_CASES))
def test_decode_trace(name):
    """Oracle == unmodified reference, step by step (SpecTree / GreedyTree and the GreedySTree / SpecInferTree policy
    variants), sharing one CPU RNG seed so that every torch draw lines up."""
    gm_name, mode, dkey, tkey, M, pseed, plen, iters, rng_seed = ALL_CASES[name]
    rec = DEC[name] if name in DEC else VAR[name]
    gm = cases.load_growmap(gm_name)
    draft, target = _oracle_engines(dkey, tkey, M)
    prompt = cases.make_prompt(pseed, plen)
    torch.manual_seed(rng_seed)
    if mode in ("spec", "specinfer"):
        cls = O.SpecTreeOracle if mode == "spec" else O.SpecInferTreeOracle
        tree = cls(draft, target, prompt, gm, temperature=0.6, top_p=1.0, max_length=M)
        assert cases.sha(tree.r) == rec["r_sha"] and cases.sha(tree.rand) == rec["rand_sha"]
    elif mode == "greedys":
        tree = O.GreedySTreeOracle(draft, target, prompt, gm, temperature=0.6, top_p=1.0, max_length=M)
    else:
        tree = O.GreedyTreeOracle(draft, target, prompt, gm, max_length=M)
    S = gm["size"]
    assert cases.sha(tree.draft_logits[0]) == rec["draft_logits0_sha"]
    tot = plen + S - 1
    assert torch.equal(tree.attn_mask[:tot, :tot] == 0, rec["mask_visible0"])
    assert torch.equal(O.visible_from_rule(M, plen, gm["mask"]), rec["mask_visible0"])
    assert torch.equal(tree.position_ids, rec["position_ids0"])
    for it, g in enumerate(rec["iters"]):
        P = tree.ground_truth_len
        assert P == g["P"]
        tree.construct_grow_map()
        assert torch.equal(tree.tokens[P:P + S - 1], g["tree_tokens"]), f"iter {it} tree tokens"
        assert cases.sha(tree.draft_logits[:S]) == g["draft_logits_sha"]
        valid, a, _, terminal = tree.verify()
        assert a == g["accept_len"] and terminal == g["terminal"]
        assert torch.equal(valid, g["valid_tokens"])
        if mode in ("spec", "specinfer", "greedys"):
            assert cases.sha(tree.target_logits) == g["target_logits_sha"]
        assert cases.sha(draft.kv_cache.k_cache) == g["draft_k_sha"]
        assert cases.sha(draft.kv_cache.v_cache) == g["draft_v_sha"]
        assert cases.sha(target.kv_cache.k_cache) == g["target_k_sha"]
        assert cases.sha(target.kv_cache.v_cache) == g["target_v_sha"]
        assert draft.kv_cache.kv_offset == g["draft_kv_offset"] and target.kv_cache.kv_offset == g["target_kv_offset"]
        assert torch.equal(tree.position_ids, g["position_ids"])
        if not terminal:
            n = tree.ground_truth_len
            tot = n + S - 1
            assert torch.equal(tree.attn_mask[:tot, :tot] == 0, g["mask_visible_next"])
            assert torch.equal(O.visible_from_rule(M, n, gm["mask"]), g["mask_visible_next"])


def test_multinomial_words_is_an_exact_inverse_cdf_sampler():
    """The integer inverse-CDF used for SpecInfer-style draws (oracle + sq_sample_replace): every draw lands on a token
    with q > 0, the extreme words hit the first / last support token, and the empirical law follows q."""
    g = torch.Generator().manual_seed(3)
    q = torch.softmax((torch.randn(2, 512, generator=g) * 3).half() / 0.6, dim=-1)
    q[1, :100] = 0                                                    # leading zero-probability tokens
    words = torch.tensor([[0, (1 << 32) - 1], [0, (1 << 32) - 1]])
    ends = O.multinomial_words(q, words)
    for r in range(2):
        nz = (q[r] > 0).nonzero().flatten()
        assert int(ends[r, 0]) == int(nz[0]) and int(ends[r, 1]) == int(nz[-1])
    n = 200_000
    w = torch.randint(0, 1 << 32, (2, n), generator=g, dtype=torch.int64)
    idx = O.multinomial_words(q, w)
    for r in range(2):
        assert bool((q[r][idx[r]] > 0).all())
        emp = torch.bincount(idx[r], minlength=512).double() / n
        ref = q[r].double() / q[r].double().sum()
        assert float((emp - ref).abs().sum()) < 0.06                  # total variation*2 of a 200k-sample histogram
    # same answer as a float64 inverse CDF away from the boundaries
    cdf = (q[0].double() / q[0].double().sum()).cumsum(0)
    u = (w[0, :2000].double() + 0.5) / 4294967296.0
    f64 = torch.searchsorted(cdf, u, right=True).clamp(max=511)
    assert float((f64 == idx[0, :2000]).double().mean()) > 0.995
