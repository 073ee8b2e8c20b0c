"""Generate tests/golden/*.pt by running the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container only:   python tests/golden/make_golden.py
(the reference cannot travel to the GPU box; the vectors it produced are committed).
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref = ref_shim.load_reference()          # puts /root/reference first on sys.path (Engine, Tree, utils = reference)

# tests/cases.py imports `oracle.*` from the repo root; import it by path WITHOUT letting the repo
# root shadow the reference's top-level packages.
import importlib.util  # noqa: E402

ROOT = os.path.dirname(TESTS)
sys.path.append(ROOT)                    # appended => /root/reference still wins for Engine/Tree/utils
spec = importlib.util.spec_from_file_location("cases", os.path.join(TESTS, "cases.py"))
cases = importlib.util.module_from_spec(spec)
spec.loader.exec_module(cases)
assert ref.U.__file__.startswith("/root/reference"), ref.U.__file__

U = ref.U
T = 0.6


def gen_utils():
    out = {}
    for name, (seed, rows, k, peaked) in {
        "swor_flat_1x19": (1, 1, 19, False), "swor_flat_19x13": (2, 19, 13, False),
        "swor_peaked_34x6": (3, 34, 6, True), "swor_peaked_27x4": (4, 27, 4, True),
        "swor_peaked_8x8": (5, 8, 8, True),
    }.items():
        logits, rand = cases.sampling_case(seed, rows, peaked)
        pos = U.sampling_without_replacement(logits, rand, k, T)
        q = torch.softmax(logits / T, dim=-1)
        out[name] = {"seed": seed, "rows": rows, "k": k, "peaked": peaked, "T": T, "positions": pos,
                     "q_sha": cases.sha(q), "q_head": q[:, :64].clone()}
        pos2 = U.sampling_argmax(logits, k)
        out[name.replace("swor", "argmax")] = {"seed": seed, "rows": rows, "k": k, "peaked": peaked, "positions": pos2}
    for seed in (7, 8):
        p, q = cases.residual_case(seed)
        res = U.get_residual(p.clone(), q.clone())
        out[f"residual_{seed}"] = {"seed": seed, "residual": res}
    # p == q  ->  0/0 = NaN everywhere (the reference's NaN => terminal condition, SpecTree.py:219)
    p, _ = cases.residual_case(9)
    out["residual_nan"] = {"seed": 9, "residual_isnan_all": bool(torch.isnan(U.get_residual(p.clone(), p.clone())).all())}
    out["causal_8"] = U._make_causal_mask((1, 8), torch.float16, "cpu")
    g = torch.Generator().manual_seed(21)
    lg = (torch.randn(4, 1000, generator=g) * 3).to(torch.float16)
    out["top_p_0.9"] = {"seed": 21, "out": U.get_sampling_logits(lg.clone(), 0.9, T)}
    torch.save(out, os.path.join(HERE, "utils_golden.pt"))
    print("utils_golden.pt", list(out))


def build_engines(dkey, tkey, M):
    dcfg, dw = cases.model_weights(dkey)
    tcfg, tw = cases.model_weights(tkey)
    draft = ref_shim.make_engine(ref, dcfg, dw, M, "FI")
    target = ref_shim.make_engine(ref, tcfg, tw, M, "TG")
    return draft, target


def run_decode(name, table=None):
    table = cases.DECODE_CASES if table is None else table
    gm_name, mode, dkey, tkey, M, pseed, plen, iters, rng_seed = table[name]
    gm = cases.load_growmap(gm_name)
    draft, target = build_engines(dkey, tkey, M)
    prompt = cases.make_prompt(pseed, plen)
    torch.manual_seed(rng_seed)
    dtype = torch.float16
    attn_mask = torch.full((M, M), torch.finfo(dtype).min, dtype=dtype)
    sequence = torch.arange(M).long().unsqueeze(-1)
    new_tokens_buffer = torch.zeros(M).long()
    parents_buffer = torch.zeros(M).long()
    position_ids = torch.zeros(M).long()
    branches = gm["branches"]
    steps = len(gm["roots"])
    gather = {i: torch.cat([torch.arange(b) + j * max(branches[i]) for j, b in enumerate(branches[i])]).long()
              for i in range(steps - 1)}
    if mode == "spec":
        samp = {i: (lambda k: (lambda lg, rd: U.sampling_without_replacement(lg, rd, k, T)))(max(branches[i]))
                for i in range(steps - 1)}
        tree = ref.ST.SpecTree(prefix=prompt, device="cpu", temperature=T, top_p=1.0, draft_kv_len=0, target_kv_len=0,
                               draft_model_engine=draft, target_model_engine=target, max_length=M, max_target_seq=M,
                               grow_map=gm, attn_mask=attn_mask, sequence=sequence, new_tokens_buffer=new_tokens_buffer,
                               parents_buffer=parents_buffer, position_ids=position_ids,
                               residual_graph=lambda p, q: U.get_residual(p, q), sampling_callables=samp,
                               sample_gather_indices=gather)
    elif mode == "specinfer":
        tree = ref.SIT.SpecInferTree(prefix=prompt, device="cpu", temperature=T, top_p=1.0, draft_kv_len=0, target_kv_len=0,
                                     draft_model_engine=draft, target_model_engine=target, max_length=M, max_target_seq=M,
                                     grow_map=gm, attn_mask=attn_mask, sequence=sequence,
                                     new_tokens_buffer=new_tokens_buffer, parents_buffer=parents_buffer,
                                     position_ids=position_ids, residual_graph=lambda p, q: U.get_residual(p, q),
                                     sampling_callables=None, sample_gather_indices=gather)
    else:
        samp = {i: (lambda k: (lambda lg: U.sampling_argmax(lg, k)))(max(branches[i])) for i in range(steps - 1)}
        cls = ref.GST.GreedySTree if mode == "greedys" else ref.GT.GreedyTree
        tree = cls(prefix=prompt, device="cpu", temperature=T, top_p=1.0, draft_kv_len=0, target_kv_len=0,
                                 draft_model_engine=draft, target_model_engine=target, max_length=M, max_target_seq=M,
                                 grow_map=gm, attn_mask=attn_mask, sequence=sequence, new_tokens_buffer=new_tokens_buffer,
                                 parents_buffer=parents_buffer, position_ids=position_ids,
                                 residual_graph=None, sampling_callables=samp, sample_gather_indices=gather)
    S = gm["size"]
    rec = {"case": table[name], "iters": []}
    rec["draft_logits0_sha"] = cases.sha(tree.draft_logits[0])
    rec["draft_logits0_head"] = tree.draft_logits[0][:64].clone()
    tot = plen + S - 1
    rec["mask_visible0"] = (tree.attn_mask[:tot, :tot] == 0)             # the window SpecTree built (bool)
    rec["position_ids0"] = tree.position_ids.clone()
    if mode in ("spec", "specinfer"):
        rec["r_sha"] = cases.sha(tree.r)
        rec["rand_sha"] = cases.sha(tree.rand)
    for it in range(iters):
        P = tree.ground_truth_len
        tree.construct_grow_map()
        tree_tokens = tree.tokens[P:P + S - 1].clone()
        dl_sha = cases.sha(tree.draft_logits[:S])
        valid, a, _, terminal = tree.verify()
        item = {"P": P, "tree_tokens": tree_tokens, "draft_logits_sha": dl_sha, "accept_len": a,
                "valid_tokens": valid.clone(), "terminal": terminal,
                "target_logits_sha": cases.sha(tree.target_logits),
                "target_logits_head": tree.target_logits[:, :16].clone(),
                "draft_k_sha": cases.sha(draft.engine.kv_cache.k_cache), "draft_v_sha": cases.sha(draft.engine.kv_cache.v_cache),
                "target_k_sha": cases.sha(target.engine.kv_cache.k_cache), "target_v_sha": cases.sha(target.engine.kv_cache.v_cache),
                "draft_kv_offset": draft.engine.kv_cache.kv_offset, "target_kv_offset": target.engine.kv_cache.kv_offset,
                "position_ids": tree.position_ids.clone()}
        if not terminal:
            tot = tree.ground_truth_len + S - 1
            item["mask_visible_next"] = (tree.attn_mask[:tot, :tot] == 0)
        rec["iters"].append(item)
        print(name, "iter", it, "P", P, "accept_len", a, "new", a + (0 if terminal else 1) - P, "terminal", terminal)
        if terminal:
            break
    return rec


def gen_decode():
    out = {name: run_decode(name) for name in cases.DECODE_CASES}
    torch.save(out, os.path.join(HERE, "decode_golden.pt"))
    print("decode_golden.pt", list(out))


def gen_variants():
    out = {name: run_decode(name, cases.VARIANT_CASES) for name in cases.VARIANT_CASES}
    torch.save(out, os.path.join(HERE, "variants_golden.pt"))
    print("variants_golden.pt", list(out))


def gen_sweep():
    out = {name: run_decode(name, cases.SWEEP_CASES) for name in cases.SWEEP_CASES}
    torch.save(out, os.path.join(HERE, "sweep_golden.pt"))
    print("sweep_golden.pt", list(out))


def gen_growmaps():
    """Structure goldens for every growmap shipped (tree indices / mask bit-exact): sha of each field."""
    import glob
    out = {}
    for path in sorted(glob.glob(os.path.join(ROOT, "*_growmaps", "**", "*.pt"), recursive=True)):
        rel = os.path.relpath(path, ROOT)
        ref_path = os.path.join("/root/reference", rel)
        g = torch.load(ref_path)
        out[rel] = {"size": g["size"], "mask_sha": cases.sha(g["mask"]), "depth_sha": cases.sha(g["depth"]),
                    "levels": [len(r) for r in g["roots"]], "n_children": [sum(b) for b in g["branches"]]}
    torch.save(out, os.path.join(HERE, "growmaps_golden.pt"))
    print("growmaps_golden.pt", len(out))


if __name__ == "__main__":
    torch.set_num_threads(8)
    if "variants" in sys.argv[1:] or "sweep" in sys.argv[1:]:   # only the policy-variant / sweep-shape goldens
        if "variants" in sys.argv[1:]:
            gen_variants()
        if "sweep" in sys.argv[1:]:
            gen_sweep()
        sys.exit(0)
    gen_utils()
    gen_growmaps()
    gen_decode()
    gen_variants()
    gen_sweep()
