#!/usr/bin/env python
"""The UNMODIFIED reference (oracle/_ref, vendored by tools/vendor_ref.py) driven on a CUDA device.

THIS IS TEST / MEASUREMENT INFRASTRUCTURE, NOT PRODUCT CODE: it gives bench.py the number BASELINE.json's north_star
sets as the bar -- "the reference's own GPU tokens/sec on the same B200" -- and gives the parity tests a full-size
GPU trace of the reference.  Run as a SEPARATE PROCESS (the reference's top-level module names Engine / Tree / utils
collide with this repository's drop-in shims of the same names):

    python oracle/ref_gpu.py --spec '{"draft": "llama-68m", "target": "llama-2-7b", ...}' --steps 20 --warmup 3

It replicates tests/testbed.py of the reference call for call:
  * engines            tests/testbed.py:250-254  (GraphInferenceEngine / GraphInferenceEngineTG; from_pretrained is
                       bypassed because there is no hub / accelerate here: the model classes are instantiated on the
                       device and given the same random-init weights the sequoia_b200 arm uses)
  * residual / sampling CUDA graphs, draft graphs   tests/testbed.py:256-285
  * the decode loop    tests/testbed.py:45-95 (simulation_fast), with a step budget and CUDA-event timing added
The five compatibility shims (torch 2.11 / transformers 5.5 instead of 2.1.2 / 4.36.2) are those of
tests/golden/ref_shim.py; no reference file is edited.

Prints ONE JSON line.  With --trace FILE it also saves, for the first --parity prompts, the first-iteration drafted
tree tokens / accept length / accepted tokens (full-size parity evidence consumed by bench.py).
"""
import argparse
import json
import os
import sys
import time
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_DIR = os.path.join(HERE, "_ref")


def load_reference(ref_dir=REF_DIR):
    if not os.path.isfile(os.path.join(ref_dir, "utils.py")):
        raise FileNotFoundError(f"{ref_dir} is empty: run tools/vendor_ref.py in the build container")
    import transformers.models.llama.modeling_llama  # noqa: F401  (shim 1: before the accelerate stub)
    if "accelerate" not in sys.modules:
        acc = types.ModuleType("accelerate")
        acc.cpu_offload = lambda model, execution_device=None: model
        acc.Accelerator = object
        sys.modules["accelerate"] = acc
    sys.path.insert(0, ref_dir)
    import Engine.Llama_modules as LM
    import Engine.offload_engine as OE
    LM.apply_rotary_pos_emb = OE.apply_rotary_pos_emb            # shim 2: transformers-4.36 RoPE signature
    import Engine.Engine as EE
    import Engine.Llama_model as LMod
    import Engine.Llama_KV as LKV
    import Tree.SpecTree as ST
    import Tree.GreedyTree as GT
    import utils as U
    for m in (LM, EE, ST, GT, U):
        assert os.path.abspath(m.__file__).startswith(os.path.abspath(ref_dir)), m.__file__
    LMod.LlamaForCausalLM_FI._tied_weights_keys = None
    LMod.LlamaForCausalLM_TG._tied_weights_keys = None
    return types.SimpleNamespace(LM=LM, EE=EE, LMod=LMod, LKV=LKV, ST=ST, GT=GT, U=U)


def hf_config(cfg):
    from transformers import LlamaConfig
    c = LlamaConfig(hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                    num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
                    num_key_value_heads=cfg.num_key_value_heads, vocab_size=cfg.vocab_size,
                    rms_norm_eps=cfg.rms_norm_eps, max_position_embeddings=cfg.max_position_embeddings,
                    rope_theta=cfg.rope_theta, attention_bias=False, hidden_act="silu")
    c.rope_scaling = None                                         # shim 3
    c.rope_theta = cfg.rope_theta
    return c


def make_engine(ref, cfg, weights, max_length, kind, device):
    """shim 4: 'FI' -> GraphInferenceEngine (draft), 'TG' -> GraphInferenceEngineTG (target), weights given."""
    hc = hf_config(cfg)
    model_cls = ref.LMod.LlamaForCausalLM_FI if kind == "FI" else ref.LMod.LlamaForCausalLM_TG
    # Construct on the meta device and ASSIGN the given tensors (what from_pretrained(device_map=...) does through
    # accelerate): a 70B model must not exist twice in HBM.  Buffers created in __init__ (the RoPE tables, non-persistent)
    # are then rebuilt on the real device by re-running the reference's own rotary module constructor.
    with torch.device("meta"):
        model = model_cls(hc)
    missing, unexpected = model.load_state_dict(weights, strict=False, assign=True)
    assert not unexpected, unexpected
    assert all("rotary" in m or "inv_freq" in m for m in missing), missing
    for mod in model.modules():
        if hasattr(mod, "rotary_emb") and hasattr(mod, "_init_rope"):
            with torch.device(device):
                mod._init_rope()
            mod.rotary_emb.to(torch.float16)                      # model.to(float16) casts the cached tables likewise
    bad = [n for n, t in list(model.named_parameters()) + list(model.named_buffers()) if t.is_meta]
    assert not bad, f"still on meta: {bad[:4]}"
    model.eval()
    inner_cls = ref.EE.InferenceEngine if kind == "FI" else ref.EE.InferenceEngineTG
    inner = object.__new__(inner_cls)
    inner.device, inner.dtype, inner.max_length = device, torch.float16, max_length
    inner.model, inner.model_config = model, hc
    inner.kv_cache = ref.LKV.KV_Cache(config=hc, max_length=max_length, device=device, dtype=torch.float16)
    outer_cls = ref.EE.GraphInferenceEngine if kind == "FI" else ref.EE.GraphInferenceEngineTG
    outer = object.__new__(outer_cls)
    outer.device, outer.dtype, outer.max_length = device, torch.float16, max_length
    outer.engine = inner
    if kind == "FI":
        outer.callables, outer.mempool = {}, None
    return outer


def testbed_setup(ref, draft, grow_map, M, T, greedy, device):
    """tests/testbed.py:256-285 (tests/testbed_greedy.py for the greedy policy): the reference's own CUDA graphs."""
    U = ref.U
    if not str(device).startswith("cuda"):          # --device cpu: plumbing self-test only (tests/test_ref_gpu_cpu.py)
        branch_lists = grow_map["branches"]
        n = len(grow_map["roots"]) - 1
        samp = {i: (lambda k: (lambda lg, rd=None: U.sampling_argmax(lg, k) if greedy else
                               U.sampling_without_replacement(lg, rd, k, T)))(max(branch_lists[i])) for i in range(n)}
        gather = {i: torch.cat([torch.arange(b) + j * max(branch_lists[i]) for j, b in enumerate(branch_lists[i])]).long()
                  for i in range(n)}
        return (lambda p, q: U.get_residual(p, q)), samp, gather
    residual_graph = U.cuda_graph_for_residual(device=device)
    idx_lists, branch_lists = grow_map["roots"], grow_map["branches"]
    draft_step = len(idx_lists)
    graph_capture_list = [sum(x) for x in branch_lists]
    graph_capture_list.append(1)
    draft.initialize_cuda_graph(graph_capture_list)
    sampling_callables, sample_gather_indices = {}, {}
    factory = U.cuda_graph_for_sampling_argmax if greedy else U.cuda_graph_for_sampling_without_replacement
    for i in range(draft_step - 1):
        sampling_callables[i] = factory(device=device, max_length=M, idx_len=len(idx_lists[i]),
                                        num_samples=max(branch_lists[i]), temperature=T, tree_size=grow_map["size"])
    for i in range(draft_step - 1):
        k = max(branch_lists[i])
        sample_gather_indices[i] = torch.cat([torch.arange(b, device=device, dtype=torch.long) + j * k
                                              for j, b in enumerate(branch_lists[i])])
    return residual_graph, sampling_callables, sample_gather_indices


def run(spec, steps, warmup, n_parity, trace_path, device="cuda:0", sdp="no_cudnn"):
    # SDPA backend selection (not a code change): the reference was written for torch 2.1.2, whose dispatcher knew flash /
    # mem-efficient / math.  torch 2.11 adds a cuDNN backend that faults ("misaligned address") on the strided,
    # arbitrarily offset fp16 mask views Tree/SpecTree.py:116-123 hands to F.scaled_dot_product_attention.
    if str(device).startswith("cuda"):
        if sdp in ("no_cudnn", "math"):
            torch.backends.cuda.enable_cudnn_sdp(False)
        if sdp == "math":
            torch.backends.cuda.enable_flash_sdp(False)
            torch.backends.cuda.enable_mem_efficient_sdp(False)
    ref = load_reference()
    sys.path.append(ROOT)                           # appended: Engine / Tree / utils stay the reference's
    from sequoia_b200.model import NAMED_CONFIGS, _RandomInit, full_state_dict
    assert sys.modules["utils"].__file__.startswith(os.path.abspath(REF_DIR))
    cuda = str(device).startswith("cuda")
    if cuda:
        torch.cuda.set_device(device)
    sync = torch.cuda.synchronize if cuda else (lambda: None)
    M, T, top_p, greedy = spec["M"], spec["T"], spec["top_p"], spec["greedy"]
    prefix, max_len = spec["prefix"], spec["max_len"]
    grow_map = torch.load(os.path.join(ROOT, spec["growmap"]))
    S = grow_map["size"]
    t0 = time.time()
    engines = []
    for name, seed, kind in ((spec["draft"], spec.get("draft_seed", 1), "FI"), (spec["target"], spec.get("target_seed", 2), "TG")):
        cfg = spec["_cfgs"][name] if "_cfgs" in spec else NAMED_CONFIGS[name]
        sd = full_state_dict(cfg, _RandomInit(cfg, seed, torch.device(device)))
        engines.append(make_engine(ref, cfg, sd, M, kind, device))
        del sd
    draft, target = engines
    residual_graph, sampling_callables, sample_gather_indices = testbed_setup(ref, draft, grow_map, M, T, greedy, device)
    init_s = time.time() - t0
    g = torch.Generator().manual_seed(17)
    prompts = [torch.randint(3, 32000, (prefix,), generator=g) for _ in range(64)]      # data_converter.synthetic_prompts
    dtype = torch.float16
    attn_mask = torch.full((M, M), torch.finfo(dtype).min, dtype=dtype, device=device)   # tests/testbed.py:53-57
    sequence = torch.tensor(list(range(M)), device=device).long().unsqueeze(-1)
    new_tokens_buffer = torch.zeros(M).long().to(device)
    parents_buffer = torch.zeros(M).long().to(device)
    position_ids = torch.zeros(M).long().to(device)
    cls = ref.GT.GreedyTree if greedy else ref.ST.SpecTree

    def new_tree(pi):
        torch.manual_seed(1000 + pi)                # same per-prompt CPU stream as the sequoia_b200 arm: same r / rand
        attn_mask.fill_(torch.finfo(dtype).min)
        return cls(prefix=prompts[pi % len(prompts)].to(device), device=device, temperature=T, top_p=top_p,
                   draft_kv_len=0, target_kv_len=0, draft_model_engine=draft, target_model_engine=target, max_length=M,
                   max_target_seq=M, grow_map=grow_map, attn_mask=attn_mask, sequence=sequence,
                   new_tokens_buffer=new_tokens_buffer, parents_buffer=parents_buffer, position_ids=position_ids,
                   residual_graph=residual_graph, sampling_callables=sampling_callables,
                   sample_gather_indices=sample_gather_indices)

    # ---- parity trace: first iteration of the first prompts (untimed) --------------------------------------------------
    trace = []
    with torch.no_grad():
        for pi in range(n_parity):
            tree = new_tree(pi)
            P = prefix
            tree.construct_grow_map()
            tokens = tree.tokens[P:P + S - 1].cpu().clone()
            valid, a, _, term = tree.verify()
            trace.append({"prompt": pi, "tree_tokens": tokens, "accept_len": int(a), "terminal": bool(term),
                          "valid_tokens": valid[:a].cpu().clone(),
                          "target_logits_head": tree.target_logits[:, :64].float().cpu().clone()
                          if hasattr(tree, "target_logits") else None})
            draft.clear_kv()
            target.clear_kv()
    if trace_path:
        os.makedirs(os.path.dirname(os.path.abspath(trace_path)), exist_ok=True)
        torch.save(trace, trace_path)

    # ---- the timed loop: tests/testbed.py:59-93 with a step budget ------------------------------------------------------
    state = {"pi": 0, "tree": None, "len": 0, "term": True}

    def run_steps(k):
        done = tokens = 0
        wall = ev_ms = 0.0
        with torch.no_grad():
            while done < k:
                if state["term"] or state["len"] >= max_len:
                    if state["tree"] is not None:
                        draft.clear_kv()
                        target.clear_kv()
                    state["tree"] = new_tree(state["pi"])
                    state["pi"] += 1
                    state["len"], state["term"] = prefix, False
                tree = state["tree"]
                if cuda:
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                sync()
                t1 = time.time()
                if cuda:
                    e0.record()
                while done < k and state["len"] < max_len and not state["term"]:
                    tree.construct_grow_map()
                    valid, _, _, term = tree.verify()
                    tokens += valid.shape[0] - state["len"]
                    state["len"] = valid.shape[0]
                    state["term"] = bool(term) or int(valid[-1]) in (0, 2)
                    done += 1
                if cuda:
                    e1.record()
                sync()
                wall += time.time() - t1
                ev_ms += e0.elapsed_time(e1) if cuda else (time.time() - t1) * 1e3
        return tokens, wall, ev_ms

    run_steps(max(warmup, 3))
    tokens, wall, ev_ms = run_steps(steps)
    return {"impl": "reference_gpu", "value": round(tokens / (ev_ms / 1e3), 2), "unit": "tokens/s",
            "ms_per_step": round(ev_ms / steps, 4), "wall_ms_per_step": round(wall / steps * 1e3, 4), "steps": steps,
            "accepted_tokens_per_step": round(tokens / steps, 4), "init_s": round(init_s, 1),
            "first_iter_accept_lens": [t["accept_len"] for t in trace],
            "how": "unmodified reference (oracle/_ref) on this GPU through tests/testbed.py's own setup: its CUDA graphs for "
                   "the draft widths / sampling / residual, eager target forward, host accept loop; same random-init "
                   "weights, prompts and per-prompt seeds as the sequoia_b200 arm",
            "torch": torch.__version__, "sdpa_backends": sdp}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--spec", required=True, help="JSON: draft, target, growmap, greedy, T, top_p, M, prefix, max_len")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--parity", type=int, default=4)
    ap.add_argument("--trace", default="")
    ap.add_argument("--device", default="cuda:0")
    ap.add_argument("--sdp", default="no_cudnn", choices=["auto", "no_cudnn", "math"])
    a = ap.parse_args()
    try:
        out = run(json.loads(a.spec), a.steps, a.warmup, a.parity, a.trace, a.device, a.sdp)
    except Exception as e:  # the bench line must survive a reference that does not run on this software stack
        import traceback
        out = {"impl": "reference_gpu", "unavailable": f"{type(e).__name__}: {e}", "traceback": traceback.format_exc()[-1500:]}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
