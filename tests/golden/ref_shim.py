"""Import the UNMODIFIED Sequoia reference from /root/reference in the build container.

Only used by ``make_golden.py`` (fixture generation) -- never at test/bench time, because
/root/reference does not exist on the GPU box.  The reference targets torch 2.1.2 /
transformers 4.36.2 / accelerate 0.26.1; this image has torch 2.11 / transformers 5.5 / no
accelerate, so five shims are applied WITHOUT editing the reference:

 1. import transformers' llama module first, then stub ``accelerate`` in sys.modules;
 2. monkey-patch ``Engine.Llama_modules.apply_rotary_pos_emb`` with the 4.36 semantics
    (q, k, cos, sin, position_ids) -- the reference keeps a verbatim copy of that function in
    Engine/offload_engine.py:35-67, which is what we bind;
 3. build LlamaConfig with rope_scaling=None / rope_theta=10000.0 explicitly;
 4. bypass from_pretrained(device_map=...) by constructing the engines with object.__new__;
 5. fp16 models on CPU, plain lambdas instead of the cuda_graph_for_* factories.
"""
import sys
import types

import torch

REF = "/root/reference"


def load_reference():
    import transformers.models.llama.modeling_llama  # noqa: F401  (1) before the accelerate stub
    if "accelerate" not in sys.modules:
        acc = types.ModuleType("accelerate")
        acc.cpu_offload = lambda model, execution_device=None: model
        acc.Accelerator = object
        sys.modules["accelerate"] = acc
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import Engine.Llama_modules as LM
    import Engine.offload_engine as OE
    LM.apply_rotary_pos_emb = OE.apply_rotary_pos_emb            # (2)
    import Engine.Engine as EE
    import Engine.Llama_model as LMod
    import Engine.Llama_KV as LKV
    import Tree.SpecTree as ST
    import Tree.GreedyTree as GT
    import Tree.GreedySTree as GST
    import Tree.SpecInferTree as SIT
    import utils as U
    LMod.LlamaForCausalLM_FI._tied_weights_keys = None
    LMod.LlamaForCausalLM_TG._tied_weights_keys = None
    return types.SimpleNamespace(LM=LM, EE=EE, LMod=LMod, LKV=LKV, ST=ST, GT=GT, GST=GST, SIT=SIT, U=U)


def hf_config(cfg):
    from transformers import LlamaConfig
    c = LlamaConfig(hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                    num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
                    num_key_value_heads=cfg.num_key_value_heads, vocab_size=cfg.vocab_size,
                    rms_norm_eps=cfg.rms_norm_eps, max_position_embeddings=cfg.max_position_embeddings,
                    rope_theta=cfg.rope_theta, attention_bias=False, hidden_act="silu")
    c.rope_scaling = None                                         # (3)
    c.rope_theta = cfg.rope_theta
    return c


def make_engine(ref, cfg, weights, max_length, kind):
    """(4) kind 'FI' -> GraphInferenceEngine (draft), 'TG' -> GraphInferenceEngineTG (target)."""
    hc = hf_config(cfg)
    model_cls = ref.LMod.LlamaForCausalLM_FI if kind == "FI" else ref.LMod.LlamaForCausalLM_TG
    model = model_cls(hc).to(torch.float16)
    missing, unexpected = model.load_state_dict(weights, strict=False)
    assert not unexpected, unexpected
    assert all("rotary" in m or "inv_freq" in m for m in missing), missing
    model.eval()
    inner_cls = ref.EE.InferenceEngine if kind == "FI" else ref.EE.InferenceEngineTG
    inner = object.__new__(inner_cls)
    inner.device, inner.dtype, inner.max_length = "cpu", torch.float16, max_length
    inner.model, inner.model_config = model, hc
    inner.kv_cache = ref.LKV.KV_Cache(config=hc, max_length=max_length, device="cpu", dtype=torch.float16)
    outer_cls = ref.EE.GraphInferenceEngine if kind == "FI" else ref.EE.GraphInferenceEngineTG
    outer = object.__new__(outer_cls)
    outer.device, outer.dtype, outer.max_length = "cpu", torch.float16, max_length
    outer.engine = inner
    if kind == "FI":
        outer.callables, outer.mempool = {}, None
    return outer
