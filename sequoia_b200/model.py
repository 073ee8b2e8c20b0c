"""Tree-masked Llama forward on the sequoia_b200 kernels.

Replaces Engine/Llama_model.py + Engine/Llama_modules.py of the reference (LlamaForCausalLM_FI / _TG):
embed -> L x [RMSNorm, fused QKV GEMM, RoPE + KV append, tree attention, o_proj, residual+RMSNorm,
fused gate/up GEMM, SiLU*up, down_proj, residual+RMSNorm] -> lm_head.

* the dense weight GEMMs stay on cuBLASLt through torch.mm (not a hand-written
  kernel on this path); everything else is a launch into libsequoia_b200.so;
* q/k/v and gate/up weights are concatenated at load time so one GEMM feeds each fused kernel;
* every buffer is preallocated for n_max rows, so a forward allocates nothing and is CUDA-graph
  capturable; dynamic quantities (prefix length, kv length) are read from the device state word;
* optional tensor parallelism (Megatron layout): column-parallel qkv / gate_up, row-parallel
  o_proj / down_proj followed by an NCCL sum-allreduce (2 per layer), KV cache sharded by kv head.
"""
from __future__ import annotations

import json
import math
import os
from dataclasses import dataclass
from typing import Dict, Optional

import torch

from . import ops

F16 = torch.float16


@dataclass
class LlamaConfigLite:
    hidden_size: int
    intermediate_size: int
    num_hidden_layers: int
    num_attention_heads: int
    num_key_value_heads: int
    vocab_size: int = 32000
    rms_norm_eps: float = 1e-6
    rope_theta: float = 10000.0
    max_position_embeddings: int = 2048

    @property
    def head_dim(self):
        return self.hidden_size // self.num_attention_heads


# public HF configs of the model sizes BASELINE.json names (random-init weights of these shapes)
NAMED_CONFIGS = {
    "llama-68m": LlamaConfigLite(768, 3072, 2, 12, 12),
    "llama-160m": LlamaConfigLite(768, 3072, 12, 12, 12),
    "llama-2-7b": LlamaConfigLite(4096, 11008, 32, 32, 32, rms_norm_eps=1e-5, max_position_embeddings=4096),
    "llama-2-13b": LlamaConfigLite(5120, 13824, 40, 40, 40, rms_norm_eps=1e-5, max_position_embeddings=4096),
    "llama-2-70b": LlamaConfigLite(8192, 28672, 80, 64, 8, rms_norm_eps=1e-5, max_position_embeddings=4096),
    # the first 8 layers' worth of a 70B-shaped model: TP parity checks at the real head / FFN shapes where the
    # unsharded 138 GB model cannot sit next to a shard (bench.py tp_parity, tests/test_gpu_tp.py)
    "llama-2-70b-8l": LlamaConfigLite(8192, 28672, 8, 64, 8, rms_norm_eps=1e-5, max_position_embeddings=4096),
}


def config_from(obj) -> LlamaConfigLite:
    if isinstance(obj, LlamaConfigLite):
        return obj
    g = (lambda k, d=None: obj.get(k, d)) if isinstance(obj, dict) else (lambda k, d=None: getattr(obj, k, d))
    theta = g("rope_theta", None)
    return LlamaConfigLite(
        hidden_size=g("hidden_size"), intermediate_size=g("intermediate_size"),
        num_hidden_layers=g("num_hidden_layers"), num_attention_heads=g("num_attention_heads"),
        num_key_value_heads=g("num_key_value_heads") or g("num_attention_heads"), vocab_size=g("vocab_size", 32000),
        rms_norm_eps=g("rms_norm_eps", 1e-6), rope_theta=float(theta) if theta else 10000.0,
        max_position_embeddings=g("max_position_embeddings", 2048))


def _load_state_dict_dir(path: str) -> Dict[str, torch.Tensor]:
    sd: Dict[str, torch.Tensor] = {}
    files = sorted(os.listdir(path))
    st = [f for f in files if f.endswith(".safetensors")]
    if st:
        from safetensors.torch import load_file
        for f in st:
            sd.update(load_file(os.path.join(path, f)))
        return sd
    for f in files:                                    # HF shard names only (a model dir may also hold training_args.bin ...)
        if (f.startswith("pytorch_model") and f.endswith(".bin")) or (f.startswith("model") and f.endswith(".pt")):
            sd.update(torch.load(os.path.join(path, f), map_location="cpu"))
    if not sd:
        raise FileNotFoundError(f"no weight files (*.safetensors / *.bin) in {path}")
    return sd


class _RandomInit:
    """HF-default random init (normal std 0.02, norms = 1) generated on the device, tensor by tensor, from a seeded
    generator; the FULL tensor is always drawn and then sharded, so every TP degree sees the same model."""

    def __init__(self, cfg: LlamaConfigLite, seed: int, device):
        self.cfg, self.device = cfg, device
        self.g = torch.Generator(device=device)
        self.g.manual_seed(seed)

    def get(self, name: str, shape) -> torch.Tensor:
        if name.endswith("norm.weight") or "layernorm" in name:
            return torch.ones(shape, dtype=F16, device=self.device)
        t = torch.empty(shape, dtype=F16, device=self.device)
        t.normal_(0.0, 0.02, generator=self.g)
        return t


class _DictSource:
    def __init__(self, sd: Dict[str, torch.Tensor], device):
        self.sd, self.device = sd, device

    def get(self, name: str, shape) -> torch.Tensor:
        if name == "lm_head.weight" and name not in self.sd:          # tied embeddings: checkpoints omit lm_head
            name = "model.embed_tokens.weight"
        t = self.sd[name]
        assert tuple(t.shape) == tuple(shape), (name, t.shape, shape)
        return t.to(device=self.device, dtype=F16)


def resolve_model(model_name_or_path, device):
    """-> (LlamaConfigLite, weight source).  Accepted forms of `model_name_or_path`:
       * directory with config.json + *.safetensors / *.bin   (what from_pretrained took, Engine/Engine.py:18)
       * "random-init:<name>[:seed]" with <name> in NAMED_CONFIGS   (synthetic benchmarks, no network)
       * {"config": cfg, "state_dict": {...}}   (tests: weights shared with the oracle)."""
    if isinstance(model_name_or_path, dict):
        cfg = config_from(model_name_or_path["config"])
        return cfg, _DictSource(model_name_or_path["state_dict"], device)
    s = str(model_name_or_path)
    if s.startswith("random-init:"):
        parts = s.split(":")
        cfg = NAMED_CONFIGS[parts[1].lower()]
        seed = int(parts[2]) if len(parts) > 2 else 0
        return cfg, _RandomInit(cfg, seed, device)
    if os.path.isdir(s):
        with open(os.path.join(s, "config.json")) as f:
            cfg = config_from(json.load(f))
        return cfg, _DictSource(_load_state_dict_dir(s), device)
    raise FileNotFoundError(
        f"{s!r}: expected a local model directory or 'random-init:<{'|'.join(NAMED_CONFIGS)}>[:seed]' "
        "(there is no network access for hub downloads)")


def rope_cache(cfg: LlamaConfigLite, max_length: int, device):
    """LlamaRotaryEmbedding_FI (Engine/Llama_modules.py:16-45): fp32 tables, sliced [:max_length], cast to fp16."""
    d = cfg.head_dim
    inv_freq = 1.0 / (cfg.rope_theta ** (torch.arange(0, d, 2, dtype=torch.float32) / d))
    # the reference builds max_position_embeddings rows and slices [:max_length]; rows beyond that table would be an
    # out-of-bounds read in the RoPE kernel, so the table always covers max_length (identical values where both exist)
    t = torch.arange(max(cfg.max_position_embeddings, max_length), dtype=torch.float32)
    freqs = torch.outer(t, inv_freq)
    emb = torch.cat((freqs, freqs), dim=-1)
    return (emb.cos()[:max_length].to(F16).to(device).contiguous(), emb.sin()[:max_length].to(F16).to(device).contiguous())


class TPInfo:
    def __init__(self, group=None):
        import torch.distributed as dist
        self.group = group
        if group is None:
            self.rank, self.size = 0, 1
        else:
            self.rank, self.size = dist.get_rank(group), dist.get_world_size(group)

    def all_reduce(self, t: torch.Tensor):
        if self.size > 1:
            import torch.distributed as dist
            dist.all_reduce(t, group=self.group)


def full_state_dict(cfg: LlamaConfigLite, src) -> Dict[str, torch.Tensor]:
    """The whole model as an HF-named state dict, tensors requested from `src` in EXACTLY the order
    `load_sharded_weights` requests them -- so a seeded `_RandomInit` yields the same weights either way.  Used to hand
    the very same random-init model to the reference implementation / the CPU oracle in bench.py."""
    h, D, V = cfg.hidden_size, cfg.head_dim, cfg.vocab_size
    Hf, Hkvf, If = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.intermediate_size
    sd = {"model.embed_tokens.weight": src.get("model.embed_tokens.weight", (V, h))}
    for l in range(cfg.num_hidden_layers):
        p = f"model.layers.{l}."
        for name, shape in (("self_attn.q_proj.weight", (Hf * D, h)), ("self_attn.k_proj.weight", (Hkvf * D, h)),
                            ("self_attn.v_proj.weight", (Hkvf * D, h)), ("self_attn.o_proj.weight", (h, Hf * D)),
                            ("mlp.gate_proj.weight", (If, h)), ("mlp.up_proj.weight", (If, h)),
                            ("mlp.down_proj.weight", (h, If)), ("input_layernorm.weight", (h,)),
                            ("post_attention_layernorm.weight", (h,))):
            sd[p + name] = src.get(p + name, shape)
    sd["model.norm.weight"] = src.get("model.norm.weight", (h,))
    sd["lm_head.weight"] = src.get("lm_head.weight", (V, h))
    return sd


def load_sharded_weights(cfg: LlamaConfigLite, src, rank: int, tp: int) -> dict:
    """Megatron-style shard `rank` of `tp`: q/k/v and gate/up split by output rows (heads / FFN columns) and fused,
    o_proj / down_proj split by input columns (their partial products are summed by the allreduce)."""
    h, D, V = cfg.hidden_size, cfg.head_dim, cfg.vocab_size
    Hf, Hkvf, If = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.intermediate_size
    assert Hf % tp == 0 and Hkvf % tp == 0 and If % tp == 0, "heads / FFN width must divide the TP degree"
    H, Hkv, I, r = Hf // tp, Hkvf // tp, If // tp, rank
    out = {"embed": src.get("model.embed_tokens.weight", (V, h)), "layers": []}
    for l in range(cfg.num_hidden_layers):
        p = f"model.layers.{l}."
        wq = src.get(p + "self_attn.q_proj.weight", (Hf * D, h))[r * H * D:(r + 1) * H * D]
        wk = src.get(p + "self_attn.k_proj.weight", (Hkvf * D, h))[r * Hkv * D:(r + 1) * Hkv * D]
        wv = src.get(p + "self_attn.v_proj.weight", (Hkvf * D, h))[r * Hkv * D:(r + 1) * Hkv * D]
        wo = src.get(p + "self_attn.o_proj.weight", (h, Hf * D))[:, r * H * D:(r + 1) * H * D]
        wg = src.get(p + "mlp.gate_proj.weight", (If, h))[r * I:(r + 1) * I]
        wu = src.get(p + "mlp.up_proj.weight", (If, h))[r * I:(r + 1) * I]
        wd = src.get(p + "mlp.down_proj.weight", (h, If))[:, r * I:(r + 1) * I]
        out["layers"].append(dict(
            wqkv=torch.cat([wq, wk, wv], dim=0).contiguous(), wo=wo.contiguous(),
            wgu=torch.cat([wg, wu], dim=0).contiguous(), wd=wd.contiguous(),
            ln1=src.get(p + "input_layernorm.weight", (h,)), ln2=src.get(p + "post_attention_layernorm.weight", (h,))))
        del wq, wk, wv, wo, wg, wu, wd
    out["norm"] = src.get("model.norm.weight", (h,))
    out["lm_head"] = src.get("lm_head.weight", (V, h))
    return out


class LlamaRunner:
    """Weights + preallocated activations + attention plan of one engine."""

    def __init__(self, model_name_or_path, max_length: int, device="cuda:0", tp_group=None, n_max: Optional[int] = None):
        self.device = torch.device(device)
        self.cfg, src = resolve_model(model_name_or_path, self.device)
        cfg = self.cfg
        self.tp = TPInfo(tp_group)
        tp, r = self.tp.size, self.tp.rank
        assert cfg.num_attention_heads % tp == 0 and cfg.num_key_value_heads % tp == 0 and cfg.intermediate_size % tp == 0
        self.M = max_length
        self.n_max = n_max or max_length
        self.H, self.Hkv, self.D = cfg.num_attention_heads // tp, cfg.num_key_value_heads // tp, cfg.head_dim
        self.I = cfg.intermediate_size // tp
        h, D, V = cfg.hidden_size, self.D, cfg.vocab_size
        self.h, self.V, self.L = h, V, cfg.num_hidden_layers
        wts = load_sharded_weights(cfg, src, r, tp)
        self.embed, self.layers, self.norm, self.lm_head = wts["embed"], wts["layers"], wts["norm"], wts["lm_head"]
        self.cos, self.sin = rope_cache(cfg, max_length, self.device)
        self.eps = float(cfg.rms_norm_eps)
        n = self.n_max
        dev = self.device
        z = lambda *s: torch.zeros(*s, dtype=F16, device=dev)
        self.hidden, self.normed = z(n, h), z(n, h)
        self.qkv = z(n, (self.H + 2 * self.Hkv) * D)
        self.attn_out, self.proj = z(n, self.H * D), z(n, h)
        self.gate_up, self.act = z(n, 2 * self.I), z(n, self.I)
        self.logits = z(n, V)
        self.k_cache = torch.zeros(self.L, 1, self.Hkv, max_length, D, dtype=F16, device=dev)
        self.v_cache = torch.zeros_like(self.k_cache)
        self.plan = ops.AttnPlan(self.qkv, n, self.H, self.Hkv, D, self.k_cache, self.v_cache, self.attn_out)
        # Dense GEMMs.  Default ("auto"): the weight-streaming shapes that the hand-written tcgen05 kernel (csrc/sq_gemm.cu)
        # wins -- gate_up with the SwiGLU epilogue fused (weights interleaved + pre-tiled, `act` written directly: no
        # gate_up round trip, no silu_mul launch) and the target's lm_head (pre-tiled) -- for models whose layers are
        # HBM-stream-sized (hidden >= 2048); q/k/v, o_proj and down_proj stay on cuBLASLt (torch.mm), whose 2-CTA-MMA
        # kernels are faster at those shapes (profiles/r02_gemm_ncu.md).  SQ_GEMM=1 routes those through sq_gemm too
        # (tuning), SQ_GEMM=0 keeps everything on cuBLASLt + sq_silu_mul.
        mode = os.environ.get("SQ_GEMM", "auto")
        self.gemm = None
        self.gemm_err = torch.zeros(4, dtype=torch.int32, device=dev)
        self.lm_plan = None
        stream_sized = h >= 2048
        if mode == "1":
            self.gemm = []
            for ly in self.layers:
                self.gemm.append(dict(qkv=self._plan(self.normed, ly["wqkv"], self.qkv),
                                      o=self._plan(self.attn_out, ly["wo"], self.proj),
                                      d=self._plan(self.act, ly["wd"], self.proj)))
        # (a fused-epilogue plan cannot split K, so a narrow shard -- TP-4/8 of a 7B -- would leave most SMs idle: those
        # stay on cuBLASLt + sq_silu_mul).  The plans serve forwards of <= 128 rows; larger ones (prefill, the 768-row verify
        # of config 4: compute-bound, not a weight stream) go to cuBLASLt on the SAME weight tensor, which is why gate_up is
        # kept row-major in the interleaved order (16 gate rows | 16 up rows) rather than pre-tiled.
        gu_bn, gu_split, _ = ops.gemm_pick_tiles(2 * self.I, h, ops.GEMM_SWIGLU) if self.I % 16 == 0 and h % 64 == 0 else (0, 1, 1)
        gu_ctas = (-(-2 * self.I // gu_bn)) * gu_split if gu_bn else 0
        self.gu_interleaved = False
        if mode in ("1", "auto") and stream_sized and gu_ctas >= 120:
            self.gu_interleaved = True
            for ly in self.layers:
                ly["wgu"] = ops.interleave_gate_up(ly["wgu"][:self.I], ly["wgu"][self.I:])
                ly["gu_plan"] = ops.GemmPlan(self.normed, ly["wgu"], self.act, self.gemm_err, swiglu=True)
        if mode in ("1", "auto") and stream_sized and V % 32 == 0 and h % 64 == 0:
            self.lm_plan = ops.GemmPlan(self.normed, self.lm_head, self.logits, self.gemm_err)
        # Small draft models (csrc/sq_draft.cu).  SQ_DRAFT_FUSED = "chain" | "coop": the whole forward of a tree level as
        # PDL-chained phase kernels / ONE persistent cooperative kernel (measured SLOWER than the multi-kernel path on a
        # B200 -- 103 / 121 us vs 84 us per level, DESIGN.md section 7 -- so it is opt-in).  Default: only its attention
        # phase replaces sq_tree_attn for the draft's tree-relative forwards of <= 64 rows (SQ_DRAFT_ATTN=0 turns that off).
        self.draft_plan = None
        self.draft_fused = os.environ.get("SQ_DRAFT_FUSED", "0") not in ("0", "")
        self.draft_attn = os.environ.get("SQ_DRAFT_ATTN", "1") != "0"
        if ((self.draft_fused or self.draft_attn) and tp == 1 and not self.gu_interleaved and
                ops.draft_supported(h, self.I, self.L, self.H, self.Hkv, D, V, max_length)):
            self.draft_plan = ops.DraftPlan(h, self.I, self.H, V, max_length, self.eps, self.embed, self.layers, self.norm,
                                            self.lm_head, self.cos, self.sin, self.k_cache, self.v_cache)
        self.peer = None
        if self.tp.size > 1 and os.environ.get("SQ_TP_MODE", "fused") == "fused":
            from .peer import PeerBuffers
            self.peer = PeerBuffers(tp_group, self.device, n, h)
        self.attn_impl = int(os.environ.get("SQ_ATTN_IMPL", "0"))
        # L2 prefetch of the next GEMM's weights while the latency-bound kernels between two GEMMs leave HBM idle
        # (csrc/sq_prefetch.cu).  SQ_L2_PREFETCH = "A,B,C,D" MB budgets for the four windows of a layer (after qkv /
        # o_proj / gate_up / down_proj), "0" = off.  A hint only: results are identical with or without it.
        pf = os.environ.get("SQ_L2_PREFETCH", "0")
        self.pf_budget = None
        if pf not in ("0", "") and self.peer is None and h >= 2048:      # pointless for the small draft models
            vals = [74, 24, 28, 28] if pf == "1" else [float(x) for x in pf.split(",")]
            assert len(vals) == 4, "SQ_L2_PREFETCH: 1 | 0 | A,B,C,D (MB)"
            self.pf_budget = [int(v * 1e6) for v in vals]
            self.pf_stream = torch.cuda.Stream(device=self.device)

    def _prefetch(self, window: int, weights):
        """Fork: on the side stream, pull the leading columns of `weights` (in order, until the window's byte budget is
        spent) into L2.  Joined by `_prefetch_join` before the next GEMM."""
        if self.pf_budget is None:
            return
        left = self.pf_budget[window]
        cur = torch.cuda.current_stream()
        self.pf_stream.wait_stream(cur)
        with torch.cuda.stream(self.pf_stream):
            for w, col0 in weights:
                row_bytes = w.shape[0] * w.element_size()
                cols = min((left // row_bytes) // 64 * 64, w.shape[1] - col0)
                if cols <= 0:
                    continue
                ops.l2_prefetch(w, col0, col0 + cols)
                left -= cols * row_bytes
        self._pf_fork = True

    def _prefetch_join(self):
        if self.pf_budget is not None and getattr(self, "_pf_fork", False):
            torch.cuda.current_stream().wait_stream(self.pf_stream)
            self._pf_fork = False

    def _plan(self, a, w, c):
        try:
            return ops.GemmPlan(a, w, c, self.gemm_err)
        except Exception:
            return None                                   # shape outside the kernel's tiling (K % 64, N % 32): cuBLASLt

    def _gate_up_act(self, ly, n: int):
        """act[:n] = silu(normed[:n] @ Wg.T) * (normed[:n] @ Wu.T)   (Engine/Llama_modules.py:272)"""
        plan = ly.get("gu_plan")
        # the plan's activation tile is always 128 rows: worth it when most of them are real (config 2: 128 rows); for the
        # 65-row tree of config 3 cuBLASLt's 64-row tiles ingest half as much activation per SM (7.19 vs 7.36 ms / step)
        if plan is not None and 96 < n <= 128:
            plan.run(n)
            return
        torch.mm(self.normed[:n], ly["wgu"].t(), out=self.gate_up[:n])
        ops.silu_mul(self.gate_up, self.act, n, interleaved=self.gu_interleaved)

    def _linear(self, l: int, key: str, x: torch.Tensor, w: torch.Tensor, out: torch.Tensor, n: int):
        """out[:n] = x[:n] @ w.T"""
        if self.gemm is not None and n <= 128 and self.peer is None:
            plan = self.gemm[l][key]
            if plan is not None:
                plan.run(n)
                return
        torch.mm(x[:n], w.t(), out=out[:n])

    def weight_bytes(self) -> int:
        b = self.embed.numel() + self.lm_head.numel() + self.norm.numel()
        for ly in self.layers:
            b += sum(ly[k].numel() for k in ("wqkv", "wo", "wgu", "wd", "ln1", "ln2"))
        return 2 * b

    @torch.no_grad()
    def forward(self, n: int, tokens: torch.Tensor, position_ids: torch.Tensor, storage_ids: torch.Tensor, *,
                state=None, n0: int = 0, kv_end: int = 0, prefix_len: int = 0, dense_mask=None, mask_ld: int = 0,
                tree_bits=None, tree_words: int = 0, tree_size: int = 0, logits_out: Optional[torch.Tensor] = None,
                logits_from: int = 0, skip_lm_head: bool = False) -> Optional[torch.Tensor]:
        """Forward `n` rows.  Row r is token tokens[base+r] at position position_ids[base+r], written to cache slot
        storage_ids[base+r], base = (state ? P-1 : 0) + n0.  Attends slots [0, (state ? P-1 : 0) + kv_end).
        Logits of rows [logits_from, n) are written to `logits_out` (default: the internal buffer) and returned."""
        assert 0 < n <= self.n_max
        H, Hkv, D, M = self.H, self.Hkv, self.D, self.M
        small = (self.draft_plan is not None and state is not None and n <= ops.DraftPlan.MAX_ROWS and dense_mask is None
                 and self.attn_impl == 0)
        if small and self.draft_fused and logits_from == 0 and not skip_lm_head:
            out = logits_out if logits_out is not None else self.logits[:n]
            self.draft_plan.forward(n, tokens, position_ids, storage_ids, state, n0, kv_end, tree_bits, tree_words,
                                    tree_size, out)
            return out
        hid, nrm = self.hidden[:n], self.normed[:n]
        ops.embed_rows(self.embed, tokens, n, self.hidden, state=state, n0=n0)
        ops.rmsnorm(self.hidden, self.layers[0]["ln1"], self.normed, n, self.eps)
        for l, ly in enumerate(self.layers):
            self._prefetch_join()
            self._linear(l, "qkv", self.normed, ly["wqkv"], self.qkv, n)
            if self.pf_budget is not None:          # window A: RoPE + attention
                self._prefetch(0, [(ly["wo"], 0), (ly["wgu"], 0)])
            ops.rope_kv_append(self.qkv, H, Hkv, D, self.cos, self.sin, position_ids, storage_ids, n,
                               self.k_cache[l], self.v_cache[l], M, state=state, n0=n0)
            if small and self.draft_attn:
                self.draft_plan.attention(l, n, self.qkv, self.attn_out, state, n0, kv_end, tree_bits, tree_words, tree_size)
            else:
                ops.tree_attn(self.plan, l, n, state=state, n0=n0, kv_end=kv_end, prefix_len=prefix_len,
                              dense_mask=dense_mask, mask_ld=mask_ld, tree_bits=tree_bits, tree_words=tree_words,
                              tree_size=tree_size, impl=self.attn_impl)
            nxt = self.layers[l + 1]["ln1"] if l + 1 < self.L else self.norm
            if self.peer is not None:
                torch.mm(self.attn_out[:n], ly["wo"].t(), out=self.peer.buf[0][:n])
                self.peer.allreduce_add_rmsnorm(0, self.hidden, ly["ln2"], self.normed, n, self.eps)
                self._gate_up_act(ly, n)
                torch.mm(self.act[:n], ly["wd"].t(), out=self.peer.buf[1][:n])
                self.peer.allreduce_add_rmsnorm(1, self.hidden, nxt, self.normed, n, self.eps)
                continue
            self._prefetch_join()
            self._linear(l, "o", self.attn_out, ly["wo"], self.proj, n)
            if self.pf_budget is not None:          # window B: residual + RMSNorm; continue gate_up where window A stopped
                wo_b = ly["wo"].numel() * 2
                done = 0 if self.pf_budget[0] <= wo_b else \
                    min(((self.pf_budget[0] - wo_b) // (ly["wgu"].shape[0] * 2)) // 64 * 64, ly["wgu"].shape[1])
                self._prefetch(1, [(ly["wgu"], done)])
            self.tp.all_reduce(self.proj[:n])
            ops.add_rmsnorm(self.hidden, self.proj, ly["ln2"], self.normed, n, self.eps)
            self._prefetch_join()
            self._gate_up_act(ly, n)
            self._prefetch_join()
            self._linear(l, "d", self.act, ly["wd"], self.proj, n)
            if self.pf_budget is not None:          # window D: residual + RMSNorm before the next layer's qkv / lm_head
                nw = self.layers[l + 1]["wqkv"] if l + 1 < self.L else (None if skip_lm_head else self.lm_head)
                if nw is not None:
                    self._prefetch(3, [(nw, 0)])
            self.tp.all_reduce(self.proj[:n])
            ops.add_rmsnorm(self.hidden, self.proj, nxt, self.normed, n, self.eps)
        if skip_lm_head:                               # TP follower ranks: only rank 0 consumes logits
            self._prefetch_join()
            return None
        self._prefetch_join()
        m = n - logits_from
        out = logits_out if logits_out is not None else self.logits[:m]
        if self.lm_plan is not None and m <= 128 and out.stride(-1) == 1 and out.stride(0) % 8 == 0 and out.data_ptr() % 16 == 0:
            self.lm_plan.run(m, a_row0=logits_from, out=out)
        else:
            torch.mm(self.normed[logits_from:n], self.lm_head.t(), out=out)
        return out
