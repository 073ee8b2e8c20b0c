"""CPU ORACLE for the Sequoia tree-speculative-decoding hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  The product (``sequoia_b200``) never imports anything from
``oracle/`` and fails loudly when its CUDA library is missing.

It is a plain torch-on-CPU restatement of the reference algorithm (the reference itself
is pure PyTorch, so "the same arithmetic" means the same torch ops in the same order and
dtype: fp16 tensors, fp32 softmax internals, fp16 roundings between ops).  Each function
cites the reference file:line it follows (paths relative to the Sequoia repository).

Parity pinning: the reference ships NO golden vectors / unit tests for this path
(SURVEY.md section 4), so this oracle is pinned against outputs of the reference itself,
imported in the build container with the 5-point compatibility shim and run on CPU:
``tests/golden/make_golden.py`` generated ``tests/golden/*.pt``; ``tests/test_oracle_golden.py``
checks this file against every one of them bit-exactly (same torch build => same bits).

Arithmetic that lives in third-party code (torch 2.1.2 / transformers 4.36.2 pinned by the
reference README:13-15; torch 2.11 / transformers 5.5 here): softmax, topk, multinomial,
matmul, index ops (torch) and apply_rotary_pos_emb / repeat_kv (transformers 4.36
semantics, restated below from the verbatim copy the reference keeps in
Engine/offload_engine.py:35-67).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F
from torch.nn.functional import softmax

FP16_MIN = torch.finfo(torch.float16).min  # -65504, the reference's "masked" value


# --------------------------------------------------------------------------------------
# utils.py
# --------------------------------------------------------------------------------------
def get_residual(p: torch.Tensor, q: torch.Tensor) -> torch.Tensor:
    """utils.py:5-8  relu(p-q) / sum(relu(p-q)) in the tensors' own dtype (fp16)."""
    residual = (p - q).relu_()
    residual = residual / (residual.sum(dim=-1).unsqueeze(-1))
    return residual


def sampling_without_replacement(sampling_logits: torch.Tensor, rand: torch.Tensor,
                                 num_samples: int, temperature: float) -> torch.Tensor:
    """utils.py:10-18  exponential-race sampling w/o replacement, everything in fp16."""
    sampling_q = softmax(sampling_logits / temperature, dim=-1)
    position = (rand.log() / sampling_q).topk(k=num_samples).indices.flatten()
    return position


def sampling_argmax(sampling_logits: torch.Tensor, num_samples: int) -> torch.Tensor:
    """utils.py:29-32."""
    return sampling_logits.topk(k=num_samples).indices.flatten()


def get_sampling_logits(logits: torch.Tensor, top_p: float, T: float, replicate: bool = False):
    """utils.py:65-77  nucleus filter; a no-op when top_p >= 1.0 (all named configs)."""
    if replicate:
        logits = logits.clone()
    if top_p < 1.0:
        sorted_logits, sorted_indices = torch.sort(logits, descending=True)
        cumulative_probs = torch.cumsum(softmax(sorted_logits / T, dim=-1), dim=-1)
        filt = cumulative_probs > top_p
        filt[..., 1:] = filt[..., :-1].clone()
        filt[..., 0] = 0
        indices_to_remove = filt.scatter(-1, sorted_indices, filt)
        logits[indices_to_remove] = float("-inf")
    return logits


def make_causal_mask(tgt_len: int, dtype=torch.float16) -> torch.Tensor:
    """utils.py:95-107 (_make_causal_mask): 0 on/below the diagonal, dtype-min above."""
    mask = torch.full((tgt_len, tgt_len), torch.finfo(dtype).min)
    cond = torch.arange(tgt_len)
    mask.masked_fill_(cond < (cond + 1).view(tgt_len, 1), 0)
    return mask.to(dtype)


def sample_gather_index(branches: Sequence[int]) -> torch.Tensor:
    """tests/testbed.py:277-285  indices selecting the first b_j draws of row j."""
    mx = max(branches)
    out = [torch.arange(b, dtype=torch.long) + j * mx for j, b in enumerate(branches)]
    return torch.cat(out) if out else torch.zeros(0, dtype=torch.long)


# --------------------------------------------------------------------------------------
# Tree/Tree.py + SpecTree.__init__ mask / position layout
# --------------------------------------------------------------------------------------
def build_full_attn_mask(max_length: int, tree_mask01: torch.Tensor, dtype=torch.float16) -> torch.Tensor:
    """Tree/Tree.py:13-27 + Tree/SpecTree.py:44-54.

    (2M,2M) additive mask: top-left MxM causal, the tree block (rows/cols 1..S-1 of the
    growmap mask; 1 = ancestor-or-self = visible) pasted at [M-S+1:M, M-S+1:M], all else min.
    """
    M = max_length
    S = tree_mask01.shape[0]
    full = torch.full((M, M), torch.finfo(dtype).min, dtype=dtype).repeat(2, 2)
    full[:M, :M] = make_causal_mask(M, dtype)
    tm = (tree_mask01 == 0).type(dtype)
    tm.masked_fill_(tm > 0, torch.finfo(dtype).min)
    full[M - S + 1:M, M - S + 1:M] = tm[1:, 1:]
    return full


def window_mask(full: torch.Tensor, max_length: int, total_nodes: int) -> torch.Tensor:
    """Tree/SpecTree.py:57-58,270-271  the sliding (M,M) window for tot = P + S - 1."""
    M = max_length
    return full[M - total_nodes: 2 * M - total_nodes, M - total_nodes: 2 * M - total_nodes]


def visible_from_rule(max_length: int, P: int, tree_mask01: torch.Tensor) -> torch.Tensor:
    """Closed form of the same window as a bool (tot, tot) matrix:
    slot c is visible from row slot r iff   c <= min(r, P-1)   or
    (r >= P and c >= P-1 and mask[r-(P-1), c-(P-1)] == 1).
    This is the rule the CUDA kernels evaluate instead of reading a dense fp16 mask."""
    S = tree_mask01.shape[0]
    tot = P + S - 1
    r = torch.arange(tot).view(-1, 1)
    c = torch.arange(tot).view(1, -1)
    vis = c <= torch.minimum(r, torch.tensor(P - 1))
    node_r = (r - (P - 1)).clamp(min=0)
    node_c = (c - (P - 1)).clamp(min=0)
    tree = tree_mask01.bool()[node_r.expand(tot, tot), node_c.expand(tot, tot)]
    vis = vis | ((r >= P) & (c >= P - 1) & tree)
    return vis


# --------------------------------------------------------------------------------------
# Engine/Llama_KV.py
# --------------------------------------------------------------------------------------
class KVCacheOracle:
    """Engine/Llama_KV.py:4-103  static (L,1,H_kv,M,D) K and V."""

    def __init__(self, num_layers: int, num_kv_heads: int, head_dim: int, max_length: int,
                 dtype=torch.float16):
        self.max_length = max_length
        self.num_layers = num_layers
        self.k_cache = torch.zeros(num_layers, 1, num_kv_heads, max_length, head_dim, dtype=dtype)
        self.v_cache = torch.zeros_like(self.k_cache)
        self.kv_offset = 0

    def initialize_kv(self, k_cache, v_cache, kv_len):            # :38-46
        self.k_cache[..., :kv_len, :] = k_cache[..., :kv_len, :]
        self.v_cache[..., :kv_len, :] = v_cache[..., :kv_len, :]
        self.kv_offset = kv_len

    def gather_kv(self, indices: List[int]):                       # :50-58
        n = len(indices)
        self.k_cache[..., :n, :] = self.k_cache[..., indices, :]
        self.v_cache[..., :n, :] = self.v_cache[..., indices, :]
        self.k_cache[..., n:, :] = 0.0
        self.v_cache[..., n:, :] = 0.0
        self.kv_offset = n

    def gather_kv_incremental(self, indices: List[int], offset: int):   # :60-68
        n = len(indices)
        self.k_cache[..., offset:offset + n, :] = self.k_cache[..., indices, :]
        self.v_cache[..., offset:offset + n, :] = self.v_cache[..., indices, :]
        self.k_cache[..., offset + n:, :] = 0.0
        self.v_cache[..., offset + n:, :] = 0.0
        self.kv_offset = offset + n

    def update_kv_cache(self, new_k, new_v, layer_idx: int, storage_ids: torch.Tensor):   # :72-89
        self.k_cache[layer_idx].index_copy_(dim=-2, index=storage_ids, source=new_k)
        self.v_cache[layer_idx].index_copy_(dim=-2, index=storage_ids, source=new_v)
        if layer_idx == self.num_layers - 1:
            self.kv_offset += len(storage_ids)
        return self.k_cache[layer_idx], self.v_cache[layer_idx]

    def clear(self):                                               # :91-94
        self.k_cache.zero_()
        self.v_cache.zero_()
        self.kv_offset = 0

    def get_usable_length(self, layer_idx: int, input_length: int) -> int:   # :96-100
        if layer_idx == self.num_layers - 1:
            return self.kv_offset
        return self.kv_offset + input_length

    def set_kv_len(self, kv_len: int):                             # :102-103
        self.kv_offset = kv_len


# --------------------------------------------------------------------------------------
# Engine/Llama_modules.py + Engine/Llama_model.py  (functional, weights in a dict)
# --------------------------------------------------------------------------------------
@dataclass
class LlamaCfg:
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
    def head_dim(self) -> int:
        return self.hidden_size // self.num_attention_heads


def init_llama_weights(cfg: LlamaCfg, seed: int, dtype=torch.float16, std: float = 0.02,
                       generator_device: str = "cpu") -> Dict[str, torch.Tensor]:
    """HF-default random init (normal std 0.02, RMSNorm weight 1), drawn in
    fp32 from a seeded CPU generator in a FIXED key order, then cast.  Both the oracle and
    the product load the dict this returns, so they see identical weights."""
    g = torch.Generator(device=generator_device)
    g.manual_seed(seed)

    def nrm(*shape):
        return (torch.randn(*shape, generator=g, dtype=torch.float32) * std).to(dtype)

    h, i, v = cfg.hidden_size, cfg.intermediate_size, cfg.vocab_size
    d = cfg.head_dim
    w: Dict[str, torch.Tensor] = {}
    w["model.embed_tokens.weight"] = nrm(v, h)
    for l in range(cfg.num_hidden_layers):
        p = f"model.layers.{l}."
        w[p + "self_attn.q_proj.weight"] = nrm(cfg.num_attention_heads * d, h)
        w[p + "self_attn.k_proj.weight"] = nrm(cfg.num_key_value_heads * d, h)
        w[p + "self_attn.v_proj.weight"] = nrm(cfg.num_key_value_heads * d, h)
        w[p + "self_attn.o_proj.weight"] = nrm(h, cfg.num_attention_heads * d)
        w[p + "mlp.gate_proj.weight"] = nrm(i, h)
        w[p + "mlp.up_proj.weight"] = nrm(i, h)
        w[p + "mlp.down_proj.weight"] = nrm(h, i)
        w[p + "input_layernorm.weight"] = torch.ones(h, dtype=dtype)
        w[p + "post_attention_layernorm.weight"] = torch.ones(h, dtype=dtype)
    w["model.norm.weight"] = torch.ones(h, dtype=dtype)
    w["lm_head.weight"] = nrm(v, h)
    return w


def rope_cache(head_dim: int, max_length: int, base: float, max_position_embeddings: int,
               dtype=torch.float16) -> Tuple[torch.Tensor, torch.Tensor]:
    """Engine/Llama_modules.py:16-45  cos/sin built in fp32, sliced [:max_length], cast."""
    inv_freq = 1.0 / (base ** (torch.arange(0, head_dim, 2).float() / head_dim))
    t = torch.arange(max_position_embeddings, dtype=inv_freq.dtype)
    freqs = torch.outer(t, inv_freq)
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos()[:max_length].to(dtype), emb.sin()[:max_length].to(dtype)


def rotate_half(x: torch.Tensor) -> torch.Tensor:
    x1 = x[..., : x.shape[-1] // 2]
    x2 = x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)


def apply_rotary_pos_emb(q, k, cos, sin, position_ids, unsqueeze_dim: int = 1):
    """transformers 4.36.2 semantics (copy kept by the reference at Engine/offload_engine.py:35-67)."""
    cos = cos[position_ids].unsqueeze(unsqueeze_dim)
    sin = sin[position_ids].unsqueeze(unsqueeze_dim)
    q_embed = (q * cos) + (rotate_half(q) * sin)
    k_embed = (k * cos) + (rotate_half(k) * sin)
    return q_embed, k_embed


def repeat_kv(hidden_states: torch.Tensor, n_rep: int) -> torch.Tensor:
    b, h, s, d = hidden_states.shape
    if n_rep == 1:
        return hidden_states
    hidden_states = hidden_states[:, :, None, :, :].expand(b, h, n_rep, s, d)
    return hidden_states.reshape(b, h * n_rep, s, d)


def rmsnorm(x: torch.Tensor, weight: torch.Tensor, eps: float) -> torch.Tensor:
    """Engine/Llama_modules.py:274-288  fp32 variance, cast back, THEN multiply by weight."""
    dt = x.dtype
    xf = x.to(torch.float32)
    var = xf.pow(2).mean(-1, keepdim=True)
    xf = xf * torch.rsqrt(var + eps)
    return weight * xf.to(dt)


def attention_core_TG(q, k, v, mask) -> torch.Tensor:
    """Engine/Llama_modules.py:229-248  explicit matmul -> /sqrt(D) -> +mask -> fp32 softmax -> cast -> matmul.
    q (1,H,n,D); k,v (1,H,kv,D) already repeat_kv'd and sliced; mask (1,1,n,kv) or None."""
    d = q.shape[-1]
    w = torch.matmul(q, k.transpose(2, 3)) / math.sqrt(d)
    if mask is not None:
        w = w + mask
    w = softmax(w, dim=-1, dtype=torch.float32).to(q.dtype)
    return torch.matmul(w, v)


def attention_core_FI(q, k, v, mask) -> torch.Tensor:
    """Engine/Llama_modules.py:127-134  SDPA with an additive float mask over all M slots."""
    return F.scaled_dot_product_attention(q, k, v, attn_mask=mask, dropout_p=0.0, is_causal=False)


class LlamaOracle:
    """Engine/Llama_model.py:53-72,201-216 + Llama_modules.py:87-140 (FI) / 182-258 (TG).

    mode "FI" = draft engine (SDPA over the whole static cache),
    mode "TG" = target engine (explicit attention over kv_len slots)."""

    def __init__(self, cfg: LlamaCfg, weights: Dict[str, torch.Tensor], max_length: int, mode: str,
                 dtype=torch.float16, device="cpu"):
        assert mode in ("FI", "TG")
        self.cfg, self.w, self.max_length, self.mode, self.dtype = cfg, weights, max_length, mode, dtype
        self.kv_cache = KVCacheOracle(cfg.num_hidden_layers, cfg.num_key_value_heads, cfg.head_dim,
                                      max_length, dtype)
        self.cos, self.sin = rope_cache(cfg.head_dim, max_length, cfg.rope_theta,
                                        cfg.max_position_embeddings, dtype)
        if str(device) != "cpu":      # tests only: the same torch op sequence on the reference's own device (weights given there)
            self.kv_cache.k_cache = self.kv_cache.k_cache.to(device)
            self.kv_cache.v_cache = self.kv_cache.v_cache.to(device)
            self.cos, self.sin = self.cos.to(device), self.sin.to(device)

    @torch.no_grad()
    def forward(self, input_ids, storage_ids, position_ids, attention_mask) -> torch.Tensor:
        cfg, w = self.cfg, self.w
        H, Hkv, D = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim
        hs = F.embedding(input_ids, w["model.embed_tokens.weight"])
        bsz, q_len, _ = hs.shape
        for l in range(cfg.num_hidden_layers):
            p = f"model.layers.{l}."
            resid = hs
            x = rmsnorm(hs, w[p + "input_layernorm.weight"], cfg.rms_norm_eps)
            qs = F.linear(x, w[p + "self_attn.q_proj.weight"]).view(bsz, q_len, H, D).transpose(1, 2)
            ks = F.linear(x, w[p + "self_attn.k_proj.weight"]).view(bsz, q_len, Hkv, D).transpose(1, 2)
            vs = F.linear(x, w[p + "self_attn.v_proj.weight"]).view(bsz, q_len, Hkv, D).transpose(1, 2)
            qs, ks = apply_rotary_pos_emb(qs, ks, self.cos, self.sin, position_ids)
            ks, vs = self.kv_cache.update_kv_cache(ks, vs, l, storage_ids)
            if self.mode == "TG":
                kv_len = self.kv_cache.get_usable_length(l, len(storage_ids))
                ks = ks[..., :kv_len, :]
                vs = vs[..., :kv_len, :]
            ks = repeat_kv(ks, H // Hkv)
            vs = repeat_kv(vs, H // Hkv)
            if self.mode == "TG":
                if attention_mask is not None and attention_mask.size() != (bsz, 1, q_len, ks.shape[-2]):
                    raise ValueError(f"Attention mask should be of size {(bsz, 1, q_len, ks.shape[-2])}, "
                                     f"but is {attention_mask.size()}")
                a = attention_core_TG(qs, ks, vs, attention_mask)
            else:
                a = attention_core_FI(qs, ks, vs, attention_mask)
            a = a.transpose(1, 2).contiguous().reshape(bsz, q_len, H * D)
            a = F.linear(a, w[p + "self_attn.o_proj.weight"])
            hs = resid + a
            resid = hs
            x = rmsnorm(hs, w[p + "post_attention_layernorm.weight"], cfg.rms_norm_eps)
            g = F.silu(F.linear(x, w[p + "mlp.gate_proj.weight"]))
            u = F.linear(x, w[p + "mlp.up_proj.weight"])
            hs = resid + F.linear(g * u, w[p + "mlp.down_proj.weight"])
        hs = rmsnorm(hs, w["model.norm.weight"], cfg.rms_norm_eps)
        return F.linear(hs, w["lm_head.weight"])


class EngineOracle:
    """Engine/Engine.py GraphInferenceEngine / GraphInferenceEngineTG surface on CPU (eager only)."""

    def __init__(self, model: LlamaOracle):
        self.model = model
        self.max_length = model.max_length
        self.dtype = model.dtype
        self.kv_cache = model.kv_cache

    def inference(self, input_ids, storage_ids, position_ids, attn_mask):
        return self.model.forward(input_ids, storage_ids, position_ids, attn_mask)

    graph_inference = inference

    def clear_kv(self):
        self.kv_cache.clear()


# --------------------------------------------------------------------------------------
# Tree/SpecTree.py  and  Tree/GreedyTree.py
# --------------------------------------------------------------------------------------
@dataclass
class IterTrace:
    """What one construct_grow_map()+verify() produced (for parity tests)."""
    P: int
    tree_tokens: torch.Tensor            # (S-1,) tokens of nodes 1..S-1
    accept_list: List[int]               # absolute slots, incl. the P prefix slots
    bonus: Optional[int]
    terminal: bool
    valid_tokens: torch.Tensor


class SpecTreeOracle:
    """Tree/SpecTree.py:7-281 restated (stochastic Sequoia tree).  CPU, fp16."""

    def __init__(self, draft: EngineOracle, target: EngineOracle, prefix: torch.Tensor, grow_map: dict,
                 temperature: float = 0.6, top_p: float = 1.0, max_length: int = 256,
                 max_target_seq: Optional[int] = None, vocab_size: int = 32000,
                 bonus_noise: Optional[torch.Tensor] = None):
        dt = torch.float16
        self.dtype = dt
        self.draft, self.target = draft, target
        self.T, self.top_p = temperature, top_p
        self.M = max_length
        self.max_target_seq = max_length if max_target_seq is None else max_target_seq
        self.grow_map = grow_map
        self.S = grow_map["size"]
        self.Successors = grow_map["Successors"]
        self.draft_step = len(grow_map["roots"])
        self.roots = [torch.tensor(x, dtype=torch.long) for x in grow_map["roots"]]
        self.tokens = torch.zeros(max_length, dtype=torch.long)                     # Tree.py:5
        self.position_ids = torch.zeros(max_length, dtype=torch.long)
        P = len(prefix)
        self.tokens[:P] = prefix                                                    # Tree.py:21-27
        self.position_ids[:P] = torch.arange(P)
        self.num_nodes = P
        self.full = build_full_attn_mask(max_length, grow_map["mask"], dt)          # SpecTree.py:44-54
        self.attn_mask = window_mask(self.full, max_length, P + self.S - 1)         # :57-58
        self.ground_truth_len = P
        self.r = torch.rand(max_length, dtype=dt)                                   # :60  (CPU generator)
        self.depth = grow_map["depth"][1:]
        self.position_ids[P:P + self.S - 1] = self.depth + P - 1                    # :62
        self.storage_ids = torch.arange(max_length)
        self.draft_logits = torch.zeros((max_length, vocab_size), dtype=dt)         # :66
        out = self.draft.inference(self.tokens[:P].unsqueeze(0), self.storage_ids[:P],
                                   self.position_ids[:P].unsqueeze(0),
                                   self.attn_mask[:P][None, None, :, :])           # :68-72
        self.draft_logits[0] = out[..., -1, :][0]
        self.draft_kv_len = P
        self.target_kv_len = 0
        self.rand = torch.empty((self.S, vocab_size), dtype=dt).uniform_()          # :84 (CPU generator)
        self.gather_idx = [sample_gather_index(b) for b in grow_map["branches"][:-1]]
        # bonus_noise: optional (n_iter, V) Exp(1) draws; when given the bonus token is
        # argmax(residual / noise) (the n=1 form of torch.multinomial) so that a GPU run fed the
        # same noise is comparable; when None the oracle calls torch.multinomial like :222.
        self.bonus_noise = bonus_noise
        self.iter = 0

    # Tree/SpecTree.py:88-134
    def collective_grow_static(self, idx_list: torch.Tensor, n_branch_list: List[int], grow_step: int):
        total_branch = sum(n_branch_list)
        k = max(n_branch_list)
        new_tokens_set = sampling_without_replacement(self.draft_logits[idx_list], self.rand[idx_list], k, self.T)
        self.tokens[self.num_nodes:self.num_nodes + total_branch] = new_tokens_set[self.gather_idx[grow_step]]
        self.num_nodes += total_branch
        start_pos, end_pos = self.num_nodes - total_branch, self.num_nodes
        attn_mask = self.attn_mask[start_pos:end_pos][None, None, :, :]
        out = self.draft.graph_inference(self.tokens[self.draft_kv_len:self.num_nodes].unsqueeze(0),
                                         self.storage_ids[self.draft_kv_len:self.num_nodes],
                                         self.position_ids[start_pos:end_pos].unsqueeze(0), attn_mask)
        self.draft_kv_len = self.num_nodes
        g = self.ground_truth_len
        self.draft_logits[start_pos - g + 1:end_pos - g + 1] = out[0][-total_branch:]

    def construct_grow_map(self):                                                   # :245-259
        for i in range(self.draft_step - 1):
            self.collective_grow_static(self.roots[i], self.grow_map["branches"][i], i)

    def accept_step(self, parent_id: int):                                          # :137-157
        g = self.ground_truth_len
        logits_id = parent_id - (g - 1)
        p = self.target_logits[logits_id]
        draft_logits = self.draft_logits[logits_id]
        children = self.Successors[logits_id]
        if len(children) == 0:
            return -1, p
        for pos in children:
            token = self.tokens[pos + (g - 1)]
            q = softmax(draft_logits / self.T, dim=-1)
            r = self.r[pos + (g - 1)]
            if p[token] > r * q[token]:
                return pos + (g - 1), None
            p = get_residual(p, q)
            draft_logits[token] = torch.finfo(self.dtype).min
        return -1, p

    def verify(self):                                                               # :160-242
        g = self.ground_truth_len
        new_node_num = self.num_nodes - g + 1
        start_pos = 0 if self.target_kv_len == 0 else self.target_kv_len
        end_pos = self.num_nodes
        attn_mask = self.attn_mask[start_pos:end_pos, :end_pos][None, None, :, :].type(self.target.dtype)
        out = self.target.inference(self.tokens[start_pos:end_pos].unsqueeze(0), self.storage_ids[start_pos:end_pos],
                                    self.position_ids[start_pos:end_pos].unsqueeze(0), attn_mask)
        self.target_logits = out[0][g - 1:] if self.target_kv_len == 0 else out[0][-new_node_num:]
        assert len(self.target_logits) == new_node_num
        self.raw_target_logits = self.target_logits
        self.target_logits = get_sampling_logits(self.target_logits, self.top_p, self.T, replicate=False)
        self.target_logits = softmax(self.target_logits / self.T, dim=-1)
        tree_tokens = self.tokens[g:g + self.S - 1].clone()
        accept_list = list(range(g))
        terminal = False
        residual = None
        while True:
            pos, res = self.accept_step(accept_list[-1])
            if pos != -1:
                accept_list.append(pos)
                if self.tokens[pos] == 0 or self.tokens[pos] == 2:
                    terminal = True
                    break
            else:
                residual = res
                break
        a = len(accept_list)
        bonus = None
        if not terminal:
            if torch.isnan(residual).any():
                terminal = True
            else:
                if self.bonus_noise is None:
                    bonus = int(residual.multinomial(num_samples=1, replacement=True))
                else:
                    bonus = int(torch.argmax(residual / self.bonus_noise[self.iter]))
                self.tokens[a] = bonus
        self.last_residual = residual
        self.tokens[:a] = self.tokens[accept_list]
        self.draft.kv_cache.gather_kv_incremental(accept_list[g:], g)
        self.target.kv_cache.gather_kv_incremental(accept_list[g:], g)
        self.iter += 1
        if not terminal:
            valid = self.tokens[:a + 1]
            self.prepare_for_next_iter(accept_list, valid)
        else:
            valid = self.tokens[:a]
        self.last_trace = IterTrace(g, tree_tokens, list(accept_list), bonus, terminal, valid.clone())
        return valid, a, a, terminal

    def prepare_for_next_iter(self, accept_list: List[int], valid_tokens: torch.Tensor):   # :261-281
        a = len(accept_list)
        if a + 1 > self.max_target_seq:
            return
        self.position_ids[:a] = self.position_ids[accept_list]
        self.position_ids[a] = a
        n = len(valid_tokens)
        self.position_ids[n:n + self.S - 1] = self.depth + n - 1
        self.ground_truth_len = n
        self.num_nodes = n
        self.attn_mask = window_mask(self.full, self.M, n + self.S - 1)
        out = self.draft.graph_inference(self.tokens[a:self.num_nodes].unsqueeze(0), self.storage_ids[a:self.num_nodes],
                                         self.position_ids[a:self.num_nodes].unsqueeze(0),
                                         self.attn_mask[a:self.num_nodes][None, None, :, :])
        self.draft_logits[0] = out[..., -1, :][0]
        self.draft_kv_len = self.num_nodes
        self.target_kv_len = a


class GreedyTreeOracle(SpecTreeOracle):
    """Tree/GreedyTree.py:6-264 restated: top-k drafting, argmax verification."""

    def __init__(self, draft, target, prefix, grow_map, max_length=256, max_target_seq=None, vocab_size=32000):
        # GreedyTree.__init__ draws no random numbers (GreedyTree.py:59-83); build the shared
        # state without disturbing the caller's RNG stream.
        state = torch.get_rng_state()
        super().__init__(draft, target, prefix, grow_map, temperature=1.0, top_p=1.0, max_length=max_length,
                         max_target_seq=max_target_seq, vocab_size=vocab_size)
        torch.set_rng_state(state)
        self.r = None
        self.rand = None

    def collective_grow_static(self, idx_list, n_branch_list, grow_step):          # GreedyTree.py:86-130
        total_branch = sum(n_branch_list)
        k = max(n_branch_list)
        new_tokens_set = sampling_argmax(self.draft_logits[idx_list], k)
        self.tokens[self.num_nodes:self.num_nodes + total_branch] = new_tokens_set[self.gather_idx[grow_step]]
        self.num_nodes += total_branch
        start_pos, end_pos = self.num_nodes - total_branch, self.num_nodes
        attn_mask = self.attn_mask[start_pos:end_pos][None, None, :, :]
        out = self.draft.graph_inference(self.tokens[self.draft_kv_len:self.num_nodes].unsqueeze(0),
                                         self.storage_ids[self.draft_kv_len:self.num_nodes],
                                         self.position_ids[start_pos:end_pos].unsqueeze(0), attn_mask)
        self.draft_kv_len = self.num_nodes
        g = self.ground_truth_len
        self.draft_logits[start_pos - g + 1:end_pos - g + 1] = out[0][-total_branch:]

    def _target_tokens(self) -> torch.Tensor:                                       # GreedyTree.py:186
        return self.target_logits.argmax(dim=-1)

    def accept_step(self, parent_id: int) -> int:                                   # GreedyTree.py:132-146
        g = self.ground_truth_len
        logits_id = parent_id - (g - 1)
        target_token = self.target_token[logits_id]
        for pos in self.Successors[logits_id]:
            if self.tokens[pos + (g - 1)] == target_token:
                return pos + (g - 1)
        return -1

    def verify(self):                                                               # GreedyTree.py:151-223
        g = self.ground_truth_len
        new_node_num = self.num_nodes - g + 1
        start_pos = 0 if self.target_kv_len == 0 else self.target_kv_len
        end_pos = self.num_nodes
        attn_mask = self.attn_mask[start_pos:end_pos, :end_pos][None, None, :, :]
        out = self.target.inference(self.tokens[start_pos:end_pos].unsqueeze(0), self.storage_ids[start_pos:end_pos],
                                    self.position_ids[start_pos:end_pos].unsqueeze(0), attn_mask)
        self.target_logits = out[0][g - 1:] if self.target_kv_len == 0 else out[0][-new_node_num:]
        self.raw_target_logits = self.target_logits
        self.target_token = self._target_tokens()
        tree_tokens = self.tokens[g:g + self.S - 1].clone()
        accept_list = list(range(g))
        terminal = False
        while True:
            pos = self.accept_step(accept_list[-1])
            if pos != -1:
                accept_list.append(pos)
                if self.tokens[pos] == 0 or self.tokens[pos] == 2:
                    terminal = True
                    break
            else:
                break
        a = len(accept_list)
        self.tokens[:a] = self.tokens[accept_list]
        bonus = None
        self.iter += 1
        if not terminal:
            bonus = int(self.target_token[accept_list[-1] - g + 1])
            self.tokens[a] = bonus
            self.draft.kv_cache.gather_kv_incremental(accept_list[g:], g)
            self.target.kv_cache.gather_kv_incremental(accept_list[g:], g)
            valid = self.tokens[:a + 1]
            self.prepare_for_next_iter(accept_list, valid)
        else:
            valid = self.tokens[:a]
        self.last_trace = IterTrace(g, tree_tokens, list(accept_list), bonus, terminal, valid.clone())
        return valid, a, a, terminal


class GreedySTreeOracle(GreedyTreeOracle):
    """Tree/GreedySTree.py restated: GreedyTree drafting (top-k children) verified against a target token SAMPLED
    from softmax(top_p(target_logits) / T) instead of the argmax (GreedySTree.py:188-190).

    target_words: optional (n_iter, S, V) fp16 uniforms; when given, row k's token is the k=1 exponential race
    argmax(log(u) / p) (`sampling_without_replacement`, the same draw torch.multinomial(1) makes from its own
    noise) so that a GPU run fed the same uniforms is comparable; when None the oracle calls torch.multinomial
    exactly like the reference."""

    def __init__(self, draft, target, prefix, grow_map, temperature=0.6, top_p=1.0, max_length=256,
                 max_target_seq=None, vocab_size=32000, target_uniforms: Optional[torch.Tensor] = None):
        super().__init__(draft, target, prefix, grow_map, max_length=max_length, max_target_seq=max_target_seq,
                         vocab_size=vocab_size)
        self.T, self.top_p = temperature, top_p
        self.target_uniforms = target_uniforms

    def _target_tokens(self) -> torch.Tensor:                                       # GreedySTree.py:188-190
        lg = get_sampling_logits(self.target_logits, self.top_p, self.T, replicate=False)
        if self.target_uniforms is not None:
            return sampling_without_replacement(lg, self.target_uniforms[self.iter], 1, self.T)
        self.target_logits = softmax(lg / self.T, dim=-1)
        return self.target_logits.multinomial(num_samples=1).flatten()


def multinomial_words(q: torch.Tensor, words: torch.Tensor) -> torch.Tensor:
    """Sampling WITH replacement from fp16 probability rows by exact integer inverse-CDF.

    Every fp16 value in [0, 1] is an integer multiple of 2^-24, so w = q * 2^24 is an exact integer weight, prefix sums
    are exact and order-independent, and draw j of row i is the first index whose inclusive prefix sum exceeds
    (words[i, j] * total_i) >> 32 with words uniform in [0, 2^32).  The CUDA kernel (sq_sample_replace) does the same
    arithmetic, so given identical q rows the two agree bit for bit.  q: (rows, V) fp16; words: (rows, k) int64."""
    w = (q.double() * 16777216.0).round().long()
    cdf = w.cumsum(-1)
    total = cdf[:, -1:]
    t = (words.long() * total) >> 32
    return torch.searchsorted(cdf, t, right=True).clamp_(max=q.shape[-1] - 1)


class SpecInferTreeOracle(SpecTreeOracle):
    """Tree/SpecInferTree.py restated (the SpecInfer baseline policy behind the same tree machinery): children are
    drawn i.i.d. WITH replacement from softmax(draft/T) (:100-105), the walk accepts on `>=` and never masks a
    rejected token out of q (:143-162); everything else is SpecTree.

    sample_words: optional (n_iter, S) int64 words in [0, 2^32): node k's token is drawn with words[iter, k] through
    `multinomial_words`; forced_tokens: optional callable(iter) -> (S-1,) tree tokens to install instead of sampling
    (lock-step tests feed the GPU's tree).  With neither, torch.multinomial is called exactly like the reference."""

    def __init__(self, *a, sample_words: Optional[torch.Tensor] = None, forced_tokens=None, **kw):
        super().__init__(*a, **kw)
        self.sample_words = sample_words
        self.forced_tokens = forced_tokens

    def collective_grow_static(self, idx_list, n_branch_list, grow_step):            # SpecInferTree.py:88-134
        total_branch = sum(n_branch_list)
        k = max(n_branch_list)
        q = softmax(self.draft_logits[idx_list] / self.T, dim=-1)
        g = self.ground_truth_len
        first_child = self.num_nodes - g + 1                                         # node id of this level's first child
        if self.forced_tokens is not None:
            new = self.forced_tokens(self.iter)[first_child - 1:first_child - 1 + total_branch]
            self.tokens[self.num_nodes:self.num_nodes + total_branch] = new
        elif self.sample_words is not None:
            words = torch.zeros((len(n_branch_list), k), dtype=torch.long)
            c = first_child
            for j, nb in enumerate(n_branch_list):
                words[j, :nb] = self.sample_words[self.iter, c:c + nb]
                c += nb
            new_tokens_set = multinomial_words(q, words).flatten()
            self.tokens[self.num_nodes:self.num_nodes + total_branch] = new_tokens_set[self.gather_idx[grow_step]]
        else:
            new_tokens_set = q.multinomial(num_samples=k, replacement=True).flatten()
            self.tokens[self.num_nodes:self.num_nodes + total_branch] = new_tokens_set[self.gather_idx[grow_step]]
        self.num_nodes += total_branch
        start_pos, end_pos = self.num_nodes - total_branch, self.num_nodes
        attn_mask = self.attn_mask[start_pos:end_pos][None, None, :, :]
        out = self.draft.graph_inference(self.tokens[self.draft_kv_len:self.num_nodes].unsqueeze(0),
                                         self.storage_ids[self.draft_kv_len:self.num_nodes],
                                         self.position_ids[start_pos:end_pos].unsqueeze(0), attn_mask)
        self.draft_kv_len = self.num_nodes
        self.draft_logits[start_pos - g + 1:end_pos - g + 1] = out[0][-total_branch:]

    def accept_step(self, parent_id: int):                                           # SpecInferTree.py:143-162
        g = self.ground_truth_len
        logits_id = parent_id - (g - 1)
        p = self.target_logits[logits_id]
        draft_logits = self.draft_logits[logits_id]
        children = self.Successors[logits_id]
        if len(children) == 0:
            return -1, p
        for pos in children:
            token = self.tokens[pos + (g - 1)]
            q = softmax(draft_logits / self.T, dim=-1)
            r = self.r[pos + (g - 1)]
            if p[token] >= r * q[token]:
                return pos + (g - 1), None
            p = get_residual(p, q)
        return -1, p


# --------------------------------------------------------------------------------------
# tests/testbed.py simulation_fast (the metric loop), restated for the oracle engines
# --------------------------------------------------------------------------------------
def simulation_fast(draft: EngineOracle, target: EngineOracle, prompts: Sequence[torch.Tensor], grow_map: dict,
                    T: float = 0.6, top_p: float = 1.0, max_length: int = 384, greedy: bool = False,
                    max_new_len: int = 256, max_iters: Optional[int] = None, vocab_size: int = 32000):
    """tests/testbed.py:45-95.  Returns (num_decoding_steps, num_large_model_steps, traces)."""
    num_decoding_steps = 0
    num_large_model_steps = 0
    traces: List[List[IterTrace]] = []
    for prompt in prompts:
        input_ids = prompt.view(1, -1)
        if greedy:
            tree = GreedyTreeOracle(draft, target, input_ids[0], grow_map, max_length=max_length,
                                    max_target_seq=max_length, vocab_size=vocab_size)
        else:
            tree = SpecTreeOracle(draft, target, input_ids[0], grow_map, temperature=T, top_p=top_p,
                                  max_length=max_length, max_target_seq=max_length, vocab_size=vocab_size)
        tr: List[IterTrace] = []
        terminate = False
        it = 0
        while input_ids.shape[1] < max_new_len and not terminate:
            tree.construct_grow_map()
            valid_tokens, _, _, terminate = tree.verify()
            tr.append(tree.last_trace)
            num_decoding_steps += valid_tokens.shape[0] - input_ids.shape[1]
            num_large_model_steps += 1
            input_ids = valid_tokens.unsqueeze(0)
            if input_ids[0][-1] == 2 or input_ids[0][-1] == 0:
                terminate = True
            it += 1
            if max_iters is not None and it >= max_iters:
                break
        traces.append(tr)
        draft.clear_kv()
        target.clear_kv()
    return num_decoding_steps, num_large_model_steps, traces


def top_p_filter_integer(logits: torch.Tensor, top_p: float, T: float) -> torch.Tensor:
    """The algorithm of the `sq_top_p_filter` kernel (csrc/sq_sampling.cu) restated on the CPU -- TEST INFRASTRUCTURE: it lets
    the CPU suite check the kernel's arithmetic (exact integer masses instead of a sorted fp16 cumsum) against the reference's
    golden vector (utils.py:65-77).  fp16 probabilities are integer multiples of 2^-24, so the mass ranked before a token is
    an exact integer S; the token is removed iff fp16(S * 2^-24) > fp16(top_p).  Ranking: value descending, index ascending."""
    out = logits.clone()
    tp = float(torch.tensor(top_p, dtype=torch.float16))
    for r in range(logits.shape[0]):
        xt = (logits[r].float() * (1.0 / T)).to(torch.float16)                  # what torch's CUDA div-by-scalar computes
        p = softmax(xt.float(), dim=-1).to(torch.float16)
        w = (p.double() * 16777216.0).round().to(torch.int64)                   # exact
        order = sorted(range(xt.numel()), key=lambda i: (-float(xt[i]), i))
        before = 0
        for i in order:
            if float(torch.tensor(before / 16777216.0, dtype=torch.float32).to(torch.float16)) > tp:
                out[r, i] = float("-inf")
            before += int(w[i])
    return out
