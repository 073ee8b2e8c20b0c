"""Tensor-level wrappers over the C ABI (one function per exported kernel).

Everything here launches on torch's current CUDA stream, never allocates on the hot path unless
an output tensor is not supplied, and raises on any error (no fallback).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import check, ptr, stream_ptr

F16 = torch.float16


def _need(t: torch.Tensor, dtype, name: str):
    if t.dtype != dtype or not t.is_cuda:
        raise TypeError(f"{name}: expected CUDA tensor of {dtype}, got {t.dtype} on {t.device}")


def embed_rows(table, tokens, n, out, state=None, n0=0):
    lib = _lib.load()
    check(lib.sq_embed_rows(ptr(table), ptr(tokens), ptr(state), n0, n, table.shape[1], ptr(out), stream_ptr()),
          "sq_embed_rows")


def rmsnorm(x, weight, out, n, eps):
    lib = _lib.load()
    check(lib.sq_rmsnorm(ptr(x), ptr(weight), ptr(out), n, x.shape[-1], eps, stream_ptr()), "sq_rmsnorm")


def add_rmsnorm(resid, delta, weight, out, n, eps):
    lib = _lib.load()
    check(lib.sq_add_rmsnorm(ptr(resid), ptr(delta), ptr(weight), ptr(out), n, resid.shape[-1], eps, stream_ptr()),
          "sq_add_rmsnorm")


def silu_mul(gate_up, out, n, interleaved=False):
    """out = silu(gate) * up; interleaved: gate_up columns in blocks of 32 = 16 gate | 16 up (interleave_gate_up order)."""
    lib = _lib.load()
    check(lib.sq_silu_mul_ex(ptr(gate_up), ptr(out), n, out.shape[-1], 1 if interleaved else 0, stream_ptr()), "sq_silu_mul_ex")


def rope_kv_append(qkv, H, Hkv, D, cos, sin, position_ids, storage_ids, n, k_layer, v_layer, M, state=None, n0=0):
    lib = _lib.load()
    check(lib.sq_rope_kv_append(ptr(qkv), qkv.shape[-1], H, Hkv, D, ptr(cos), ptr(sin), ptr(position_ids),
                                ptr(storage_ids), ptr(state), n0, n, ptr(k_layer), ptr(v_layer), M, stream_ptr()),
          "sq_rope_kv_append")


def kv_gather(k_cache, v_cache, idx, n, offset, state=None, max_n=0, zero_tail=False):
    """k_cache/v_cache (L,1,Hkv,M,D); idx int32 device tensor."""
    lib = _lib.load()
    L, _, Hkv, M, D = k_cache.shape
    check(lib.sq_kv_gather(ptr(k_cache), ptr(v_cache), L, Hkv, M, D, ptr(idx), n, offset, ptr(state), max_n,
                           1 if zero_tail else 0, stream_ptr()), "sq_kv_gather")


def kv_gather_big(k_cache, v_cache, idx, n, offset, zero_tail=False):
    """Index lists too long for the on-chip staging of `kv_gather` (host-known n / offset): through a global scratch."""
    lib = _lib.load()
    L, _, Hkv, M, D = k_cache.shape
    nbytes = lib.sq_kv_gather_scratch_bytes(L, Hkv, D, n)
    scratch = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=k_cache.device)
    check(lib.sq_kv_gather_big(ptr(k_cache), ptr(v_cache), L, Hkv, M, D, ptr(idx), n, offset, ptr(scratch), nbytes,
                               1 if zero_tail else 0, stream_ptr()), "sq_kv_gather_big")


class AttnPlan:
    """TMA descriptors + split-KV workspace for one (qkv buffer, KV cache, output buffer) triple."""

    def __init__(self, qkv: torch.Tensor, n_max: int, H: int, Hkv: int, D: int, k_cache, v_cache, out):
        lib = _lib.load()
        L, _, hk, M, d = k_cache.shape
        assert hk == Hkv and d == D
        nbytes = lib.sq_attn_workspace_bytes(n_max, H, D, M)
        self.workspace = torch.zeros(nbytes, dtype=torch.uint8, device=qkv.device)
        self.handle = C.c_void_p()
        self._keep = (qkv, k_cache, v_cache, out)
        check(lib.sq_attn_plan_create(C.byref(self.handle), ptr(qkv), qkv.shape[-1], n_max, H, Hkv, D, ptr(k_cache),
                                      ptr(v_cache), L, M, ptr(out), ptr(self.workspace), nbytes), "sq_attn_plan_create")
        self.n_max, self.M = n_max, M

    def error(self) -> int:
        return _lib.load().sq_attn_plan_error(self.handle)

    def __del__(self):
        try:
            if self.handle:
                _lib.load().sq_attn_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


def tree_attn(plan: AttnPlan, layer, n, *, state=None, n0=0, kv_end=0, prefix_len=0, dense_mask=None, mask_ld=0,
              tree_bits=None, tree_words=0, tree_size=0, impl=0):
    lib = _lib.load()
    check(lib.sq_tree_attn(plan.handle, layer, n, ptr(state), n0, kv_end, prefix_len, ptr(dense_mask), mask_ld,
                           ptr(tree_bits), tree_words, tree_size, impl, stream_ptr()), "sq_tree_attn")


def softmax_T(logits: torch.Tensor, T: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _need(logits, F16, "softmax_T")
    lg = logits.reshape(-1, logits.shape[-1])
    assert lg.stride(-1) == 1
    if out is None:
        out = torch.empty((lg.shape[0], lg.shape[1]), dtype=F16, device=logits.device)
    lib = _lib.load()
    check(lib.sq_softmax_T(ptr(lg), lg.stride(0), ptr(out), out.stride(0), lg.shape[0], lg.shape[1], T, stream_ptr()),
          "sq_softmax_T")
    return out.view(logits.shape) if out.numel() == logits.numel() else out


def sample_level(logits, rand, n_parents, k_max, T, mode, *, parent_rows=None, child_first=None, n_branch=None,
                 positions=None, tokens=None, state=None):
    lib = _lib.load()
    V = logits.shape[-1]
    check(lib.sq_sample_level(ptr(logits), logits.stride(-2), ptr(rand), rand.stride(-2) if rand is not None else 0,
                              ptr(parent_rows), ptr(child_first), ptr(n_branch), n_parents, k_max, V, T, mode,
                              ptr(positions), ptr(tokens), ptr(state), stream_ptr()), "sq_sample_level")


def sample_replace(logits, words, n_parents, k_max, T, *, parent_rows=None, child_first=None, n_branch=None,
                   positions=None, tokens=None, state=None):
    """i.i.d. draws with replacement from softmax(logits/T) rows (SpecInferTree.py:100-105); words: int64 in [0, 2^32)."""
    lib = _lib.load()
    if words.dtype != torch.int64:
        raise TypeError("sample_replace: words must be int64")
    check(lib.sq_sample_replace(ptr(logits), logits.stride(-2), ptr(words), ptr(parent_rows), ptr(child_first),
                                ptr(n_branch), n_parents, k_max, logits.shape[-1], T, ptr(positions), ptr(tokens),
                                ptr(state), stream_ptr()), "sq_sample_replace")


def residual(p, q, out=None):
    _need(p, F16, "residual")
    if out is None:
        out = torch.empty_like(p)
    check(_lib.load().sq_residual(ptr(p), ptr(q), ptr(out), p.shape[-1], stream_ptr()), "sq_residual")
    return out


def top_p_filter_(logits, top_p: float, T: float):
    """get_sampling_logits (utils.py:65-77), in place on (n, V) fp16 logits."""
    _need(logits, F16, "top_p_filter_")
    n, V = logits.shape
    check(_lib.load().sq_top_p_filter(ptr(logits), logits.stride(0), n, V, float(top_p), float(T), stream_ptr()),
          "sq_top_p_filter")
    return logits


def argmax_rows(logits, out=None):
    n, V = logits.shape
    if out is None:
        out = torch.empty(n, dtype=torch.int64, device=logits.device)
    check(_lib.load().sq_argmax_rows(ptr(logits), logits.stride(0), n, V, ptr(out), stream_ptr()), "sq_argmax_rows")
    return out


ACCEPT_GE, ACCEPT_KEEP_Q = 1, 2          # sq_accept_stochastic policy bits (include/sequoia_b200.h)


def accept_stochastic(target_logits, draft_logits, r, noise, succ_off, succ, depth, S, T, tokens, position_ids,
                      accept_idx, state, max_target_seq, policy=0):
    V = target_logits.shape[-1]
    check(_lib.load().sq_accept_stochastic(ptr(target_logits), target_logits.stride(0), ptr(draft_logits),
                                           draft_logits.stride(0), ptr(r), ptr(noise), ptr(succ_off), ptr(succ),
                                           ptr(depth), S, V, T, ptr(tokens), ptr(position_ids), ptr(accept_idx),
                                           ptr(state), max_target_seq, policy, stream_ptr()), "sq_accept_stochastic")


def accept_greedy(target_token, succ_off, succ, depth, S, tokens, position_ids, accept_idx, state, max_target_seq):
    check(_lib.load().sq_accept_greedy(ptr(target_token), ptr(succ_off), ptr(succ), ptr(depth), S, ptr(tokens),
                                       ptr(position_ids), ptr(accept_idx), ptr(state), max_target_seq, stream_ptr()),
          "sq_accept_greedy")


def l2_prefetch(w: torch.Tensor, col_from: int, col_to: int):
    """Hint: pull columns [col_from, col_to) of every row of the 2-D row-major tensor `w` into L2 (current stream)."""
    es = w.element_size()
    col_to = min(col_to, w.shape[1])
    if col_to <= col_from:
        return
    check(_lib.load().sq_l2_prefetch(ptr(w), w.stride(0) * es, w.shape[0], col_from * es, (col_to - col_from) * es,
                                     stream_ptr()), "sq_l2_prefetch")


GEMM_TILED, GEMM_SWIGLU = 1, 2          # sq_gemm_plan_create_ex flags (include/sequoia_b200.h)


def gemm_pick_tiles(N: int, K: int, flags: int = 0):
    bn, sp, mc = C.c_int(), C.c_int(), C.c_int()
    check(_lib.load().sq_gemm_pick_tiles_ex(N, K, flags, C.byref(bn), C.byref(sp), C.byref(mc)), "sq_gemm_pick_tiles_ex")
    return bn.value, sp.value, mc.value


def tile_weights(w: torch.Tensor, flags: int = 0) -> torch.Tensor:
    """(N, K) row-major -> (ceil(N/BN), K/64, BN, 64) contiguous, rows beyond N zero (BN = the plan's tile width)."""
    N, K = w.shape
    bn, _, _ = gemm_pick_tiles(N, K, flags)
    tiles = (N + bn - 1) // bn
    if tiles * bn != N:
        pad = torch.zeros(tiles * bn, K, dtype=w.dtype, device=w.device)
        pad[:N] = w
        w = pad
    return w.view(tiles, bn, K // 64, 64).permute(0, 2, 1, 3).contiguous()


def interleave_gate_up(wg: torch.Tensor, wu: torch.Tensor) -> torch.Tensor:
    """(I, K) gate and up weights -> (2I, K) with rows 32b..32b+15 = gate[16b..], rows 32b+16..32b+31 = up[16b..]: the
    row order the fused SwiGLU epilogue of sq_gemm expects."""
    I, K = wg.shape
    assert I % 16 == 0 and wu.shape == wg.shape
    return torch.stack([wg.reshape(I // 16, 16, K), wu.reshape(I // 16, 16, K)], dim=1).reshape(2 * I, K).contiguous()


class GemmPlan:
    """C[:n] = A[:n] @ W.T on the weight-streaming tcgen05 kernel (csrc/sq_gemm.cu); n > 128 runs one launch per 128 rows."""

    def __init__(self, a: torch.Tensor, w: torch.Tensor, c: torch.Tensor, err_flag: Optional[torch.Tensor] = None,
                 tiled: bool = False, swiglu: bool = False):
        """tiled=True: `w` (N, K) is re-laid out once into the plan's own HBM-friendly copy (`self.w_tiled`, tile (n-tile,
        k-block) = one contiguous BN x 64 block) and the plan streams that copy.
        swiglu=True: `w` = interleave_gate_up(gate, up); the output (n, N/2) is silu(gate) * up (fused epilogue)."""
        lib = _lib.load()
        assert a.dtype == F16 and w.dtype == F16 and c.dtype == F16 and w.is_contiguous()
        assert a.stride(-1) == 1 and c.stride(-1) == 1 and a.shape[1] == w.shape[1]
        assert c.shape[1] >= (w.shape[0] // 2 if swiglu else w.shape[0])
        self.handle = C.c_void_p()
        self.w_tiled = None
        self.N, self.K, self.swiglu = w.shape[0], w.shape[1], swiglu
        flags = (GEMM_TILED if tiled else 0) | (GEMM_SWIGLU if swiglu else 0)
        if tiled:
            self.w_tiled = w = tile_weights(w, flags)
        self._keep = (a, w, c, err_flag)
        check(lib.sq_gemm_plan_create_ex(C.byref(self.handle), ptr(a), a.stride(0), a.shape[0], ptr(w), self.N, self.K,
                                         ptr(c), c.stride(0), ptr(err_flag), flags), "sq_gemm_plan_create_ex")

    def info(self):
        bn, sp, st = C.c_int(), C.c_int(), C.c_int()
        _lib.load().sq_gemm_plan_info(self.handle, C.byref(bn), C.byref(sp), C.byref(st))
        return bn.value, sp.value, st.value

    def run(self, n: int, a_row0: int = 0, out: Optional[torch.Tensor] = None):
        """rows [a_row0, a_row0+n) of the activation buffer -> `out[:n]` (default: the plan's output buffer, same rows)."""
        if a_row0 == 0 and out is None:
            check(_lib.load().sq_gemm_run(self.handle, n, stream_ptr()), "sq_gemm_run")
            return
        if out is not None:
            assert out.dtype == F16 and out.stride(-1) == 1 and out.shape[0] >= n
        check(_lib.load().sq_gemm_run_at(self.handle, n, a_row0, ptr(out), out.stride(0) if out is not None else 0,
                                         stream_ptr()), "sq_gemm_run_at")

    def __del__(self):
        try:
            if self.handle:
                _lib.load().sq_gemm_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


# ---- fused draft forward (csrc/sq_draft.cu) -------------------------------------------------------------------------------
def draft_supported(hidden, inter, n_layers, n_heads, n_kv_heads, head_dim, vocab, max_length) -> bool:
    return bool(_lib.load().sq_draft_supported(hidden, inter, n_layers, n_heads, n_kv_heads, head_dim, vocab, max_length))


class DraftPlan:
    """One persistent cooperative kernel per draft-tree level: embed -> L decoder layers -> lm_head for <= 64 rows."""
    MAX_ROWS = 64

    def __init__(self, hidden, inter, n_heads, vocab, max_length, eps, embed, layers, final_norm, lm_head, cos, sin,
                 k_cache, v_cache):
        lib = _lib.load()
        dev = embed.device
        ws_bytes = lib.sq_draft_workspace_bytes(hidden, inter)
        if ws_bytes <= 0:
            raise _lib.SequoiaLibError("sq_draft_workspace_bytes: unsupported shape")
        self.workspace = torch.zeros(ws_bytes, dtype=torch.uint8, device=dev)
        flat = []
        for ly in layers:
            flat += [ly["wqkv"], ly["wo"], ly["wgu"], ly["wd"], ly["ln1"], ly["ln2"]]
        for t in flat + [embed, final_norm, lm_head, cos, sin, k_cache, v_cache]:
            assert t.dtype == F16 and t.is_contiguous()
        self._keep = (flat, embed, final_norm, lm_head, cos, sin, k_cache, v_cache)
        arr = (C.c_void_p * len(flat))(*[t.data_ptr() for t in flat])
        self.handle = C.c_void_p()
        check(lib.sq_draft_plan_create(C.byref(self.handle), hidden, inter, len(layers), n_heads, vocab, max_length, float(eps),
                                       ptr(embed), arr, ptr(final_norm), ptr(lm_head), ptr(cos), ptr(sin), ptr(k_cache),
                                       ptr(v_cache), ptr(self.workspace), ws_bytes), "sq_draft_plan_create")

    def forward(self, n, tokens, position_ids, storage_ids, state, n0, kv_end, tree_bits, tree_words, tree_size, logits_out):
        assert logits_out.dtype == F16 and logits_out.stride(-1) == 1 and logits_out.shape[0] >= n
        check(_lib.load().sq_draft_forward(self.handle, n, ptr(tokens), ptr(position_ids), ptr(storage_ids), ptr(state), n0,
                                           kv_end, ptr(tree_bits), tree_words, tree_size, ptr(logits_out),
                                           logits_out.stride(0), stream_ptr()), "sq_draft_forward")

    def attention(self, layer, n, qkv, attn_out, state, n0, kv_end, tree_bits, tree_words, tree_size):
        check(_lib.load().sq_draft_attention(self.handle, layer, n, ptr(qkv), ptr(attn_out), ptr(state), n0, kv_end,
                                             ptr(tree_bits), tree_words, tree_size, stream_ptr()), "sq_draft_attention")

    def __del__(self):
        try:
            if self.handle:
                _lib.load().sq_draft_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass
