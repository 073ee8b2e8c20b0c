"""utils.py of the reference on sequoia_b200 kernels: same names, arguments and return values.

The `cuda_graph_for_*` factories of the reference wrap 5-10 tiny torch kernels in a CUDA graph with static input
copies (utils.py:109-211); here each of them is ONE fused kernel, so the returned closure simply launches it
(it stays capturable inside a caller's graph)."""
from __future__ import annotations

import dataclasses

import torch

from . import ops

F16 = torch.float16


def _rows(t: torch.Tensor) -> torch.Tensor:
    t = t if t.dim() == 2 else t.reshape(-1, t.shape[-1])
    return t if t.stride(-1) == 1 else t.contiguous()


def get_residual(p: torch.Tensor, q: torch.Tensor) -> torch.Tensor:
    """utils.py:5-8."""
    if p.dim() != 1:
        rows = [ops.residual(a.contiguous(), b.contiguous()) for a, b in zip(_rows(p), _rows(q))]
        return torch.stack(rows).view(p.shape)
    return ops.residual(p.contiguous(), q.contiguous())


def sampling_without_replacement(sampling_logits: torch.Tensor, rand: torch.Tensor, num_samples: int,
                                 temperature: float) -> torch.Tensor:
    """utils.py:10-18 -> flattened (rows*num_samples,) int64 positions."""
    lg, rd = _rows(sampling_logits), _rows(rand)
    pos = torch.empty(lg.shape[0] * num_samples, dtype=torch.int64, device=lg.device)
    ops.sample_level(lg, rd, lg.shape[0], num_samples, float(temperature), 0, positions=pos)
    return pos


def sampling_argmax(sampling_logits: torch.Tensor, num_samples: int) -> torch.Tensor:
    """utils.py:29-32."""
    lg = _rows(sampling_logits)
    pos = torch.empty(lg.shape[0] * num_samples, dtype=torch.int64, device=lg.device)
    ops.sample_level(lg, None, lg.shape[0], num_samples, 1.0, 1, positions=pos)
    return pos


def sampling_with_replacement(sampling_logits: torch.Tensor, num_samples: int, temperature: float) -> torch.Tensor:
    """utils.py:20-28 (SpecInfer baseline; torch.multinomial as in the reference)."""
    q = ops.softmax_T(_rows(sampling_logits), float(temperature))
    return q.multinomial(num_samples=num_samples, replacement=False).flatten()


def get_sampling_logits(logits: torch.Tensor, top_p: float, T: float, replicate=False):
    """utils.py:65-77 (identity when top_p >= 1.0, which every named configuration uses)."""
    if replicate:
        logits = logits.clone()
    if top_p < 1.0:
        shape = logits.shape
        ops.top_p_filter_(logits.view(-1, shape[-1]), float(top_p), float(T))     # sq_top_p_filter kernel, in place
    return logits


def make_tree_attention_mask(prefix_len: int, gen_len: int, ancestors, device="cpu", dtype=torch.float32):
    """utils.py:52-62."""
    tree_mask = torch.full((gen_len, gen_len + prefix_len), torch.finfo(dtype).min, dtype=dtype).to(device=device)
    for idx, ancestor in enumerate(ancestors):
        if len(ancestor) > 0:
            tree_mask[idx][ancestor] = 0.0
    return tree_mask[None, None, :, :]


@dataclasses.dataclass
class ChildrenAccept:
    accept_mark: int = None
    token: int = None
    position: int = None
    successor_order: int = -1
    residual: torch.FloatTensor = None


def _make_causal_mask(input_ids_shape, dtype: torch.dtype, device):
    """utils.py:95-107."""
    _, tgt_len = input_ids_shape
    mask = torch.full((tgt_len, tgt_len), torch.finfo(dtype).min, device=device)
    mask_cond = torch.arange(mask.size(-1), device=device)
    mask.masked_fill_(mask_cond < (mask_cond + 1).view(mask.size(-1), 1), 0)
    return mask.to(dtype)


def cuda_graph_for_residual(device="cuda:0", dtype=torch.float16, dim=32000, n_warmups=3, mempool=None):
    """utils.py:109-136 -> run(p, q)."""
    def run(p, q):
        return get_residual(p, q)
    return run


def cuda_graph_for_sampling_without_replacement(device="cuda:0", dtype=torch.float16, dim=32000, max_length=384,
                                                n_warmups=3, mempool=None, idx_len=8, num_samples=16, temperature=0.6,
                                                tree_size=64):
    """utils.py:138-177 -> run(draft_logits, rand_vector)."""
    def run(draft_logits, rand_vector):
        return sampling_without_replacement(draft_logits, rand_vector, num_samples, temperature)
    return run


def cuda_graph_for_sampling_argmax(device="cuda:0", dtype=torch.float16, dim=32000, max_length=384, n_warmups=3,
                                   mempool=None, idx_len=8, num_samples=16, temperature=0.6, tree_size=64):
    """utils.py:179-211 -> run(draft_logits)."""
    def run(draft_logits):
        return sampling_argmax(draft_logits, num_samples)
    return run


def cuda_graph_for_sampling_with_replacement(device="cuda:0", dtype=torch.float16, dim=32000, max_length=384,
                                             n_warmups=3, mempool=None, idx_len=8, num_samples=16, temperature=0.6,
                                             tree_size=64):
    """utils.py:214-246 -> run(draft_logits)."""
    def run(draft_logits):
        return sampling_with_replacement(draft_logits, num_samples, temperature)
    return run
