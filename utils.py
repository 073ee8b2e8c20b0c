"""Drop-in for the reference's top-level utils.py import path (tests/testbed.py:16)."""
from sequoia_b200.sampling import (ChildrenAccept, _make_causal_mask, cuda_graph_for_residual,  # noqa: F401
                                   cuda_graph_for_sampling_argmax, cuda_graph_for_sampling_with_replacement,
                                   cuda_graph_for_sampling_without_replacement, get_residual, get_sampling_logits,
                                   make_tree_attention_mask, sampling_argmax, sampling_with_replacement,
                                   sampling_without_replacement)
