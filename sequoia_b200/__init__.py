"""sequoia_b200: B200-native implementation of Sequoia's tree-speculative-decoding hot path.

Python host over PyTorch tensors (device memory, streams, graphs, NCCL plumbing) calling hand-written sm_100a CUDA
through the C ABI in include/sequoia_b200.h (libsequoia_b200.so).  There is no CPU / eager fallback."""
from . import _lib  # noqa: F401
from .engine import (GraphInferenceEngine, GraphInferenceEngineTG, InferenceEngine, InferenceEngineTG,  # noqa: F401
                     OffloadEngine, capture_graph)
from .kv import KV_Cache  # noqa: F401
from .tree import GreedyTree, SpecTree, Tree  # noqa: F401

__version__ = "0.1.0"
