"""Drop-in for the reference's Engine/Engine.py import path (tests/testbed.py:17)."""
from sequoia_b200.engine import (GraphInferenceEngine, GraphInferenceEngineTG, InferenceEngine,  # noqa: F401
                                 InferenceEngineTG, capture_graph)
