"""Drop-in for the reference's Tree/GreedyTree.py import path (tests/testbed_greedy.py)."""
from sequoia_b200.tree import GreedyTree  # noqa: F401
