"""Drop-in for the reference's Tree/GreedyTree.py import path (tests/testbed_greedy.py, tests/test_accept.py)."""
from sequoia_b200.tree import GreedyTree, GreedyTreeTest  # noqa: F401
