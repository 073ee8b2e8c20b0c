"""Drop-in for the reference's Tree/GreedySTree.py import path."""
from sequoia_b200.tree import GreedySTree  # noqa: F401
