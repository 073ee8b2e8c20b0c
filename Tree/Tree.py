"""Drop-in for the reference's Tree/Tree.py import path."""
from sequoia_b200.tree import Tree  # noqa: F401
