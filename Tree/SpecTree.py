"""Drop-in for the reference's Tree/SpecTree.py import path (tests/testbed.py, tests/test_accept.py)."""
from sequoia_b200.tree import SpecTree, SpecTreeTest  # noqa: F401
