"""Drop-in for the reference's Tree/SpecTree.py import path (tests/testbed.py:14)."""
from sequoia_b200.tree import SpecTree  # noqa: F401
