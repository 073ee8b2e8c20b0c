"""Drop-in for the reference's Tree/SpecInferTree.py import path (tests/testbed.py --Mode baseline-style drivers)."""
from sequoia_b200.tree import SpecInferTree  # noqa: F401
