"""Drop-in for the reference's Engine/offload_engine.py import path (tests/testbed.py:18)."""
from sequoia_b200.engine import OffloadEngine  # noqa: F401
