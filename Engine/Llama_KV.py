"""Drop-in for the reference's Engine/Llama_KV.py import path."""
from sequoia_b200.kv import KV_Cache  # noqa: F401
