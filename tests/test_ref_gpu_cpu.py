"""Plumbing self-test of oracle/ref_gpu.py (the harness that times the UNMODIFIED reference on the GPU for bench.py's
`reference_gpu` block) with --device cpu on tiny models, in a subprocess: the reference's top-level module names
(Engine / Tree / utils) collide with this repository's drop-in shims, so it can never share a process with the tests.
Skipped when oracle/_ref has not been vendored (tools/vendor_ref.py needs /root/reference)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import json, sys, torch
sys.path.insert(0, r"%(root)s/oracle")
import ref_gpu
sys.path.append(r"%(root)s")
from sequoia_b200.model import LlamaConfigLite
spec = dict(draft="d", target="t", growmap="L40_growmaps/4x4-tree.pt", greedy=%(greedy)s, T=0.6, top_p=1.0, M=128, prefix=32,
            max_len=48, _cfgs={"d": LlamaConfigLite(64, 128, 1, 4, 4), "t": LlamaConfigLite(64, 128, 2, 4, 2)})
out = ref_gpu.run(spec, 4, 3, 2, r"%(trace)s", device="cpu")
tr = torch.load(r"%(trace)s")
out["trace_len"] = len(tr)
out["trace_tokens"] = int(tr[0]["tree_tokens"].numel())
print("RESULT " + json.dumps(out))
"""


@pytest.mark.skipif(not os.path.isfile(os.path.join(ROOT, "oracle", "_ref", "utils.py")), reason="oracle/_ref not vendored")
@pytest.mark.parametrize("greedy", [False, True])
def test_reference_harness_runs_the_vendored_reference(greedy, tmp_path):
    code = SCRIPT % {"root": ROOT, "greedy": greedy, "trace": str(tmp_path / "trace.pt")}
    r = subprocess.run([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")][-1]
    out = json.loads(line[len("RESULT "):])
    assert out["impl"] == "reference_gpu" and out["value"] > 0 and out["steps"] == 4
    assert out["trace_len"] == 2 and out["trace_tokens"] == 16            # 4x4 tree: 17 nodes - root
    assert all(a >= 32 for a in out["first_iter_accept_lens"])


def test_vendor_manifest_matches_files():
    """The vendored copy is byte-identical to what the recipe recorded (nobody edited the reference)."""
    import hashlib
    mf = os.path.join(ROOT, "oracle", "_ref", "MANIFEST.json")
    if not os.path.isfile(mf):
        pytest.skip("oracle/_ref not vendored")
    with open(mf) as f:
        m = json.load(f)
    for rel, sha in m["files"].items():
        with open(os.path.join(ROOT, "oracle", "_ref", rel), "rb") as fh:
            assert hashlib.sha256(fh.read()).hexdigest() == sha, rel
        src = os.path.join(m["source"], rel)
        if os.path.isfile(src):
            with open(src, "rb") as fh:
                assert hashlib.sha256(fh.read()).hexdigest() == sha, f"{rel} differs from {src}"
