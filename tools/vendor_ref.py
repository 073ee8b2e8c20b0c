#!/usr/bin/env python
"""Recipe that makes the UNMODIFIED reference runnable on the GPU box (where /root/reference does not exist).

The reference is pure Python with no build system (no setup.py / pyproject.toml => `pip install --target baseline/_ref
/root/reference` cannot work), so the recipe copies the modules of the hot path -- Engine/*.py, Tree/*.py, utils.py --
(plus the tests/testbed*.py drivers and the bundled pre-tokenised openwebtext_eval prompts) byte for byte from
/root/reference into the git-ignored `oracle/_ref/` (it travels with gpurun snapshots exactly like a
built .so; it is never committed) and writes a manifest of sha256 sums.  `__graft_entry__.build()` runs it whenever
/root/reference is present; on the GPU box the prebuilt copy is used as is.

Consumers: `oracle/ref_gpu.py` only (test / measurement infrastructure; the product never imports it).
"""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "oracle", "_ref")
PARTS = ("Engine", "Tree")
FILES = ("utils.py", "tests/testbed.py", "tests/testbed_greedy.py")      # the drivers run VERBATIM by tools/run_reference_testbed.py
DATA = ("dataset/openwebtext_eval",)                                     # bundled pre-tokenised prompts (no tokenizer / hub)


def vendor(src: str = SRC, dst: str = DST) -> bool:
    if not os.path.isdir(src):
        return os.path.isdir(dst)
    manifest = {}
    os.makedirs(dst, exist_ok=True)
    for part in PARTS:
        os.makedirs(os.path.join(dst, part), exist_ok=True)
        for f in sorted(os.listdir(os.path.join(src, part))):
            if f.endswith(".py"):
                shutil.copyfile(os.path.join(src, part, f), os.path.join(dst, part, f))
                manifest[f"{part}/{f}"] = None
    for f in FILES:
        os.makedirs(os.path.dirname(os.path.join(dst, f)), exist_ok=True)
        shutil.copyfile(os.path.join(src, f), os.path.join(dst, f))
        manifest[f] = None
    for d in DATA:
        if os.path.isdir(os.path.join(dst, d)):
            shutil.rmtree(os.path.join(dst, d))
        shutil.copytree(os.path.join(src, d), os.path.join(dst, d))
    for rel in manifest:
        with open(os.path.join(dst, rel), "rb") as fh:
            manifest[rel] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(dst, "MANIFEST.json"), "w") as fh:
        json.dump({"source": src, "files": manifest}, fh, indent=1)
    return True


if __name__ == "__main__":
    ok = vendor()
    print(f"oracle/_ref {'ready' if ok else 'unavailable (no /root/reference and no prebuilt copy)'}")
    sys.exit(0 if ok else 1)
