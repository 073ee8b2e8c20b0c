#!/usr/bin/env python
"""Run the reference's OWN driver -- tests/testbed.py (or tests/testbed_greedy.py), byte-identical to the reference's file
(vendored by tools/vendor_ref.py, sha256 in oracle/_ref/MANIFEST.json) -- against THIS repository's drop-in modules.

This is north_star's "tests/testbed.py drops in unchanged": the driver imports `Engine.Engine`, `Tree.SpecTree`, `utils`,
`data_converter` (here: the sequoia_b200 drop-ins at the repository root) and runs its simulation_fast loop.  What the
offline GPU box cannot provide is supplied from outside the driver, without touching it:
  * `accelerate` is not installed            -> a stub module whose Accelerator().prepare() is the identity;
  * the Llama-2 tokenizer needs the HF hub   -> AutoTokenizer.from_pretrained returns a tiny local tokenizer; with
                                                `--dataset openwebtext` the driver only uses it to PAD the bundled,
                                                already tokenised prompts (dataset/openwebtext_eval of the reference);
  * weights                                   -> `--model / --target random-init:<name>[:seed]` (engine feature);
  * this image's `datasets` expects a newer torchvision (`torchvision.io.VideoReader`) -> a placeholder attribute.

    python tools/run_reference_testbed.py [--driver testbed.py] -- --model random-init:llama-68m:1 \
        --target random-init:llama-2-7b:2 --growmap <abs path> --T 0.6 --P 1.0 --M 384 --dataset openwebtext --start 0 --end 20
"""
import os
import runpy
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TESTS = os.path.join(ROOT, "oracle", "_ref", "tests")


def main():
    argv = sys.argv[1:]
    driver = "testbed.py"
    if argv and argv[0] == "--driver":
        driver, argv = argv[1], argv[2:]
    if argv and argv[0] == "--":
        argv = argv[1:]
    path = os.path.join(REF_TESTS, driver)
    if not os.path.isfile(path):
        sys.exit(f"{path} missing: run tools/vendor_ref.py in the build container")
    import transformers  # noqa: F401  (before the accelerate stub, see tests/golden/ref_shim.py)
    import transformers.models.llama.modeling_llama  # noqa: F401
    acc = types.ModuleType("accelerate")

    class Accelerator:                       # tests/testbed.py:288-289: accelerator.prepare(dataloader)
        def prepare(self, *objs):
            return objs[0] if len(objs) == 1 else objs

    acc.Accelerator = Accelerator
    acc.cpu_offload = lambda model, execution_device=None: model
    sys.modules.setdefault("accelerate", acc)

    from tokenizers import Tokenizer
    from tokenizers.models import WordLevel
    from transformers import AutoTokenizer, PreTrainedTokenizerFast

    def local_tokenizer(*a, **k):
        tok = Tokenizer(WordLevel({"<unk>": 0, "<s>": 1, "</s>": 2}, unk_token="<unk>"))
        return PreTrainedTokenizerFast(tokenizer_object=tok, unk_token="<unk>", bos_token="<s>", eos_token="</s>")

    AutoTokenizer.from_pretrained = staticmethod(local_tokenizer)
    try:                                     # this image's `datasets` torch formatter imports a class its torchvision lacks
        import torchvision.io as tvio
        if not hasattr(tvio, "VideoReader"):
            tvio.VideoReader = type("VideoReader", (), {})
    except Exception:
        pass
    sys.path.insert(0, ROOT)                 # Engine / Tree / utils / data_converter = this repository's drop-ins
    os.chdir(REF_TESTS)                      # the driver's relative paths ("..", "../dataset/openwebtext_eval")
    sys.argv = [path] + argv
    runpy.run_path(path, run_name="__main__")
    import Engine.Engine as E
    assert E.__file__.startswith(ROOT) and "_ref" not in E.__file__, "the driver must have run on the sequoia_b200 drop-ins"
    print(f"[run_reference_testbed] {driver} (verbatim) ran on {E.__file__}")


if __name__ == "__main__":
    main()
