"""Prompt sources for the drivers (import path of the reference's data_converter.py, tests/testbed.py:12).

The reference pulls its evaluation text from the HF hub (network).  Those loaders are kept importable with the same
names and signatures, implemented over one helper; the benchmark of this repository uses `synthetic_prompts`, which
needs neither network nor tokenizer (torch.randint(3, 32000, (128,)) prompts, seed 17).
"""
from __future__ import annotations

from typing import List

import torch


def synthetic_prompts(n: int, length: int = 128, vocab_size: int = 32000, seed: int = 17) -> List[torch.Tensor]:
    """Deterministic random token prompts (ids in [3, vocab) so that EOS=2 / pad=0 never appear)."""
    g = torch.Generator().manual_seed(seed)
    return [torch.randint(3, vocab_size, (length,), generator=g) for _ in range(n)]


def _hub_text_dataset(tokenizer, hub_args: tuple, hub_kwargs: dict, text_column: str, seq_len: int, padding):
    from datasets import load_dataset  # lazy: needs network access
    ds = load_dataset(*hub_args, **hub_kwargs)
    drop = [c for c in ds.column_names]

    def tok(batch):
        return tokenizer(batch[text_column], return_tensors="pt", max_length=seq_len, padding=padding, truncation=True)

    ds = ds.map(tok, batched=True, remove_columns=drop)
    ds.set_format(type="torch", columns=["input_ids", "attention_mask"])
    return ds


def convert_wiki_dataset(tokenizer, seq_len=256):
    return _hub_text_dataset(tokenizer, ("wikimedia/wikipedia", "20231101.en"), {"split": "train[0:2000]"}, "text",
                             seq_len, True)


def convert_cnn_dataset(tokenizer, seq_len=256):
    return _hub_text_dataset(tokenizer, ("cnn_dailymail", "1.0.0"), {"split": "test[0:2000]"}, "article", seq_len, True)


def convert_wikimqa_dataset(tokenizer, seq_len=256):
    return _hub_text_dataset(tokenizer, ("THUDM/LongBench", "2wikimqa_e"), {"split": "test"}, "context", seq_len,
                             "max_length")


def convert_qasper_dataset(tokenizer, seq_len=256):
    return _hub_text_dataset(tokenizer, ("THUDM/LongBench", "qasper_e"), {"split": "test"}, "context", seq_len,
                             "max_length")


def convert_c4_dataset_eval(tokenizer, seq_len=256):
    return _hub_text_dataset(tokenizer, ("allenai/c4",),
                             {"data_files": {"validation": "en/c4-validation.00000-of-00008.json.gz"},
                              "split": "validation[:2000]"}, "text", seq_len, True)


def convert_dataset(tokenizer, file_path):
    """Pre-tokenised json rows {'input_ids': [...]} -> {'input_ids', 'labels'} (pad positions labelled -100)."""
    from datasets import load_dataset
    ds = load_dataset("json", data_files=file_path, split="train")
    pad = tokenizer.pad_token_id

    def to_lm(batch):
        ids = torch.tensor(batch["input_ids"], dtype=torch.float32)
        labels = ids.clone()
        if pad is not None:
            labels[labels == pad] = -100
        return {"input_ids": ids, "labels": labels}

    extra = [c for c in ds.column_names if c not in ("input_ids",)]
    ds = ds.map(to_lm, batched=True, remove_columns=extra)
    ds.set_format(type="torch", columns=["input_ids", "labels"])
    return ds
