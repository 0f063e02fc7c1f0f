"""Calibration data -- mirror of the reference's quant/gptq/src/data_utils.py::get_data.

The file branch (data_utils.py:134-136: a torch-saved list of [1, L] int64 tensors) and a synthetic
`random:<vocab>` source work offline; the hub datasets (wikitext2 / c4 / fineweb_edu) need network access and
the `datasets` package, exactly like the reference.
"""
from __future__ import annotations

import os
from typing import List

import torch


def synthetic_tokens(num_sequences: int, sequence_length: int, vocab_size: int, seed: int = 1) -> List[torch.Tensor]:
    """The benchmark's calibration set: torch.randint(0, vocab, (1, L)) x N, generator seed 1."""
    g = torch.Generator().manual_seed(seed)
    return [torch.randint(0, vocab_size, (1, sequence_length), generator=g) for _ in range(num_sequences)]


def get_data(data_name_or_path: str, num_tokens: int, sequence_length: int, tokenizer=None, train: bool = True):
    if os.path.isfile(data_name_or_path):                                     # data_utils.py:134-136
        data = torch.load(data_name_or_path)[: num_tokens // sequence_length]
        return [sample[:, :sequence_length] for sample in data]
    if data_name_or_path.startswith("random:"):
        return synthetic_tokens(num_tokens // sequence_length, sequence_length, int(data_name_or_path.split(":")[1]))
    if data_name_or_path in ("wikitext2", "c4", "fineweb_edu"):
        try:
            from datasets import load_dataset  # noqa: F401
        except Exception as e:  # pragma: no cover
            raise RuntimeError(f"{data_name_or_path} needs the `datasets` package and network access") from e
        return _hub_dataset(data_name_or_path, num_tokens, sequence_length, tokenizer, train)
    raise ValueError("Unknown dataset.")


def _hub_dataset(name, num_tokens, sequence_length, tokenizer, train):  # pragma: no cover (needs network)
    from datasets import load_dataset
    n = num_tokens // sequence_length
    if name == "wikitext2":
        ds = load_dataset("wikitext", "wikitext-2-raw-v1", split="train" if train else "test")
        ids = tokenizer("\n\n".join(ds["text"]), return_tensors="pt", add_special_tokens=False).input_ids
    elif name == "c4":
        ds = load_dataset("allenai/c4", "default", data_files={"train": "en/c4-train.00000-of-01024.json.gz"}, split="train")
        ids = tokenizer("\n\n".join(ds["text"][: 4 * n]), return_tensors="pt", add_special_tokens=False).input_ids
    else:
        ds = load_dataset("HuggingFaceFW/fineweb-edu", "sample-10BT", split="train", streaming=True)
        buf, tot = [], 0
        for row in ds:
            t = tokenizer(row["text"], return_tensors="pt", add_special_tokens=False).input_ids
            buf.append(t)
            tot += t.numel()
            if tot >= num_tokens:
                break
        ids = torch.cat(buf, dim=1)
    return [ids[:, i * sequence_length:(i + 1) * sequence_length] for i in range(min(n, ids.shape[1] // sequence_length))]
