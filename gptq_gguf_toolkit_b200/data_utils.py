"""Calibration data -- host-side mirror of the reference's quant/gptq/src/data_utils.py (same entry point `get_data`, same
sampling rules, so that a run sees the SAME token sequences as the reference for the same arguments):

  file path      a torch-saved list of [1, L] int64 tensors, the first num_tokens // L of them, trimmed to L      (:134-136)
  wikitext2      train: num_tokens // L random windows of the "\\n\\n"-joined train split, positions drawn with the global
                 `random` module (the caller seeds it, model_utils.fix_seed); eval: consecutive windows of the test split (:35-58)
  c4             train: shard 00000 at the pinned revision; documents are appended (joined by the tokens of "\\n\\n") until a
                 sample reaches L tokens, which is then TRIMMED to L and a new sample starts (collect_samples_with_join, :14-32);
                 eval: the first 1100 validation documents, consecutive windows                                     (:89-121)
  fineweb_edu    sample-10BT: first half (train) / second half (eval) of the rows, shuffled with seed 0; every document is cut
                 into pieces of at most L tokens and ALL pieces are kept, the short tails included, until num_tokens tokens are
                 loaded -- samples therefore have different lengths                                                 (:61-86)
  random:<vocab> synthetic token ids (this repository's offline benchmark, not in the reference)

The hub datasets need the `datasets` package and network access, exactly like the reference.
"""
from __future__ import annotations

import os
import random
from typing import Iterable, List

import torch

C4_REVISION = "607bd4c8450a42878aa9ddc051a65a055450ef87"      # the reference pins this revision (:100, :113)


def synthetic_tokens(num_sequences: int, sequence_length: int, vocab_size: int, seed: int = 1) -> List[torch.Tensor]:
    """The benchmark's calibration set: torch.randint(0, vocab, (1, L)) x N, generator seed 1."""
    g = torch.Generator().manual_seed(seed)
    return [torch.randint(0, vocab_size, (1, sequence_length), generator=g) for _ in range(num_sequences)]


def _ids(tokenizer, text: str) -> torch.Tensor:
    return tokenizer(text, return_tensors="pt", add_special_tokens=False).input_ids


def _load_dataset(*args, **kwargs):
    try:
        from datasets import load_dataset
    except Exception as e:  # pragma: no cover
        raise RuntimeError("the hub datasets need the `datasets` package and network access") from e
    return load_dataset(*args, **kwargs)


def _consecutive_windows(tokens: torch.Tensor, sequence_length: int) -> List[torch.Tensor]:
    return [tokens[:, i * sequence_length:(i + 1) * sequence_length] for i in range(tokens.numel() // sequence_length)]


def collect_samples_with_join(data_iter: Iterable, tokenizer, num_samples: int, sequence_length: int, text_key: str = "text"):
    """data_utils.py:14-32: grow a sample document by document; once it holds >= L tokens keep its first L and start over,
    otherwise append the separator tokens and continue."""
    data: List[torch.Tensor] = []
    sep = None
    cur = torch.tensor([], dtype=torch.int64)
    for sample in data_iter:
        cur = torch.cat([cur, _ids(tokenizer, sample[text_key])], dim=1)
        if cur.numel() >= sequence_length:
            data.append(cur[:, :sequence_length])
            cur = torch.tensor([], dtype=torch.int64)
        else:
            if sep is None:
                sep = _ids(tokenizer, "\n\n")
            cur = torch.cat([cur, sep], dim=1)
        if len(data) >= num_samples:
            break
    return data


def get_wikitext2(num_samples: int, sequence_length: int, tokenizer, train: bool = True):
    split = "train" if train else "test"
    tokens = _ids(tokenizer, "\n\n".join(_load_dataset("wikitext", "wikitext-2-raw-v1", split=split)["text"]))
    if not train:
        return _consecutive_windows(tokens, sequence_length)
    data = []
    for _ in range(num_samples):
        i = random.randint(0, tokens.shape[1] - sequence_length - 1)      # the global RNG, like the reference (:47)
        data.append(tokens[:, i:i + sequence_length])
    return data


def get_fineweb_edu(num_tokens: int, sequence_length: int, tokenizer, train: bool = True):
    dataset = _load_dataset("HuggingFaceFW/fineweb-edu", "sample-10BT", split="train")
    half = dataset.num_rows // 2
    dataset = dataset.select(range(half) if train else range(half, dataset.num_rows)).shuffle(seed=0)
    data, left = [], num_tokens
    it = iter(dataset)
    while left > 0:
        piece = _ids(tokenizer, next(it)["text"])
        piece = piece[:, :min(piece.shape[1], left)]
        while piece.shape[1] > sequence_length:          # long documents become several samples, nothing is thrown away
            data.append(piece[:, :sequence_length])
            piece = piece[:, sequence_length:]
            left -= sequence_length
        data.append(piece)                               # the tail: shorter than (or exactly) L
        left -= piece.shape[1]
    return data


def get_c4(num_samples: int, sequence_length: int, tokenizer, train: bool = True):
    if train:
        dataset = _load_dataset("allenai/c4", "default", data_files={"train": "en/c4-train.00000-of-01024.json.gz"},
                                split="train", revision=C4_REVISION)
        return collect_samples_with_join(iter(dataset), tokenizer, num_samples, sequence_length)
    dataset = _load_dataset("allenai/c4", "default", data_files={"validation": "en/c4-validation.00000-of-00008.json.gz"},
                            split="validation[:1100]", revision=C4_REVISION)
    return _consecutive_windows(_ids(tokenizer, "\n\n".join(dataset["text"])), sequence_length)


def get_data(data_name_or_path: str, num_tokens: int, sequence_length: int, tokenizer=None, train: bool = True):
    """data_utils.py:125-146.  Only fineweb_edu is loaded at token granularity; the others take num_tokens // L samples."""
    if os.path.isfile(data_name_or_path):
        data = torch.load(data_name_or_path)[: num_tokens // sequence_length]
        return [sample[:, :sequence_length] for sample in data]
    if data_name_or_path.startswith("random:"):
        return synthetic_tokens(num_tokens // sequence_length, sequence_length, int(data_name_or_path.split(":")[1]))
    if data_name_or_path == "wikitext2":
        return get_wikitext2(num_tokens // sequence_length, sequence_length, tokenizer, train)
    if data_name_or_path == "c4":
        return get_c4(num_tokens // sequence_length, sequence_length, tokenizer, train)
    if data_name_or_path == "fineweb_edu":
        return get_fineweb_edu(num_tokens, sequence_length, tokenizer, train)
    raise ValueError("Unknown dataset.")
