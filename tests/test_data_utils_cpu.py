"""CPU: the calibration-data sampling rules (gptq_gguf_toolkit_b200/data_utils.py) against the reference's own
quant/gptq/src/data_utils.py run on the SAME fake hub datasets and tokenizer (no network): identical sample lists for
wikitext2 (random windows from the seeded global RNG), c4 (join-until-L, trimmed) and fineweb_edu (first half, shuffle seed 0,
per-document pieces with their short tails), train and eval splits.  Skipped where the reference is not mounted."""
import os
import random
import sys
import types

import pytest
import torch

REF = "/root/reference/quant/gptq"


class FakeTok:
    def __call__(self, text, return_tensors="pt", add_special_tokens=False):
        ids = [ord(c) % 251 for c in text]
        return types.SimpleNamespace(input_ids=torch.tensor([ids], dtype=torch.int64))


class FakeDS:
    def __init__(self, rows):
        self.rows = list(rows)
        self.num_rows = len(self.rows)

    def select(self, idx):
        return FakeDS([self.rows[i] for i in idx])

    def shuffle(self, seed):
        r = random.Random(seed)
        rows = list(self.rows)
        r.shuffle(rows)
        return FakeDS(rows)

    def __iter__(self):
        return iter(self.rows)

    def __getitem__(self, key):
        return [r[key] for r in self.rows]


def fake_load_dataset(name, config=None, split="train", **kw):
    g = random.Random(hash((name, split)) % 1000)
    n = 60 if "fineweb" in name else 40
    rows = [{"text": "".join(chr(97 + g.randrange(26)) for _ in range(g.randrange(5, 90)))} for _ in range(n)]
    if split.endswith("[:1100]"):
        rows = rows[:1100]
    return FakeDS(rows)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference not mounted")
@pytest.mark.parametrize("name", ["wikitext2", "c4", "fineweb_edu"])
@pytest.mark.parametrize("train", [True, False])
def test_sampling_rules_equal_reference(monkeypatch, name, train):
    from gptq_gguf_toolkit_b200 import data_utils as ours
    sys.path.insert(0, REF)
    try:
        import datasets
        monkeypatch.setattr(datasets, "load_dataset", fake_load_dataset)
        for m in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
            del sys.modules[m]
        from src import data_utils as ref
    finally:
        sys.path.remove(REF)
    monkeypatch.setattr(ref, "load_dataset", fake_load_dataset)
    monkeypatch.setattr(ours, "_load_dataset", fake_load_dataset)
    tok = FakeTok()
    random.seed(0)
    want = ref.get_data(name, 640, 32, tok, train)
    random.seed(0)
    got = ours.get_data(name, 640, 32, tok, train)
    assert len(got) == len(want) and len(got) > 0
    assert all(torch.equal(a, b) for a, b in zip(got, want))
    if name == "fineweb_edu":
        assert len({t.shape[1] for t in got}) > 1, "fineweb_edu keeps the short tails of the documents"


def test_file_and_synthetic_sources(tmp_path):
    from gptq_gguf_toolkit_b200.data_utils import get_data
    data = [torch.arange(40).view(1, 40) + i for i in range(7)]
    p = str(tmp_path / "calib.pt")
    torch.save(data, p)
    got = get_data(p, 5 * 32, 32)
    assert len(got) == 5 and all(t.shape == (1, 32) for t in got) and torch.equal(got[2], data[2][:, :32])
    syn = get_data("random:100", 4 * 16, 16)
    assert len(syn) == 4 and int(torch.stack(syn).max()) < 100
    with pytest.raises(ValueError):
        get_data("nope", 10, 5)
