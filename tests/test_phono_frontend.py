"""ASL-Phono corpus front-end (slnlp_b200/phono.py) - host logic, no GPU.

The four composition strategies are pinned to vectors generated from the reference's own
DatasetBuilder.compose_* (tests/golden/make_phono_golden.py); corpus reading, the per-sign
frequency filter and the torchtext-0.6 vocabulary order are checked on a corpus written to a
temporary directory in the reference's file layout (dataset_builder.py:66-135)."""
import json
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "sign-language-nlp_b200"))
from slnlp_b200 import phono  # noqa: E402

GOLD = json.load(open(os.path.join(HERE, "golden", "phono_compose.json")))


@pytest.mark.parametrize("strategy", phono.STRATEGIES)
def test_composition_strategies_match_the_reference(strategy):
    assert phono.compose(GOLD["rows"], GOLD["fields"], strategy) == GOLD[strategy]


def test_field_subset_and_unknown_strategy():
    assert phono.compose(GOLD["rows"], GOLD["subset_fields"], "as_words") == GOLD["subset_as_words"]
    with pytest.raises(AssertionError, match="Unknown composition strategy"):
        phono.compose(GOLD["rows"], GOLD["fields"], "as_sentences")


def _write_corpus(tmp_path):
    """5 signs; 'rare' has one file only.  Frames reuse the golden rows."""
    rows = GOLD["rows"]
    spec = {"book": 3, "drink": 2, "hello": 4, "rare": 1, "zebra": 2}
    k = 0
    for sign, n in spec.items():
        for j in range(n):
            nf = 3 + (k % 5)
            frames = [{"phonology": rows[(k + i) % len(rows)], "other": i} for i in range(nf)]
            with open(tmp_path / f"{sign}-{j:02d}.json", "w") as f:
                json.dump({"label": sign, "frames": frames}, f)
            k += 1
    (tmp_path / "notes.txt").write_text("not a sample")
    return spec


def test_corpus_to_tensors(tmp_path):
    spec = _write_corpus(tmp_path)
    ds = phono.build_dataset(str(tmp_path), GOLD["fields"], samples_min_freq=2, composition_strategy="as_words")
    n = sum(v for v in spec.values() if v >= 2)
    assert len(ds) == n and "rare-00.json" not in ds.files
    X, lengths, y = ds.tokens, ds.lengths, ds.labels_
    assert X.dtype == torch.int64 and X.shape == (n, int(lengths.max()))
    # padding with <pad> = 1 beyond each length, real tokens before it
    for i in range(n):
        assert (X[i, lengths[i]:] == 1).all() and (X[i, :lengths[i]] >= 2).all()
    # vocabulary: specials first, then descending frequency, ties in ascending string order
    sv = ds.vocab_X
    assert sv.itos[:2] == ["<unk>", "<pad>"]
    freq = [sv.freqs[w] for w in sv.itos[2:]]
    assert freq == sorted(freq, reverse=True)
    for a, b in zip(sv.itos[2:], sv.itos[3:]):
        if sv.freqs[a] == sv.freqs[b]:
            assert a < b
    assert sv.stoi["never-seen"] == 0 and sv.stoi["<bos>"] == 0     # torchtext defaultdict -> <unk>
    # labels: hello (4 files) is the most frequent class -> index 2
    tv = ds.vocab_y
    assert tv.itos[2] == "hello" and set(tv.itos[2:]) == {"book", "drink", "hello", "zebra"}
    first = ds.files.index("hello-00.json")
    assert int(y[first]) == 2
    # round trip of one sample through the vocabulary
    with open(tmp_path / "book-01.json") as f:
        frames = [fr["phonology"] for fr in json.load(f)["frames"]]
    i = ds.files.index("book-01.json")
    toks = phono.compose(frames, GOLD["fields"], "as_words")
    assert [sv.itos[t] for t in X[i, :lengths[i]].tolist()] == toks
    # the dataset object behaves like the reference's AslDataset for the estimator
    (xi, li), yi = ds[i]
    assert li == len(toks) and yi == tv.stoi["book"]
    test, train = ds.split(lengths=0.25, seed=1)
    assert len(test) + len(train) == n


def test_min_freq_one_keeps_everything_and_strategies_change_the_vocabulary(tmp_path):
    spec = _write_corpus(tmp_path)
    ds1 = phono.build_dataset(str(tmp_path), GOLD["fields"], samples_min_freq=1, composition_strategy="as_words")
    assert len(ds1) == sum(spec.values())
    ds2 = phono.build_dataset(str(tmp_path), GOLD["fields"], samples_min_freq=1, composition_strategy="as_words_norm")
    assert len(ds2.vocab_X) != len(ds1.vocab_X) or ds2.vocab_X.itos != ds1.vocab_X.itos
    with pytest.raises(AssertionError):
        phono.build_dataset(str(tmp_path / "missing"), GOLD["fields"])


def test_helper_load_dataset_reads_a_corpus_directory(tmp_path):
    import helper
    _write_corpus(tmp_path)
    ds = helper.load_dataset(dataset_args={"dataset_dir": str(tmp_path), "fields": GOLD["fields"], "samples_min_freq": 2,
                                           "composition_strategy": "as_words"})
    assert len(ds) == 11 and ds.vocab_y.itos[2] == "hello"
