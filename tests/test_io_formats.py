"""On-disk formats: rollouts .pt, stats JSONL, gradient dumps and the comparison table."""
import os

import pytest
import torch

from dynamictreeattn_b200 import io_formats as io
from dynamictreeattn_b200.token_trie import TokenTrie
from dynamictreeattn_b200.tree_time_model import TreeTimeModel
from trie_cases import random_seqs

REF_TABLE = "/root/reference/grad/Qwen3-0.6B-TB-vs-DB-bf16.txt"


def test_sequences_round_trip_and_folder(tmp_path):
    a, b = random_seqs(1, 7), random_seqs(2, 5)
    io.save_sequences(str(tmp_path / "b_batch.pt"), b)
    io.save_sequences(str(tmp_path / "a_batch.pt"), a)
    (tmp_path / "notes.txt").write_text("ignored")
    got = io.load_sequences(str(tmp_path / "a_batch.pt"))
    assert len(got) == len(a) and all(torch.equal(x, y) and x.dtype == torch.long for x, y in zip(got, a))
    folder = io.load_sequence_folder(str(tmp_path))
    assert [n for n, _ in folder] == ["a_batch", "b_batch"]          # sorted by path, extension stripped
    assert all(torch.equal(x, y) for x, y in zip(folder[1][1], b))
    TokenTrie(folder[0][1])                                          # what the loader returns feeds the builder as is


def test_sequences_reject_bad_files(tmp_path):
    with pytest.raises(ValueError, match="Unsupported data format"):
        io.load_sequences(str(tmp_path / "x.json"))
    torch.save({"not": "a list"}, str(tmp_path / "d.pt"))
    with pytest.raises(ValueError):
        io.load_sequences(str(tmp_path / "d.pt"))
    torch.save([torch.zeros(2, 3, dtype=torch.long)], str(tmp_path / "m.pt"))
    with pytest.raises(ValueError, match="1-D"):
        io.load_sequences(str(tmp_path / "m.pt"))
    torch.save([torch.zeros(3)], str(tmp_path / "f.pt"))
    with pytest.raises(ValueError, match="integer"):
        io.load_sequences(str(tmp_path / "f.pt"))
    torch.save([torch.arange(4, dtype=torch.int32)], str(tmp_path / "i.pt"))
    assert io.load_sequences(str(tmp_path / "i.pt"))[0].dtype == torch.long


def test_stats_jsonl_feeds_time_model(tmp_path):
    rows = []
    for seed in range(20):
        trie = TokenTrie(random_seqs(seed, 12, vocab=4, max_len=24))
        trie.backward_permute()
        st = trie.get_stats("backward", 8)
        st["loss"] = 1.0
        st["time"] = 1e-3 * st["n_tree_tokens"] + 1e-5 * st["sum_depth"]
        st["name"] = f"batch{seed}"
        rows.append(st)
    path = str(tmp_path / "stats.jsonl")
    io.write_stats_jsonl(path, rows)
    back = io.read_stats_jsonl(path)
    assert back == rows
    tm = TreeTimeModel()
    tm.add_data(back)                       # >= 16 rows: fits on add
    assert tm.coeffs is not None
    assert abs(tm.pred(back[3]) - back[3]["time"]) < 1e-6 + 1e-3 * back[3]["time"]


class _Tiny(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a = torch.nn.Linear(4, 3)
        self.unused = torch.nn.Parameter(torch.zeros(2))


def test_gradient_dump_compare_and_cli(tmp_path, capsys):
    torch.manual_seed(0)
    m = _Tiny()
    m.a(torch.randn(5, 4)).sum().backward()
    p0, p1 = str(tmp_path / "g0.pt"), str(tmp_path / "g1.pt")
    io.save_gradients(m, p0)
    g = io.load_gradients(p0)
    assert set(g) == {"a.weight", "a.bias", "unused"} and g["unused"] is None
    g1 = {k: (None if v is None else v.clone()) for k, v in g.items()}
    g1["a.weight"] = g1["a.weight"] * 1.5
    torch.save(g1, p1)
    rows = io.compare_gradients(g, g1)
    assert [r[0] for r in rows] == ["a.weight", "a.bias"]            # largest ratio first, None skipped
    assert abs(rows[0][1] - 0.5) < 1e-6 and rows[1][1] == 0.0          # x1.5 rounded in fp32
    assert abs(rows[0][3] - 1.5 * rows[0][2]) < 1e-5
    assert io._main(["compare", p0, p1, "--out", str(tmp_path / "t.txt")]) == 0
    out = capsys.readouterr().out
    assert "Comparing 2 common parameters" in out and (tmp_path / "t.txt").read_text() == out
    g1["a.bias"] = torch.zeros(7)
    with pytest.raises(ValueError, match="shapes differ"):
        io.compare_gradients(g, g1)


@pytest.mark.skipif(not os.path.exists(REF_TABLE), reason="reference artefact not present")
def test_table_layout_matches_reference_artefact():
    """Header and row layout are those of the reference's recorded table."""
    ref = open(REF_TABLE, encoding="utf-8").read().splitlines()
    rows = []
    for line in ref[5:8]:
        name, ratio, nb, ne = line.split()
        rows.append((name, float(ratio), float(nb), float(ne), 0.0))
    ours = io.format_gradient_table(rows).splitlines()
    assert ours[1] == "Comparing 3 common parameters" and ref[1].startswith("Comparing ")
    assert ours[3] == ref[3] and ours[4] == ref[4]
    assert ours[5:8] == ref[5:8]
