"""On-disk formats either side of the hot path.

The reference's experiment scripts exchange four kinds of files; these helpers read and write the same
layouts so its artefacts (and a ``TreeTimeModel`` refit) work unchanged against this engine:

* rollouts: ``torch.save(List[LongTensor])`` in a ``.pt`` file, or a folder of them
  (reference ``run.py:134-137`` / ``run_all.py:21-31``);
* per-batch statistics: one JSON object per line — ``TokenTrie.get_stats`` plus ``loss``, ``time`` and
  ``name`` (``run.py:56-60, 110-114``, ``run_all.py:162-165``) — the input of ``TreeTimeModel.add_data``;
* gradient dumps: ``torch.save({param_name: grad.cpu() | None})`` (``run.py:154-161``);
* the gradient comparison table of ``grad/Qwen3-0.6B-TB-vs-DB-bf16.txt``: per common parameter
  ``|dg|/|g| = ||g_exp - g_base|| / ||g_base||`` with both norms, sorted by the ratio, largest first.

    python -m dynamictreeattn_b200.io_formats compare BASELINE.pt EXPERIMENT.pt [--out table.txt]

Host-side only: nothing here touches the GPU path.
"""
from __future__ import annotations

import argparse
import json
import os
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch

__all__ = ["load_sequences", "load_sequence_folder", "save_sequences", "write_stats_jsonl", "read_stats_jsonl",
           "save_gradients", "load_gradients", "compare_gradients", "format_gradient_table"]


# ------------------------------------------------------------------------------------------------
# rollouts
# ------------------------------------------------------------------------------------------------
def _check_sequences(data, where: str) -> List[torch.Tensor]:
    if not isinstance(data, (list, tuple)):
        raise ValueError(f"{where}: expected a list of 1-D LongTensors, got {type(data).__name__}")
    out = []
    for i, t in enumerate(data):
        if not torch.is_tensor(t) or t.dim() != 1:
            raise ValueError(f"{where}: element {i} is not a 1-D tensor")
        if t.dtype != torch.long:
            if t.is_floating_point():
                raise ValueError(f"{where}: element {i} has dtype {t.dtype}, expected an integer tensor")
            t = t.long()
        out.append(t.cpu())
    return out


def load_sequences(path: str) -> List[torch.Tensor]:
    """One rollout batch: ``.pt`` holding ``List[LongTensor]`` (reference ``run.py:134-137``).  Text files need a
    tokenizer and are outside the path; they raise like the reference does for unknown formats."""
    if not path.endswith(".pt"):
        raise ValueError(f"Unsupported data format: {path}")
    return _check_sequences(torch.load(path, map_location="cpu"), path)


def load_sequence_folder(folder: str) -> List[Tuple[str, List[torch.Tensor]]]:
    """Every ``*.pt`` of a folder as ``(name_without_extension, sequences)``, sorted by path
    (reference ``run_all.py:21-31``)."""
    files = sorted(os.path.join(folder, f) for f in os.listdir(folder) if f.endswith(".pt"))
    return [(os.path.basename(f)[:-3], load_sequences(f)) for f in files]


def save_sequences(path: str, seqs: Sequence[torch.Tensor]) -> None:
    torch.save([t.detach().cpu().long() for t in seqs], path)


# ------------------------------------------------------------------------------------------------
# statistics (JSONL)
# ------------------------------------------------------------------------------------------------
def write_stats_jsonl(path: str, rows: Iterable[dict]) -> None:
    """One ``json.dumps(stat)`` per line (reference ``run_all.py:162-165``)."""
    with open(path, "w") as f:
        for stat in rows:
            f.write(json.dumps(stat) + "\n")


def read_stats_jsonl(path: str) -> List[dict]:
    with open(path) as f:
        return [json.loads(line) for line in f if line.strip()]


# ------------------------------------------------------------------------------------------------
# gradients
# ------------------------------------------------------------------------------------------------
def save_gradients(model, path: str) -> None:
    """``{name: grad.cpu() or None}`` for every named parameter (reference ``run.py:154-161``)."""
    grads = {}
    for name, param in model.named_parameters():
        grads[name] = param.grad.detach().cpu() if param.grad is not None else None
    torch.save(grads, path)


def load_gradients(path: str) -> Dict[str, Optional[torch.Tensor]]:
    d = torch.load(path, map_location="cpu")
    if not isinstance(d, dict):
        raise ValueError(f"{path}: expected a dict of gradients, got {type(d).__name__}")
    return d


def compare_gradients(baseline: Dict[str, Optional[torch.Tensor]], experiment: Dict[str, Optional[torch.Tensor]]):
    """Rows ``(name, |dg|/|g|, |g_baseline|, |g_exp|, max_abs_diff)`` over the parameters present (and not None)
    in both dumps, sorted by the ratio, largest first; norms in fp64."""
    rows = []
    for name, gb in baseline.items():
        ge = experiment.get(name)
        if gb is None or ge is None:
            continue
        if gb.shape != ge.shape:
            raise ValueError(f"{name}: shapes differ, {tuple(gb.shape)} vs {tuple(ge.shape)}")
        b, e = gb.double(), ge.double()
        nb = float(b.norm())
        diff = e - b
        ratio = float(diff.norm()) / nb if nb > 0 else (0.0 if float(diff.norm()) == 0 else float("inf"))
        rows.append((name, ratio, nb, float(e.norm()), float(diff.abs().max()) if diff.numel() else 0.0))
    rows.sort(key=lambda r: -r[1])
    return rows


def format_gradient_table(rows) -> str:
    """The layout of the reference's ``grad/Qwen3-0.6B-TB-vs-DB-bf16.txt`` (plus a summary footer)."""
    lines = ["", f"Comparing {len(rows)} common parameters", "",
             f"{'Parameter':<60s}{'|Δg|/|g|':>13s}{'|g_baseline|':>13s}{'|g_exp|':>13s}", "-" * 104]
    for name, ratio, nb, ne, _ in rows:
        lines.append(f"{name:<60s}{ratio:>13.4e}{nb:>13.4e}{ne:>13.4e}")
    if rows:
        ratios = sorted(r[1] for r in rows)
        med = ratios[len(ratios) // 2] if len(ratios) % 2 else 0.5 * (ratios[len(ratios) // 2 - 1] + ratios[len(ratios) // 2])
        lines += ["-" * 104, f"max {ratios[-1]:.4e}   median {med:.4e}   max |Δg| {max(r[4] for r in rows):.4e}"]
    return "\n".join(lines) + "\n"


def _main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="python -m dynamictreeattn_b200.io_formats")
    sub = ap.add_subparsers(dest="cmd", required=True)
    c = sub.add_parser("compare", help="gradient comparison table of two dumps written by save_gradients")
    c.add_argument("baseline")
    c.add_argument("experiment")
    c.add_argument("--out", default=None)
    args = ap.parse_args(argv)
    text = format_gradient_table(compare_gradients(load_gradients(args.baseline), load_gradients(args.experiment)))
    if args.out:
        with open(args.out, "w") as f:
            f.write(text)
    print(text, end="")
    return 0


if __name__ == "__main__":
    raise SystemExit(_main())
