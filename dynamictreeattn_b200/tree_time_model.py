"""Five-feature non-negative least-squares cost model of one trie step.

Drop-in for the reference module ``tree_time_model.py`` (tree_time_model.py:5-48):

    T = c0*n_leaf_sequences + c1*n_tree_tokens + c2*n_f1_tokens
        + c3*sum_prefix_len + c4*sum_depth

fitted with NNLS on a sliding window of 16..1024 JSONL stat records; before the
first fit ``pred`` falls back to ``n_tree_tokens`` (tree_time_model.py:41-43).
For the B200 engine c2 and c3 model overheads the single-pass design no longer
has (F1 recompute, prefix-KV concat); NNLS simply drives them to ~0 when the
model is refit on B200 timings.
"""
from __future__ import annotations

import json
from typing import Iterable, List

import numpy as np
from scipy.optimize import nnls

__all__ = ["TreeTimeModel", "FEATURES"]

FEATURES = ("n_leaf_sequences", "n_tree_tokens", "n_f1_tokens", "sum_prefix_len", "sum_depth")
_OPTIONAL = {"n_f1_tokens"}          # absent from forward-mode stats


def _row(stats: dict) -> List[float]:
    return [stats.get(k, 0) if k in _OPTIONAL else stats[k] for k in FEATURES]


class TreeTimeModel:
    MIN_N_DATA_POINTS = 16
    MAX_N_DATA_POINTS = 1024

    def __init__(self):
        self.coeffs = None
        self.data: List[dict] = []

    def fit(self) -> float:
        """NNLS fit on the current window; returns the mean squared error."""
        X = np.array([_row(s) for s in self.data])
        y = np.array([s["time"] for s in self.data])
        self.coeffs, _ = nnls(X, y)
        resid = X @ self.coeffs - y
        return float(np.mean(resid ** 2))

    def add_data(self, data: Iterable[dict]) -> None:
        self.data.extend(data)
        if len(self.data) > self.MAX_N_DATA_POINTS:
            del self.data[: len(self.data) - self.MAX_N_DATA_POINTS]
        if len(self.data) >= self.MIN_N_DATA_POINTS:
            self.fit()

    def pred(self, stats: dict):
        if self.coeffs is None:
            return stats["n_tree_tokens"]
        c = self.coeffs
        # same association order as the reference so predictions are bit-identical
        return (c[0] * stats["n_leaf_sequences"] + c[1] * stats["n_tree_tokens"]
                + c[2] * stats.get("n_f1_tokens", 0) + c[3] * stats["sum_prefix_len"]
                + c[4] * stats["sum_depth"])

    # convenience used by the DP runtime / bench (not in the reference)
    def add_jsonl(self, path: str) -> None:
        with open(path) as f:
            self.add_data([json.loads(line) for line in f if line.strip()])
