"""Seeded trie generators for the tests: re-exported from the package (dynamictreeattn_b200/workloads.py)."""
from dynamictreeattn_b200.workloads import (bary, config1_seqs, cpu_sample_trie, random_seqs, synth_trie,  # noqa: F401
                                            tau2_trie, trunk_fan)
