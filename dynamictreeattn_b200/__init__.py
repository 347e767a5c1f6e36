"""dynamictreeattn_b200 — B200-native tree attention over prefix-merged token tries.

The package mirrors the reference's module names (token_trie, trie,
tree_training_engine, vocab_parallel, data_parallel, tree_time_model, dense) so a
user of Whisper-6/DynamicTreeAttn can switch imports and keep their code.
Host-side integer planning is pure Python/numpy; everything on the device path
goes through the hand-written sm_100a kernels in ``csrc/`` via a C-ABI shared
library (``include/dta_b200.h``) and fails loudly if that library is missing.
"""
__version__ = "0.1.0"
