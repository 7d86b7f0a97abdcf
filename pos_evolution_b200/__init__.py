"""pos_evolution_b200 -- B200-native attestation aggregation and LMD-GHOST fork choice behind the
pyspec function signatures of ethereum/pos-evolution.  Host layer in Python, all
arithmetic in hand-written sm_100a CUDA reached through the C ABI of include/b200pos.h."""
__all__ = ["engine", "bls", "spec"]
