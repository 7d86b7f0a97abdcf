"""pos_evolution_b200 -- B200-native attestation aggregation and LMD-GHOST fork choice behind the
pyspec function signatures of ethereum/pos-evolution.  Host layer in Python, all
arithmetic in hand-written sm_100a CUDA reached through the C ABI of include/b200pos.h."""
import os as _os

# the pipelined epoch API keeps up to 8 epochs in flight on ~27 CUDA streams; with the default 8 hardware queues streams alias and
# create false dependencies between pipeline slots.  Must be set before the process's first CUDA call.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

__all__ = ["engine", "bls", "spec"]
