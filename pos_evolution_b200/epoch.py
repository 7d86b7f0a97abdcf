"""EpochProcessor -- the batch (array-level) public API for one rank / one GPU.

One call processes a whole epoch of attestations for the validators this rank owns:
  individual G2 signatures --bls.Aggregate per committee--> aggregate signatures
  --FastAggregateVerify (registry-indexed pubkey gather + hash-to-G2 + pairing)--> verdicts
  --update_latest_messages (accepted aggregates only)--> LMD table on the device
  --get_weight scatter--> per-block direct votes  [--NCCL all-reduce over ranks--]  --get_head--> head index.
(pos-evolution.md:722-754, :963-979, :1102-1116, :1435-1441.)  Device-resident inputs
(`process_epoch_dev`) are torch CUDA tensors; `process_epoch_host` takes pinned host tensors and does
the H2D/D2H copies itself -- that is the end-to-end call bench.py times.

Multi-GPU: validators (and with them committees / aggregates) are sharded across ranks with no
data-path exchange until the vote weights: u64[n_blocks] direct votes are summed with one
torch.distributed all_reduce (NCCL over NVLink; int64 two's-complement sum == u64 sum), after which
every rank finishes get_head on its replica of the block tree.
"""
import torch

from .engine import Engine


class EpochProcessor:
    def __init__(self, engine: Engine, n_agg: int, n_sig: int, bits_stride: int, n_blocks: int, process_group=None, device=None):
        self.eng = engine
        self.dev = device if device is not None else torch.device("cuda", engine.device)
        self.pg = process_group
        self.n_agg, self.n_sig, self.n_blocks = n_agg, n_sig, n_blocks
        d = self.dev
        self.d_agg_sig = torch.zeros((n_agg, 96), dtype=torch.uint8, device=d)
        self.d_agg_status = torch.zeros(n_agg, dtype=torch.int32, device=d)
        self.d_ok = torch.zeros(n_agg, dtype=torch.uint8, device=d)
        self.d_votes = torch.zeros(n_blocks, dtype=torch.int64, device=d)
        self.d_head = torch.zeros(1, dtype=torch.int32, device=d)
        # staging for the host entry point
        self.d_sigs = torch.zeros((n_sig, 96), dtype=torch.uint8, device=d)
        self.d_bits = torch.zeros((n_agg, bits_stride), dtype=torch.uint8, device=d)
        self.d_msgs = torch.zeros((n_agg, 32), dtype=torch.uint8, device=d)
        self.d_target_epoch = torch.zeros(n_agg, dtype=torch.int64, device=d)
        self.d_block_idx = torch.zeros(n_agg, dtype=torch.int32, device=d)
        pin = torch.cuda.is_available()
        self.h_ok = torch.zeros(n_agg, dtype=torch.uint8, pin_memory=pin)
        self.h_head = torch.zeros(1, dtype=torch.int32, pin_memory=pin)

    def set_committees(self, members, off):
        """members u32[n_sig] (committee order), off u32[n_agg+1]; signature j belongs to member j."""
        self.d_members = torch.as_tensor(members.astype("int32"), device=self.dev)
        self.d_off = torch.as_tensor(off.astype("int32"), device=self.dev)

    def process_epoch_dev(self, d_sigs, d_bits, d_msgs, d_target_epoch, d_block_idx, justified_idx=0, boost_idx=-1, boost_score=0):
        e = self.eng
        if hasattr(e, "epoch_dev"):
            e.epoch_dev(d_sigs, self.d_members, self.d_off, d_bits, d_msgs, d_target_epoch, d_block_idx, self.d_agg_sig, self.d_agg_status, self.d_ok)
        else:                                           # engines without the fused entry point (tests' stand-ins)
            e.aggregate_dev(d_sigs, self.d_off, self.d_agg_sig, self.d_agg_status)
            e.fast_aggregate_verify_dev(self.d_members, self.d_off, d_bits, d_msgs, self.d_agg_sig, self.d_ok)
            e.latest_messages_update_dev(self.d_members, self.d_off, d_bits, d_target_epoch, d_block_idx, self.d_ok)
        e.vote_weights_dev(self.d_votes)
        if self.pg is not None and torch.distributed.get_world_size(self.pg) > 1:
            torch.distributed.all_reduce(self.d_votes, group=self.pg)
        e.head_from_votes_dev(self.d_votes, self.d_head, justified_idx, boost_idx, boost_score)
        return self.d_ok, self.d_head

    def process_epoch_host(self, h_sigs, h_bits, h_msgs, h_target_epoch, h_block_idx, justified_idx=0, boost_idx=-1, boost_score=0):
        """Pinned host tensors in, (verdict bytes, head index) out on the host."""
        self.d_sigs.copy_(h_sigs, non_blocking=True)
        self.d_bits.copy_(h_bits, non_blocking=True)
        self.d_msgs.copy_(h_msgs, non_blocking=True)
        self.d_target_epoch.copy_(h_target_epoch, non_blocking=True)
        self.d_block_idx.copy_(h_block_idx, non_blocking=True)
        self.process_epoch_dev(self.d_sigs, self.d_bits, self.d_msgs, self.d_target_epoch, self.d_block_idx, justified_idx, boost_idx, boost_score)
        self.h_ok.copy_(self.d_ok, non_blocking=True)
        self.h_head.copy_(self.d_head, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self.h_ok, int(self.h_head[0])

    @property
    def h2d_bytes(self):
        return (self.d_sigs.numel() + self.d_bits.numel() + self.d_msgs.numel() + 8 * self.d_target_epoch.numel() + 4 * self.d_block_idx.numel())

    @property
    def d2h_bytes(self):
        return self.h_ok.numel() + 4
