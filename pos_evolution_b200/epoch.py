"""EpochProcessor -- the batch (array-level) public API for one rank / one GPU.

One call processes a whole epoch of attestations for the validators this rank owns:
  individual G2 signatures --bls.Aggregate per committee--> aggregate signatures
  --FastAggregateVerify (registry-indexed pubkey gather + hash-to-G2 + pairing)--> verdicts
  --update_latest_messages (accepted aggregates only)--> LMD table on the device
  --get_weight scatter--> per-block direct votes  [--NCCL all-reduce over ranks--]  --get_head--> head index.
(pos-evolution.md:722-754, :963-979, :1102-1116, :1435-1441.)

Two ways to drive it:
  * `process_epoch_dev` / `process_epoch_host`: synchronous in stream order, one epoch at a time;
  * `submit_dev` / `submit_host` + `drain`: software-pipelined over `depth` slots (default 3).  The signature
    decompression of epoch k+1 (grid-filling, integer-pipe bound) overlaps with the latency-bound tails of epochs k and
    k-1 (subgroup check, second Miller loop, final exponentiation, LMD update; each on its slot's own high-priority
    stream inside the library) and with their fork choice (on this object's fork-choice stream); in the host form the
    H2D copy of epoch k+1 overlaps too.  Under that contention one tail takes longer than one decompression, which is
    why two of them are kept in flight.  `submit_*` returns the ticket of the epoch submitted depth-1 calls earlier.

Multi-GPU: validators (and with them committees / aggregates) are sharded across ranks with no data-path exchange
until the vote weights: u64[n_blocks] direct votes are summed with one torch.distributed all_reduce (NCCL over NVLink;
int64 two's-complement sum == u64 sum), after which every rank finishes get_head on its replica of the block tree.
"""
import collections

import torch

from .engine import Engine


class _Ticket:
    """Result of one submitted epoch: verdict bytes + head index (+ the aggregate signatures on the device), valid after
    .wait() and until `depth` further epochs have been submitted."""

    def __init__(self, slot, d_ok, d_head, d_agg_sig, event, h_ok=None, h_head=None):
        self.slot, self.d_ok, self.d_head, self.d_agg_sig, self.event, self.h_ok, self.h_head = slot, d_ok, d_head, d_agg_sig, event, h_ok, h_head

    def wait(self):
        self.event.synchronize()
        if self.h_ok is not None:
            return self.h_ok.clone(), int(self.h_head[0])
        return self.d_ok, int(self.d_head[0])


class EpochProcessor:
    def __init__(self, engine: Engine, n_agg: int, n_sig: int, bits_stride: int, n_blocks: int, process_group=None, device=None, depth: int = 3):
        self.eng = engine
        self.dev = device if device is not None else torch.device("cuda", engine.device)
        self.pg = process_group
        self.n_agg, self.n_sig, self.n_blocks = n_agg, n_sig, n_blocks
        d = self.dev
        cuda = self.dev.type == "cuda"
        assert 2 <= depth <= 4, "depth must be 2..B2_EPOCH_SLOTS"
        self.depth = S = depth
        z = lambda shape, dt: torch.zeros(shape, dtype=dt, device=d)            # noqa: E731
        self.d_agg_sig = [z((n_agg, 96), torch.uint8) for _ in range(S)]
        self.d_agg_status = [z(n_agg, torch.int32) for _ in range(S)]
        self.d_ok = [z(n_agg, torch.uint8) for _ in range(S)]
        self.d_head = [z(1, torch.int32) for _ in range(S)]
        self.d_votes = z(n_blocks, torch.int64)
        # per-slot copies of everything the asynchronous tail / side streams read, and staging for the host entry points
        self.s_sigs = [z((n_sig, 96), torch.uint8) for _ in range(S)]
        self.s_bits = [z((n_agg, bits_stride), torch.uint8) for _ in range(S)]
        self.s_msgs = [z((n_agg, 32), torch.uint8) for _ in range(S)]
        self.s_epoch = [z(n_agg, torch.int64) for _ in range(S)]
        self.s_blk = [z(n_agg, torch.int32) for _ in range(S)]
        self.h_ok = [torch.zeros(n_agg, dtype=torch.uint8, pin_memory=cuda) for _ in range(S)]
        self.h_head = [torch.zeros(1, dtype=torch.int32, pin_memory=cuda) for _ in range(S)]
        self.copy_stream = torch.cuda.Stream(device=d) if cuda else None
        self.fc_stream = torch.cuda.Stream(device=d) if cuda else None
        self.ev_fc = [None] * S       # per slot: fork choice (+ D2H) of the epoch that last used the slot
        self.k = 0
        self._inflight = collections.deque()      # tickets of the submitted, not yet returned epochs
        self._team_form = False

    def set_committees(self, members, off):
        """members u32[n_sig] (committee order), off u32[n_agg+1]; signature j belongs to member j."""
        self.d_members = torch.as_tensor(members.astype("int32"), device=self.dev)
        self.d_off = torch.as_tensor(off.astype("int32"), device=self.dev)

    # ------------------------------------------------------------------ synchronous form
    def process_epoch_dev(self, d_sigs, d_bits, d_msgs, d_target_epoch, d_block_idx, justified_idx=0, boost_idx=-1, boost_score=0):
        e = self.eng
        if hasattr(e, "epoch_dev"):
            e.epoch_dev(d_sigs, self.d_members, self.d_off, d_bits, d_msgs, d_target_epoch, d_block_idx, self.d_agg_sig[0], self.d_agg_status[0], self.d_ok[0])
        else:                                           # engines without the fused entry point (tests' stand-ins)
            e.aggregate_dev(d_sigs, self.d_off, self.d_agg_sig[0], self.d_agg_status[0])
            e.fast_aggregate_verify_dev(self.d_members, self.d_off, d_bits, d_msgs, self.d_agg_sig[0], self.d_ok[0])
            e.latest_messages_update_dev(self.d_members, self.d_off, d_bits, d_target_epoch, d_block_idx, self.d_ok[0])
        self._fork_choice(0, justified_idx, boost_idx, boost_score)
        return self.d_ok[0], self.d_head[0]

    def _fork_choice(self, slot, justified_idx, boost_idx, boost_score):
        e = self.eng
        e.vote_weights_dev(self.d_votes)
        if self.pg is not None and torch.distributed.get_world_size(self.pg) > 1:
            torch.distributed.all_reduce(self.d_votes, group=self.pg)
        e.head_from_votes_dev(self.d_votes, self.d_head[slot], justified_idx, boost_idx, boost_score)

    def process_epoch_host(self, h_sigs, h_bits, h_msgs, h_target_epoch, h_block_idx, justified_idx=0, boost_idx=-1, boost_score=0):
        """Pinned host tensors in, (verdict bytes, head index) out on the host; returns when the epoch is done."""
        s = 0
        self.s_sigs[s].copy_(h_sigs, non_blocking=True)
        self.s_bits[s].copy_(h_bits, non_blocking=True)
        self.s_msgs[s].copy_(h_msgs, non_blocking=True)
        self.s_epoch[s].copy_(h_target_epoch, non_blocking=True)
        self.s_blk[s].copy_(h_block_idx, non_blocking=True)
        self.process_epoch_dev(self.s_sigs[s], self.s_bits[s], self.s_msgs[s], self.s_epoch[s], self.s_blk[s], justified_idx, boost_idx, boost_score)
        self.h_ok[s].copy_(self.d_ok[0], non_blocking=True)
        self.h_head[s].copy_(self.d_head[0], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self.h_ok[s], int(self.h_head[s][0])

    # ------------------------------------------------------------------ pipelined form
    def _submit(self, slot, d_sigs, d_bits, d_msgs, fc, host, last=False):
        e = self.eng
        if last != self._team_form:                       # last epoch of a batch: nothing will overlap its tail, so take the short form
            e.epoch_set_pairing_form(last)
            self._team_form = last
        e.epoch_start_dev(slot, d_sigs, self.d_members, self.d_off, d_bits, d_msgs, self.d_agg_status[slot])
        e.epoch_tail_dev(slot, self.d_members, self.d_off, d_bits, self.s_epoch[slot], self.s_blk[slot], self.d_agg_sig[slot],
                         self.d_agg_status[slot], self.d_ok[slot])
        # fork choice of this epoch: on its own stream, enqueued BEFORE the next epoch's tail (whose LMD update waits for this scatter)
        with torch.cuda.stream(self.fc_stream):
            e.epoch_wait_dev(slot)
            self._fork_choice(slot, *fc)
            if host:
                self.h_ok[slot].copy_(self.d_ok[slot], non_blocking=True)
                self.h_head[slot].copy_(self.d_head[slot], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
        self.ev_fc[slot] = ev
        self._inflight.append(_Ticket(slot, self.d_ok[slot], self.d_head[slot], self.d_agg_sig[slot], ev,
                                      self.h_ok[slot] if host else None, self.h_head[slot] if host else None))
        self.k += 1
        return self._inflight.popleft() if len(self._inflight) >= self.depth else None

    def _claim_slot(self, stream):
        """`stream` is about to overwrite the slot's inputs/outputs: its previous epoch (tail and fork choice) must be done."""
        slot = self.k % self.depth
        self.eng.epoch_wait_dev(slot)                     # (on the current stream == `stream`)
        if self.ev_fc[slot] is not None:
            stream.wait_event(self.ev_fc[slot])
        return slot

    def submit_dev(self, d_sigs, d_bits, d_msgs, d_target_epoch, d_block_idx, justified_idx=0, boost_idx=-1, boost_score=0, last=False):
        """Enqueue one epoch (device-resident inputs, which must stay untouched until its ticket has been waited for).
        Returns the ticket of the epoch submitted depth-1 calls earlier (None while the pipeline fills).
        `last=True` says no further epoch follows soon: its pairing tail is enqueued in the short-critical-path (team) form."""
        slot = self._claim_slot(torch.cuda.current_stream())
        self.s_epoch[slot].copy_(d_target_epoch)
        self.s_blk[slot].copy_(d_block_idx)
        return self._submit(slot, d_sigs, d_bits, d_msgs, (justified_idx, boost_idx, boost_score), False, last)

    def submit_host(self, h_sigs, h_bits, h_msgs, h_target_epoch, h_block_idx, justified_idx=0, boost_idx=-1, boost_score=0, last=False):
        """Pinned host tensors in (untouched until the returned-later ticket has been waited for).  The H2D copies go on a
        separate stream so that they overlap with the epochs in flight."""
        cur = torch.cuda.current_stream()
        with torch.cuda.stream(self.copy_stream):
            slot = self._claim_slot(self.copy_stream)
            self.s_sigs[slot].copy_(h_sigs, non_blocking=True)
            self.s_bits[slot].copy_(h_bits, non_blocking=True)
            self.s_msgs[slot].copy_(h_msgs, non_blocking=True)
            self.s_epoch[slot].copy_(h_target_epoch, non_blocking=True)
            self.s_blk[slot].copy_(h_block_idx, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
        cur.wait_event(ev)
        return self._submit(slot, self.s_sigs[slot], self.s_bits[slot], self.s_msgs[slot], (justified_idx, boost_idx, boost_score), True, last)

    def drain(self):
        """Tickets of the epochs still in flight, oldest first; the current stream waits for all of them."""
        out = list(self._inflight)
        self._inflight.clear()
        for t in out:
            torch.cuda.current_stream().wait_event(t.event)
        return out

    @property
    def h2d_bytes(self):
        return (self.s_sigs[0].numel() + self.s_bits[0].numel() + self.s_msgs[0].numel() + 8 * self.s_epoch[0].numel() + 4 * self.s_blk[0].numel())

    @property
    def d2h_bytes(self):
        return self.h_ok[0].numel() + 4
