"""EpochProcessor -- the batch (array-level) public API for one rank / one GPU.

One call processes a whole epoch of attestations:
  individual G2 signatures --bls.Aggregate per committee--> aggregate signatures
  --FastAggregateVerify (registry-indexed pubkey gather + hash-to-G2 + pairing)--> verdicts
  --update_latest_messages (accepted aggregates only)--> LMD table on the device
  --get_weight scatter--> per-block direct votes  [--NCCL all-reduce over ranks--]  --get_head--> head index.
(pos-evolution.md:722-754, :963-979, :1102-1116, :1435-1441.)

Two ways to drive it:
  * `process_epoch_dev` / `process_epoch_host`: synchronous in stream order, one epoch at a time;
  * `submit_dev` / `submit_host` + `drain`: software-pipelined over `depth` slots (default 3, at most B2_EPOCH_SLOTS = 16).
    The signature decompression of epoch k+1 (grid-filling, integer-pipe bound) overlaps with the latency-bound tails of
    the epochs before it (subgroup check, second Miller loop, final exponentiation, LMD update; each on its slot's own
    high-priority streams inside the library) and with their fork choice (on this object's fork-choice stream); in the
    host form the H2D copy of epoch k+1 overlaps too.  `submit_*` returns the ticket of the epoch submitted depth-1
    calls earlier.

Multi-GPU, two layouts:
  * `shard=None` (weak scaling): every rank owns its OWN validator set and epoch; the only exchange is the u64[n_blocks]
    all-reduce of direct vote weights (int64 two's-complement sum == u64 sum) between the vote scatter and get_head.
  * `shard=(rank, world)` (strong scaling -- BASELINE.json configs 4/5): ONE validator set and ONE epoch for the whole
    box.  Committees are disjoint within an epoch (pos-evolution.md:455, :472-474), so rank g aggregates and verifies the
    aggregates [g*n_agg/world, (g+1)*n_agg/world) -- its slots -- from ITS slice of the individual signatures, against a
    replicated registry.  Exchange per epoch: one all-gather of (aggregate signature 96 B + verdict 1 B) per aggregate,
    after which EVERY rank applies update_latest_messages for ALL accepted aggregates to its replica of the LMD table
    (12 B per set bit, microseconds); then get_head: rank g scatters the votes of validators [g*N/world, (g+1)*N/world)
    only, one u64[n_blocks] all-reduce, head on every rank.  No collective touches the BLS data path.
"""
import collections

import torch

from .engine import Engine

MAX_DEPTH = 16      # == B2_EPOCH_SLOTS (include/b200pos.h)


class _Ticket:
    """Result of one submitted epoch: verdict bytes + head index (+ the aggregate signatures), valid after .wait() and until
    `depth` further epochs have been submitted."""

    def __init__(self, slot, d_ok, d_head, d_agg_sig, event, h_ok=None, h_head=None, h_agg_sig=None):
        self.slot, self.d_ok, self.d_head, self.d_agg_sig, self.event = slot, d_ok, d_head, d_agg_sig, event
        self.h_ok, self.h_head, self.h_agg_sig = h_ok, h_head, h_agg_sig

    def wait(self):
        self.event.synchronize()
        if self.h_ok is not None:
            return self.h_ok.clone(), int(self.h_head[0])
        return self.d_ok, int(self.d_head[0])

    def aggregate_signatures(self):
        """bls.Aggregate's result for every committee of the epoch (uint8[n_agg, 96]); host tensor in the host form."""
        self.event.synchronize()
        return self.h_agg_sig.clone() if self.h_agg_sig is not None else self.d_agg_sig


class EpochProcessor:
    def __init__(self, engine: Engine, n_agg: int, n_sig: int, bits_stride: int, n_blocks: int, process_group=None, device=None, depth: int = 3,
                 shard=None, n_validators=None, tail_form: str = "auto"):
        """n_agg / n_sig: aggregates and individual signatures of the WHOLE epoch (in sharded mode this rank handles 1/world of
        them).  tail_form: "thread" (fewest instructions and smallest footprint; needs depth x step > the ~30-40 ms a tail lasts),
        "team" (three lanes per pairing: shortest critical path) or "auto" (= thread: measured best at 1, 2, 4 and 8 GPUs, see
        DESIGN.md section 6b).  Independently of it the last `team_last` epochs of a batch take the team form (drain_hint)."""
        self.eng = engine
        self.dev = device if device is not None else torch.device("cuda", engine.device)
        self.pg = process_group
        self.n_agg, self.n_sig, self.n_blocks, self.bits_stride = n_agg, n_sig, n_blocks, bits_stride
        d = self.dev
        cuda = self.dev.type == "cuda"
        assert 2 <= depth <= MAX_DEPTH, "depth must be 2..B2_EPOCH_SLOTS"
        self.depth = S = depth
        if shard is not None:
            self.rank, self.world = int(shard[0]), int(shard[1])
            assert 0 <= self.rank < self.world and n_agg % self.world == 0, "aggregates must divide evenly over the ranks"
            assert n_validators is not None, "sharded epochs need the size of the (replicated) registry"
        else:
            self.rank, self.world = 0, 1
        self.sharded = shard is not None and self.world > 1
        self.n_val = n_validators
        self.a0, self.a1 = self.rank * n_agg // self.world, (self.rank + 1) * n_agg // self.world
        self.n_loc = self.a1 - self.a0
        if self.sharded:
            self.v0, self.v1 = self.rank * n_validators // self.world, (self.rank + 1) * n_validators // self.world
        z = lambda shape, dt: torch.zeros(shape, dtype=dt, device=d)            # noqa: E731
        self.d_agg_sig = [z((n_agg, 96), torch.uint8) for _ in range(S)]
        self.d_agg_status = [z(self.n_loc, torch.int32) for _ in range(S)]
        self.d_ok = [z(n_agg, torch.uint8) for _ in range(S)]
        self.d_head = [z(1, torch.int32) for _ in range(S)]
        self.d_votes = z(n_blocks, torch.int64)
        if self.sharded:
            # per-rank outputs packed for ONE all-gather per epoch: n_loc aggregate signatures, then n_loc verdict bytes
            self.d_out_loc = [z(self.n_loc * 97, torch.uint8) for _ in range(S)]
            self.d_out_all = [z((self.world, self.n_loc * 97), torch.uint8) for _ in range(S)]
        # per-slot copies of everything the asynchronous tail / side streams read, and staging for the host entry points
        self.n_sig_loc = None        # set by set_committees
        self._alloc_sig_slots = lambda n: [z((n, 96), torch.uint8) for _ in range(S)]       # noqa: E731
        self.s_sigs = None
        self.s_bits = [z((n_agg, bits_stride), torch.uint8) for _ in range(S)]
        self.s_msgs = [z((n_agg, 32), torch.uint8) for _ in range(S)]
        self.s_epoch = [z(n_agg, torch.int64) for _ in range(S)]
        self.s_blk = [z(n_agg, torch.int32) for _ in range(S)]
        self.h_ok = [torch.zeros(n_agg, dtype=torch.uint8, pin_memory=cuda) for _ in range(S)]
        self.h_head = [torch.zeros(1, dtype=torch.int32, pin_memory=cuda) for _ in range(S)]
        self.h_agg_sig = [torch.zeros((n_agg, 96), dtype=torch.uint8, pin_memory=cuda) for _ in range(S)]
        self.copy_stream = torch.cuda.Stream(device=d) if cuda else None
        self.fc_stream = torch.cuda.Stream(device=d) if cuda else None
        self.ev_fc = [None] * S       # per slot: fork choice (+ D2H) of the epoch that last used the slot
        self.k = 0
        self._inflight = collections.deque()      # tickets of the submitted, not yet returned epochs
        self._team_form = False
        assert tail_form in ("auto", "thread", "team")
        # Measured on 2 and 8 GPUs (profiles/r2c_sweep_*.json): the thread-per-aggregate tail wins at every N as long as enough epochs are
        # in flight to cover its latency (N = 2: 18.5 ms thread / depth 4 against 21.7 ms team; N = 8: 6.32 against 6.80) -- the one-warp
        # team blocks with their 49 KB of shared memory displace more decompression blocks than their shorter latency is worth.
        self.always_team = tail_form == "team"
        # ... except at the end of a batch: the last `team_last` epochs take the team form (see drain_hint)
        self.team_last = 1

    def set_committees(self, members, off):
        """members u32[n_sig] (committee order), off u32[n_agg+1] for the WHOLE epoch; signature j belongs to member j."""
        off64 = off.astype("int64")
        self.d_members = torch.as_tensor(members.astype("int32"), device=self.dev)
        self.d_off = torch.as_tensor(off.astype("int32"), device=self.dev)
        self.m0, self.m1 = int(off64[self.a0]), int(off64[self.a1])
        self.n_sig_loc = self.m1 - self.m0
        # this rank's rows: members of its aggregates and offsets relative to its first member
        self.d_members_loc = self.d_members[self.m0:self.m1]
        self.d_off_loc = torch.as_tensor((off64[self.a0:self.a1 + 1] - off64[self.a0]).astype("int32"), device=self.dev)
        self.s_sigs = self._alloc_sig_slots(self.n_sig_loc)

    def local_signatures(self, sigs):
        """The slice of an epoch's individual signatures (uint8[n_sig, 96], committee order) this rank aggregates."""
        return sigs[self.m0:self.m1]

    # ------------------------------------------------------------------ the exchange + fork choice of one epoch
    def _rows(self, t):
        return t[self.a0:self.a1]

    def _exchange(self, slot, d_bits, d_target_epoch, d_block_idx):
        """Sharded epoch, after the local tail: all-gather (aggregate signature, verdict) of every rank's aggregates, then
        update_latest_messages for ALL accepted aggregates of the epoch on this rank's replica of the LMD table."""
        n = self.n_loc
        loc, allr = self.d_out_loc[slot], self.d_out_all[slot]
        if self.pg is None:                           # one rank of a sharded epoch on its own (bench.py --emulate-world): no peers to gather from
            allr[self.rank].copy_(loc)
        else:
            torch.distributed.all_gather_into_tensor(allr.view(-1), loc, group=self.pg)
        self.d_agg_sig[slot].view(self.world, n * 96).copy_(allr[:, :n * 96])
        self.d_ok[slot].view(self.world, n).copy_(allr[:, n * 96:])
        self.eng.latest_messages_update_dev(self.d_members, self.d_off, d_bits, d_target_epoch, d_block_idx, self.d_ok[slot])

    def _fork_choice(self, slot, justified_idx, boost_idx, boost_score):
        e = self.eng
        if self.sharded:
            e.vote_weights_range_dev(self.v0, self.v1, self.d_votes)
        else:
            e.vote_weights_dev(self.d_votes)
        if self.pg is not None and torch.distributed.get_world_size(self.pg) > 1:
            torch.distributed.all_reduce(self.d_votes, group=self.pg)
        e.head_from_votes_dev(self.d_votes, self.d_head[slot], justified_idx, boost_idx, boost_score)

    def enable_fused_get_head(self):
        """Sharded mode: set up get_head as ONE kernel per rank with the vote all-reduce fused in over NVLink peer memory
        (Engine.fc_exchange_setup: CUDA IPC handles exchanged over the process group).  Call on every rank, after tree_load."""
        assert self.sharded, "the fused multi-GPU get_head is for a validator set sharded over the ranks"
        ok, why = 1, ""
        try:
            self.eng.fc_exchange_setup(self.rank, self.world, self.pg)
        except Exception as e:                        # e.g. CUDA IPC not permitted between the processes of this box
            ok, why = 0, repr(e)
        flag = torch.tensor([ok], dtype=torch.int32, device=self.dev)
        torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN, group=self.pg)      # all ranks or none: the call is collective
        self._fused_head = bool(int(flag.item()))
        self.fused_head_error = why
        return self._fused_head

    def get_head(self, justified_idx=0, boost_idx=-1, boost_score=0, fused=None):
        """get_head over the current LMD table through the multi-rank path; returns the head index on the host.  COLLECTIVE.
        fused (default: whether enable_fused_get_head was called): one kernel per rank -- vote scatter of this rank's validators,
        64-bit reductions into every rank's accumulator over NVLink, flag exchange, tree phase, zero-copy result.  Otherwise the
        three-stage form: scatter kernel, NCCL all-reduce of u64[n_blocks], tree kernel, device-to-host copy."""
        if fused is None:
            fused = getattr(self, "_fused_head", False)
        if fused:
            return self.eng.get_head_multi(self.v0, self.v1, justified_idx, boost_idx, boost_score)
        self._fork_choice(0, justified_idx, boost_idx, boost_score)
        return int(self.d_head[0].item())

    # ------------------------------------------------------------------ synchronous form
    def _local_outputs(self, slot):
        """(aggregate-signature rows, verdict rows) the library writes for this rank's aggregates."""
        if self.sharded:
            n = self.n_loc
            return self.d_out_loc[slot][:n * 96], self.d_out_loc[slot][n * 96:]
        return self.d_agg_sig[slot], self.d_ok[slot]

    def process_epoch_dev(self, d_sigs, d_bits, d_msgs, d_target_epoch, d_block_idx, justified_idx=0, boost_idx=-1, boost_score=0):
        """d_sigs: this rank's signatures (all of them when not sharded); every other array covers the whole epoch."""
        e = self.eng
        o_sig, o_ok = self._local_outputs(0)
        if self.sharded:
            e.epoch_dev(d_sigs, self.d_members_loc, self.d_off_loc, self._rows(d_bits), self._rows(d_msgs), None, None, o_sig, self.d_agg_status[0], o_ok)
            self._exchange(0, d_bits, d_target_epoch, d_block_idx)
        elif hasattr(e, "epoch_dev"):
            e.epoch_dev(d_sigs, self.d_members, self.d_off, d_bits, d_msgs, d_target_epoch, d_block_idx, o_sig, self.d_agg_status[0], o_ok)
        else:                                           # engines without the fused entry point (tests' stand-ins)
            e.aggregate_dev(d_sigs, self.d_off, o_sig, self.d_agg_status[0])
            e.fast_aggregate_verify_dev(self.d_members, self.d_off, d_bits, d_msgs, o_sig, o_ok)
            e.latest_messages_update_dev(self.d_members, self.d_off, d_bits, d_target_epoch, d_block_idx, o_ok)
        self._fork_choice(0, justified_idx, boost_idx, boost_score)
        return self.d_ok[0], self.d_head[0]

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
        self.h_agg_sig[s].copy_(self.d_agg_sig[0], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self.h_ok[s], int(self.h_head[s][0])

    # ------------------------------------------------------------------ pipelined form
    def _submit(self, slot, d_sigs, d_bits, d_msgs, fc, host, last=False):
        e = self.eng
        team = last or self.always_team
        if team != self._team_form:                       # last epoch of a batch: nothing will overlap its tail, so take the short form
            e.epoch_set_pairing_form(team)
            self._team_form = team
        o_sig, o_ok = self._local_outputs(slot)
        if self.sharded:
            e.epoch_start_dev(slot, d_sigs, self.d_members_loc, self.d_off_loc, self._rows(d_bits), self._rows(d_msgs), self.d_agg_status[slot])
            e.epoch_tail_dev(slot, self.d_members_loc, self.d_off_loc, self._rows(d_bits), None, None, o_sig, self.d_agg_status[slot], o_ok)
        else:
            e.epoch_start_dev(slot, d_sigs, self.d_members, self.d_off, d_bits, d_msgs, self.d_agg_status[slot])
            e.epoch_tail_dev(slot, self.d_members, self.d_off, d_bits, self.s_epoch[slot], self.s_blk[slot], o_sig, self.d_agg_status[slot], o_ok)
        # fork choice of this epoch: on its own stream, enqueued BEFORE the next epoch's tail (whose LMD update waits for this scatter)
        with torch.cuda.stream(self.fc_stream):
            e.epoch_wait_dev(slot)
            if self.sharded:
                self._exchange(slot, d_bits, self.s_epoch[slot], self.s_blk[slot])
            self._fork_choice(slot, *fc)
            if host:
                self.h_ok[slot].copy_(self.d_ok[slot], non_blocking=True)
                self.h_head[slot].copy_(self.d_head[slot], non_blocking=True)
                self.h_agg_sig[slot].copy_(self.d_agg_sig[slot], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
        self.ev_fc[slot] = ev
        self._inflight.append(_Ticket(slot, self.d_ok[slot], self.d_head[slot], self.d_agg_sig[slot], ev,
                                      self.h_ok[slot] if host else None, self.h_head[slot] if host else None,
                                      self.h_agg_sig[slot] if host else None))
        self.k += 1
        return self._inflight.popleft() if len(self._inflight) >= self.depth else None

    def _claim_slot(self, stream):
        """`stream` is about to overwrite the slot's inputs/outputs: its previous epoch (tail and fork choice) must be done."""
        slot = self.k % self.depth
        self.eng.epoch_wait_dev(slot)                     # (on the current stream == `stream`)
        if self.ev_fc[slot] is not None:
            stream.wait_event(self.ev_fc[slot])
        return slot

    def drain_hint(self, remaining: int) -> bool:
        """True when an epoch with `remaining` epochs still to come after it should take the short-latency (team) tail: the last epochs
        of a batch cannot hide a thread-form tail (≈ 30 ms) behind decompressions that are no longer coming."""
        return remaining < self.team_last

    def submit_dev(self, d_sigs, d_bits, d_msgs, d_target_epoch, d_block_idx, justified_idx=0, boost_idx=-1, boost_score=0, last=False):
        """Enqueue one epoch (device-resident inputs, which must stay untouched until its ticket has been waited for; d_sigs =
        this rank's signatures, the other arrays cover the whole epoch).  Returns the ticket of the epoch submitted depth-1
        calls earlier (None while the pipeline fills).  `last=True` says no further epoch follows soon: its pairing tail is
        enqueued in the short-critical-path (team) form."""
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
        """host -> device bytes per epoch of this rank (submit_host): its signatures + the epoch's bits, messages, epochs, blocks"""
        return (self.s_sigs[0].numel() + self.s_bits[0].numel() + self.s_msgs[0].numel() + 8 * self.s_epoch[0].numel() + 4 * self.s_blk[0].numel())

    @property
    def d2h_bytes(self):
        """device -> host bytes per epoch (submit_host): verdicts + head index + the aggregate signatures"""
        return self.h_ok[0].numel() + 4 + self.h_agg_sig[0].numel()
