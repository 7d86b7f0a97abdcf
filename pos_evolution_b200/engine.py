"""Array-level host API over the C ABI: one Engine = one b2_ctx = one GPU.

Everything here takes/returns flat numpy arrays (host entry points) or torch CUDA tensors
(`*_dev`, asynchronous on torch's current stream).  The pyspec-signature layer (bls.py, spec.py)
and bench.py are built on it.  No CPU fallback: constructing an Engine without a B200 raises.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import B2Error


def _c(a, dtype):
    a = np.ascontiguousarray(a, dtype=dtype)
    return a


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


class Engine:
    def __init__(self, device: int = 0):
        self.lib = _lib.load()
        h = ctypes.c_void_p()
        rc = self.lib.b2_init(int(device), ctypes.byref(h))
        if rc != 0:
            raise B2Error(rc, "b2_init(device=%d) failed -- a B200 (sm_100) GPU is required; there is no CPU fallback" % device)
        self.h = h
        self.device = device
        self.n_validators = 0
        self.n_blocks = 0

    def close(self):
        if getattr(self, "h", None):
            self.lib.b2_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise B2Error(rc, self.lib.b2_last_error(self.h).decode())

    def sync(self):
        self._ck(self.lib.b2_sync(self.h))

    @property
    def launch_count(self) -> int:
        return int(self.lib.b2_launch_count(self.h))

    # ------------------------------------------------------------------ registry
    def registry_load(self, pubkeys48, effective_balance, flags=None):
        """pubkeys48: uint8[N,48]; effective_balance: uint64[N]; flags: uint8[N] (bit0 active, bit1 slashed).
        Returns uint8[N]: 1 where KeyValidate passed."""
        pk = _c(pubkeys48, np.uint8).reshape(-1, 48)
        n = pk.shape[0]
        eff = _c(effective_balance, np.uint64)
        fl = _c(flags if flags is not None else np.ones(n, dtype=np.uint8), np.uint8)
        assert eff.shape == (n,) and fl.shape == (n,)
        valid = np.zeros(n, dtype=np.uint8)
        self._ck(self.lib.b2_registry_load(self.h, _p(pk), _p(eff), _p(fl), n, _p(valid)))
        self.n_validators = n
        return valid

    def key_validate(self, pubkeys48):
        """bls.KeyValidate for explicit pubkeys (uint8[n,48]) -> uint8[n]; the registry on the device is not touched."""
        pk = _c(pubkeys48, np.uint8).reshape(-1, 48)
        out = np.zeros(pk.shape[0], dtype=np.uint8)
        self._ck(self.lib.b2_key_validate(self.h, _p(pk), pk.shape[0], _p(out)))
        return out

    def registry_update_balances(self, effective_balance, flags):
        eff, fl = _c(effective_balance, np.uint64), _c(flags, np.uint8)
        self._ck(self.lib.b2_registry_update_balances(self.h, _p(eff), _p(fl), eff.shape[0]))

    # ------------------------------------------------------------------ BLS batches
    @staticmethod
    def _batch(members, off, bits):
        members = _c(members, np.uint32)
        off = _c(off, np.uint32)
        bits = _c(bits, np.uint8)
        n_agg = off.shape[0] - 1
        assert bits.ndim == 2 and bits.shape[0] == n_agg
        return members, off, bits, n_agg, bits.shape[1]

    def g1_aggregate(self, members, off, bits):
        members, off, bits, n_agg, stride = self._batch(members, off, bits)
        out = np.zeros((n_agg, 48), dtype=np.uint8)
        status = np.zeros(n_agg, dtype=np.uint8)
        self._ck(self.lib.b2_g1_aggregate(self.h, _p(members), _p(off), _p(bits), stride, n_agg, _p(out), _p(status)))
        return out, status

    def aggregate(self, sigs96, seg_off):
        sigs = _c(sigs96, np.uint8).reshape(-1, 96)
        seg_off = _c(seg_off, np.uint32)
        n_seg = seg_off.shape[0] - 1
        assert int(seg_off[-1]) == sigs.shape[0]
        out = np.zeros((n_seg, 96), dtype=np.uint8)
        status = np.zeros(n_seg, dtype=np.int32)
        self._ck(self.lib.b2_aggregate(self.h, _p(sigs), _p(seg_off), n_seg, _p(out), _p(status)))
        return out, status

    def fast_aggregate_verify(self, members, off, bits, msgs32, sigs96):
        members, off, bits, n_agg, stride = self._batch(members, off, bits)
        msgs = _c(msgs32, np.uint8).reshape(n_agg, 32)
        sigs = _c(sigs96, np.uint8).reshape(n_agg, 96)
        ok = np.zeros(n_agg, dtype=np.uint8)
        self._ck(self.lib.b2_fast_aggregate_verify(self.h, _p(members), _p(off), _p(bits), stride, _p(msgs), _p(sigs), n_agg, _p(ok)))
        return ok

    def fast_aggregate_verify_pks(self, pks48, pk_off, msgs32, sigs96):
        pk_off = _c(pk_off, np.uint32)
        n_agg = pk_off.shape[0] - 1
        pks = _c(pks48, np.uint8).reshape(-1, 48)
        assert int(pk_off[-1]) == pks.shape[0]
        msgs = _c(msgs32, np.uint8).reshape(n_agg, 32)
        sigs = _c(sigs96, np.uint8).reshape(n_agg, 96)
        ok = np.zeros(n_agg, dtype=np.uint8)
        self._ck(self.lib.b2_fast_aggregate_verify_pks(self.h, _p(pks), _p(pk_off), _p(msgs), _p(sigs), n_agg, _p(ok)))
        return ok

    @staticmethod
    def _sk_limbs(sks):
        out = np.zeros((len(sks), 8), dtype=np.uint32)
        for i, k in enumerate(sks):
            k = int(k)
            for j in range(8):
                out[i, j] = (k >> (32 * j)) & 0xFFFFFFFF
        return out

    def sk_to_pk(self, sks):
        sk8 = sks if isinstance(sks, np.ndarray) else self._sk_limbs(sks)
        sk8 = _c(sk8, np.uint32).reshape(-1, 8)
        out = np.zeros((sk8.shape[0], 48), dtype=np.uint8)
        self._ck(self.lib.b2_sk_to_pk(self.h, _p(sk8), sk8.shape[0], _p(out)))
        return out

    def sign(self, sks, msg_idx, msgs32):
        sk8 = sks if isinstance(sks, np.ndarray) else self._sk_limbs(sks)
        sk8 = _c(sk8, np.uint32).reshape(-1, 8)
        msg_idx = _c(msg_idx, np.uint32)
        msgs = _c(msgs32, np.uint8).reshape(-1, 32)
        out = np.zeros((sk8.shape[0], 96), dtype=np.uint8)
        self._ck(self.lib.b2_sign(self.h, _p(sk8), _p(msg_idx), sk8.shape[0], _p(msgs), msgs.shape[0], _p(out)))
        return out

    def hash_to_g2(self, msgs32):
        msgs = _c(msgs32, np.uint8).reshape(-1, 32)
        out = np.zeros((msgs.shape[0], 96), dtype=np.uint8)
        self._ck(self.lib.b2_hash_to_g2(self.h, _p(msgs), msgs.shape[0], _p(out)))
        return out

    def sha256_batch(self, msgs, msg_len: int):
        """SHA-256 of n fixed-length messages (uint8[n, msg_len]) on the GPU -> uint8[n, 32]."""
        m = _c(msgs, np.uint8).reshape(-1, msg_len) if msg_len else np.zeros((int(msgs), 0), dtype=np.uint8)
        n = m.shape[0]
        out = np.zeros((n, 32), dtype=np.uint8)
        self._ck(self.lib.b2_sha256_batch(self.h, _p(m) if msg_len else None, msg_len, n, _p(out)))
        return out

    def attestations_decode(self, wire: bytes, woff, bits_stride: int, max_bits: int):
        """SSZ wire decode of n Attestations (pos-evolution.md:714-717) -> (bits u8[n][stride], bit_len u32[n], data128 u8[n][128],
        sig96 u8[n][96], status i32[n])."""
        woff = _c(woff, np.uint32)
        n = woff.shape[0] - 1
        w = np.frombuffer(bytes(wire), dtype=np.uint8) if len(wire) else np.zeros(1, dtype=np.uint8)
        assert int(woff[-1]) == len(wire)
        bits = np.zeros((n, bits_stride), dtype=np.uint8)
        blen = np.zeros(n, dtype=np.uint32)
        data = np.zeros((n, 128), dtype=np.uint8)
        sig = np.zeros((n, 96), dtype=np.uint8)
        st = np.zeros(n, dtype=np.int32)
        self._ck(self.lib.b2_attestations_decode(self.h, _p(w), _p(woff), n, int(bits_stride), int(max_bits), _p(bits), _p(blen), _p(data), _p(sig), _p(st)))
        return bits, blen, data, sig, st

    def signing_roots(self, data128, domains32):
        """compute_signing_root for n attestations: data128 uint8[n,128] (SSZ-serialised AttestationData), domains32 uint8[32] or uint8[n,32]."""
        d = _c(data128, np.uint8).reshape(-1, 128)
        dom = _c(domains32, np.uint8).reshape(-1, 32)
        n = d.shape[0]
        assert dom.shape[0] in (1, n)
        out = np.zeros((n, 32), dtype=np.uint8)
        self._ck(self.lib.b2_signing_roots(self.h, _p(d), _p(dom), 1 if dom.shape[0] == n and n > 1 else 0, n, _p(out)))
        return out

    def shuffle_committees(self, seed32: bytes, n_active: int, rounds: int, active=None):
        """members[i] = active[compute_shuffled_index(i, n_active, seed)] for the whole active set, on the GPU."""
        seed = np.frombuffer(bytes(seed32), dtype=np.uint8)
        assert seed.shape == (32,)
        act = _c(active, np.uint32) if active is not None else None
        out = np.zeros(n_active, dtype=np.uint32)
        self._ck(self.lib.b2_shuffle_committees(self.h, _p(seed), _p(act), n_active, rounds, _p(out)))
        return out

    # ------------------------------------------------------------------ fork choice
    def latest_messages_reset(self):
        self._ck(self.lib.b2_latest_messages_reset(self.h))

    def latest_messages_load(self, epoch, block_idx, has_msg, equivocating):
        e, b = _c(epoch, np.uint64), _c(block_idx, np.uint32)
        h, q = _c(has_msg, np.uint8), _c(equivocating, np.uint8)
        self._ck(self.lib.b2_latest_messages_load(self.h, _p(e), _p(b), _p(h), _p(q), e.shape[0]))

    def latest_messages_read(self):
        n = self.n_validators
        e, b, h = np.zeros(n, dtype=np.uint64), np.zeros(n, dtype=np.uint32), np.zeros(n, dtype=np.uint8)
        self._ck(self.lib.b2_latest_messages_read(self.h, _p(e), _p(b), _p(h), n))
        return e, b, h

    def latest_messages_update(self, members, off, bits, target_epoch, block_idx, accept=None):
        members, off, bits, n_agg, stride = self._batch(members, off, bits)
        te, bi = _c(target_epoch, np.uint64), _c(block_idx, np.uint32)
        acc = _c(accept, np.uint8) if accept is not None else None
        self._ck(self.lib.b2_latest_messages_update(self.h, _p(members), _p(off), _p(bits), stride, _p(te), _p(bi), _p(acc), n_agg))

    def participation_load(self, which: int, participation):
        p = _c(participation, np.uint8)
        self._ck(self.lib.b2_participation_load(self.h, int(which), _p(p), p.shape[0]))

    def participation_read(self, which: int):
        out = np.zeros(self.n_validators, dtype=np.uint8)
        self._ck(self.lib.b2_participation_read(self.h, int(which), _p(out), self.n_validators))
        return out

    def participation_update(self, which, members, off, bits, flag_mask, accept, increment, base_reward_per_increment):
        """process_attestation's flag loop (:745-749) for a batch; returns the per-attestation proposer_reward_numerator (u64)."""
        members, off, bits, n_agg, stride = self._batch(members, off, bits)
        fm = _c(flag_mask, np.uint8)
        acc = _c(accept, np.uint8) if accept is not None else None
        num = np.zeros(n_agg, dtype=np.uint64)
        self._ck(self.lib.b2_participation_update(self.h, int(which), _p(members), _p(off), _p(bits), stride, _p(fm), _p(acc), n_agg,
                                                  int(increment), int(base_reward_per_increment), _p(num)))
        return num

    def ffg_balances(self, flag_index: int):
        """(total_active, current_epoch_flag_balance, previous_epoch_flag_balance, total_active_unslashed) -- pos-evolution.md:793-803."""
        out = np.zeros(4, dtype=np.uint64)
        self._ck(self.lib.b2_ffg_balances(self.h, int(flag_index), _p(out)))
        return tuple(int(x) for x in out)

    def set_fork_choice_params(self, min_vote_epoch: int = 0, exclude_slashed: bool = False):
        self._ck(self.lib.b2_set_fork_choice_params(self.h, int(min_vote_epoch), 1 if exclude_slashed else 0))

    def set_verify_mode(self, rlc: bool, seed32: bytes = None):
        """FastAggregateVerify per aggregate (default) or in random-linear-combination batches of 32 with per-aggregate fallback
        (same verdict vector, about half the pairing work).  seed32: the verifier's secret randomness (os.urandom when None)."""
        import os
        seed = np.frombuffer(bytes(seed32) if seed32 is not None else os.urandom(32), dtype=np.uint8)
        assert seed.shape == (32,)
        self._ck(self.lib.b2_set_verify_mode(self.h, 1 if rlc else 0, _p(seed) if rlc else None))

    def on_attester_slashing(self, indices_1, indices_2):
        """Store.equivocating_indices |= set(indices_1) & set(indices_2)   (pos-evolution.md:1459-1461); both lists sorted."""
        a, b = _c(indices_1, np.uint32), _c(indices_2, np.uint32)
        self._ck(self.lib.b2_on_attester_slashing(self.h, _p(a), a.shape[0], _p(b), b.shape[0]))

    def tree_load(self, parent, slot, roots32, leaf_viable):
        parent, slot = _c(parent, np.uint32), _c(slot, np.uint64)
        roots = _c(roots32, np.uint8).reshape(-1, 32)
        viable = _c(leaf_viable, np.uint8)
        n = parent.shape[0]
        assert slot.shape == (n,) and roots.shape[0] == n and viable.shape == (n,)
        self._ck(self.lib.b2_tree_load(self.h, _p(parent), _p(slot), _p(roots), _p(viable), n))
        self.n_blocks = n

    def get_weights(self, boost_idx: int = -1, boost_score: int = 0):
        w = np.zeros(self.n_blocks, dtype=np.uint64)
        self._ck(self.lib.b2_get_weights(self.h, int(boost_idx), int(boost_score), _p(w)))
        return w

    def get_head(self, justified_idx: int = 0, boost_idx: int = -1, boost_score: int = 0) -> int:
        out = ctypes.c_uint32(0)
        self._ck(self.lib.b2_get_head(self.h, int(justified_idx), int(boost_idx), int(boost_score), ctypes.byref(out)))
        return int(out.value)

    # ------------------------------------------------------------------ multi-GPU get_head over NVLink peer memory
    def fc_exchange_setup(self, rank: int, world: int, process_group=None):
        """Map the vote-exchange blocks of all ranks of the box into this context (CUDA IPC handles all-gathered over `process_group`;
        world == 1 needs no group).  Call after tree_load, on every rank."""
        h = np.zeros(64, dtype=np.uint8)
        self._ck(self.lib.b2_fc_exchange_export(self.h, _p(h)))
        if world > 1:
            import torch.distributed as dist
            gathered = [None] * world
            dist.all_gather_object(gathered, h.tobytes(), group=process_group)
            allh = np.frombuffer(b"".join(gathered), dtype=np.uint8).copy()
        else:
            allh = h
        self._ck(self.lib.b2_fc_exchange_open(self.h, int(rank), int(world), _p(allh)))

    def get_head_multi(self, v_begin: int, v_end: int, justified_idx: int = 0, boost_idx: int = -1, boost_score: int = 0) -> int:
        """COLLECTIVE get_head: this rank scatters validators [v_begin, v_end); the all-reduce runs inside the kernel over NVLink."""
        out = ctypes.c_uint32(0)
        self._ck(self.lib.b2_get_head_multi(self.h, int(v_begin), int(v_end), int(justified_idx), int(boost_idx), int(boost_score), ctypes.byref(out)))
        return int(out.value)

    def debug_head_clocks(self):
        """clock64 stamps of the phases of the last get_head (profiling aid) -> uint64[32]"""
        out = np.zeros(32, dtype=np.uint64)
        self._ck(self.lib.b2_debug_head_clocks(self.h, _p(out)))
        return out

    # ------------------------------------------------------------------ device-pointer entry points (torch CUDA tensors)
    @staticmethod
    def _stream():
        import torch
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def aggregate_dev(self, d_sigs, d_seg_off, d_out96, d_status):
        n_seg = d_seg_off.numel() - 1
        n_sig = d_sigs.numel() // 96
        self._ck(self.lib.b2_aggregate_dev(self.h, d_sigs.data_ptr(), d_seg_off.data_ptr(), n_seg, n_sig, d_out96.data_ptr(),
                                           d_status.data_ptr(), self._stream()))

    def fast_aggregate_verify_dev(self, d_members, d_off, d_bits, d_msgs, d_sigs, d_ok):
        n_agg = d_off.numel() - 1
        stride = d_bits.shape[1]
        self._ck(self.lib.b2_fast_aggregate_verify_dev(self.h, d_members.data_ptr(), d_off.data_ptr(), d_bits.data_ptr(), stride,
                                                       d_msgs.data_ptr(), d_sigs.data_ptr(), n_agg, d_ok.data_ptr(), self._stream()))

    def latest_messages_update_dev(self, d_members, d_off, d_bits, d_target_epoch, d_block_idx, d_accept):
        n_agg = d_off.numel() - 1
        self._ck(self.lib.b2_latest_messages_update_dev(self.h, d_members.data_ptr(), d_off.data_ptr(), d_bits.data_ptr(), d_bits.shape[1],
                                                        d_target_epoch.data_ptr(), d_block_idx.data_ptr(),
                                                        d_accept.data_ptr() if d_accept is not None else None, n_agg, self._stream()))

    def epoch_dev(self, d_sigs, d_members, d_off, d_bits, d_msgs, d_target_epoch, d_block_idx, d_agg_sig, d_agg_status, d_ok):
        n_agg = d_off.numel() - 1
        self._ck(self.lib.b2_epoch_dev(self.h, d_sigs.data_ptr(), d_members.data_ptr(), d_off.data_ptr(), d_bits.data_ptr(), d_bits.shape[1],
                                       d_msgs.data_ptr(), d_target_epoch.data_ptr() if d_target_epoch is not None else None,
                                       d_block_idx.data_ptr() if d_block_idx is not None else None, n_agg, d_sigs.numel() // 96,
                                       d_agg_sig.data_ptr(), d_agg_status.data_ptr(), d_ok.data_ptr(), self._stream()))

    def epoch_start_dev(self, slot, d_sigs, d_members, d_off, d_bits, d_msgs, d_agg_status):
        n_agg = d_off.numel() - 1
        self._ck(self.lib.b2_epoch_start_dev(self.h, int(slot), d_sigs.data_ptr(), d_members.data_ptr(), d_off.data_ptr(), d_bits.data_ptr(),
                                             d_bits.shape[1], d_msgs.data_ptr(), n_agg, d_sigs.numel() // 96, d_agg_status.data_ptr(), self._stream()))

    def epoch_tail_dev(self, slot, d_members, d_off, d_bits, d_target_epoch, d_block_idx, d_agg_sig, d_agg_status, d_ok):
        n_agg = d_off.numel() - 1
        """d_target_epoch / d_block_idx None: no LMD update in the tail (sharded epoch: the caller applies it after the verdict exchange)."""
        self._ck(self.lib.b2_epoch_tail_dev(self.h, int(slot), d_members.data_ptr(), d_off.data_ptr(), d_bits.data_ptr(), d_bits.shape[1],
                                            d_target_epoch.data_ptr() if d_target_epoch is not None else None,
                                            d_block_idx.data_ptr() if d_block_idx is not None else None, n_agg, d_agg_sig.data_ptr(),
                                            d_agg_status.data_ptr(), d_ok.data_ptr()))

    def epoch_set_pairing_form(self, team: bool):
        self._ck(self.lib.b2_epoch_set_pairing_form(self.h, 1 if team else 0))

    def epoch_wait_dev(self, slot):
        self._ck(self.lib.b2_epoch_wait_dev(self.h, int(slot), self._stream()))

    def vote_weights_dev(self, d_votes):
        self._ck(self.lib.b2_vote_weights_dev(self.h, d_votes.data_ptr(), self._stream()))

    def vote_weights_range_dev(self, v_begin, v_end, d_votes):
        """direct votes of the validators [v_begin, v_end) only (this rank's shard of one validator set)"""
        self._ck(self.lib.b2_vote_weights_range_dev(self.h, int(v_begin), int(v_end), d_votes.data_ptr(), self._stream()))

    def gather_probe_dev(self, d_members, d_off, d_bits, d_checksum, tma: bool = True):
        """gather stage of K2 alone: d_checksum[a] = XOR of the pubkey-record words aggregate a selects (TMA-staged or plain-load form)"""
        n_agg = d_off.numel() - 1
        self._ck(self.lib.b2_gather_probe_dev(self.h, d_members.data_ptr(), d_off.data_ptr(), d_bits.data_ptr(), d_bits.shape[1], n_agg,
                                              1 if tma else 0, d_checksum.data_ptr(), self._stream()))

    def guard_flags(self) -> int:
        """synchronise; return and clear the device-side input guard word (B2_GUARD_* bits) of the *_dev entry points"""
        out = ctypes.c_uint32(0)
        self._ck(self.lib.b2_guard_flags(self.h, ctypes.byref(out)))
        return int(out.value)

    def head_from_votes_dev(self, d_votes, d_head, justified_idx=0, boost_idx=-1, boost_score=0, d_weight=None):
        self._ck(self.lib.b2_head_from_votes_dev(self.h, d_votes.data_ptr(), int(justified_idx), int(boost_idx), int(boost_score),
                                                 d_weight.data_ptr() if d_weight is not None else None, d_head.data_ptr(), self._stream()))
