"""pyspec-signature host layer over the B200 engine.

Keeps the names, argument meaning and error behaviour (AssertionError = invalid) of the functions
quoted in /root/reference/pos-evolution.md, line numbers cited per function, so that it can stand
in for the executable consensus spec on the path
    process_attestation (:722) -> is_valid_indexed_attestation -> bls.FastAggregateVerify
    on_attestation (:963/:1423) -> update_latest_messages (:1435) -> get_head (:1102).
Python here only marshals: committee shuffling (SHA-256 via hashlib, vectorised numpy -- the GPU
version is SURVEY.md section 8(f)-1), SSZ signing roots, participation-flag bookkeeping.  Signature
aggregation/verification and the fork-choice weights/head run on the GPU through engine.Engine.
This module never imports oracle/.
"""
import hashlib
from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Set

import numpy as np

from .engine import Engine

FAR_FUTURE_EPOCH = 2**64 - 1
GENESIS_EPOCH = 0
ZERO32 = bytes(32)
DOMAIN_BEACON_PROPOSER = bytes.fromhex("00000000")
DOMAIN_BEACON_ATTESTER = bytes.fromhex("01000000")
TIMELY_SOURCE_FLAG_INDEX, TIMELY_TARGET_FLAG_INDEX, TIMELY_HEAD_FLAG_INDEX = 0, 1, 2
PARTICIPATION_FLAG_WEIGHTS = [14, 26, 14]
PROPOSER_WEIGHT = 8
WEIGHT_DENOMINATOR = 64
BASE_REWARD_FACTOR = 64
JUSTIFICATION_BITS_LENGTH = 4


@dataclass(frozen=True)
class Preset:
    name: str
    SLOTS_PER_EPOCH: int
    MAX_COMMITTEES_PER_SLOT: int
    TARGET_COMMITTEE_SIZE: int
    SHUFFLE_ROUND_COUNT: int
    EPOCHS_PER_HISTORICAL_VECTOR: int
    SLOTS_PER_HISTORICAL_ROOT: int
    MAX_VALIDATORS_PER_COMMITTEE: int = 2048
    MIN_ATTESTATION_INCLUSION_DELAY: int = 1
    MIN_SEED_LOOKAHEAD: int = 1
    MAX_EFFECTIVE_BALANCE: int = 32 * 10**9
    EFFECTIVE_BALANCE_INCREMENT: int = 10**9
    PROPOSER_SCORE_BOOST: int = 40


MAINNET = Preset("mainnet", 32, 64, 128, 90, 65536, 8192)
MINIMAL = Preset("minimal", 8, 4, 4, 10, 64, 64)


# ----------------------------------------------------------------------------- containers
@dataclass
class Validator:                       # pos-evolution.md:36-45
    pubkey: bytes
    effective_balance: int
    slashed: bool = False
    activation_epoch: int = 0
    exit_epoch: int = FAR_FUTURE_EPOCH


@dataclass(frozen=True)
class Checkpoint:                      # :219-221
    epoch: int = 0
    root: bytes = ZERO32


@dataclass(frozen=True)
class AttestationData:                 # :689-697
    slot: int
    index: int
    beacon_block_root: bytes
    source: Checkpoint
    target: Checkpoint


@dataclass
class Attestation:                     # :714-717
    aggregation_bits: List[bool]
    data: AttestationData
    signature: bytes


@dataclass
class IndexedAttestation:
    attesting_indices: List[int]
    data: AttestationData
    signature: bytes


@dataclass
class AttesterSlashing:                # pos-evolution.md:1160-1162
    attestation_1: "IndexedAttestation"
    attestation_2: "IndexedAttestation"


@dataclass(frozen=True)
class LatestMessage:                   # :287-289
    epoch: int
    root: bytes


@dataclass
class Fork:
    previous_version: bytes = bytes(4)
    current_version: bytes = bytes(4)
    epoch: int = 0


@dataclass
class BeaconBlock:                     # :671-676 (fields fork choice reads)
    slot: int
    parent_root: bytes


@dataclass
class BeaconState:                     # :338-374 (fields the path reads/writes)
    slot: int
    fork: Fork
    genesis_validators_root: bytes
    validators: List[Validator]
    balances: List[int]
    randao_mixes: List[bytes]
    block_roots: List[bytes]
    previous_epoch_participation: List[int]
    current_epoch_participation: List[int]
    previous_justified_checkpoint: Checkpoint = Checkpoint()
    current_justified_checkpoint: Checkpoint = Checkpoint()
    finalized_checkpoint: Checkpoint = Checkpoint()
    justification_bits: List[int] = field(default_factory=lambda: [0] * JUSTIFICATION_BITS_LENGTH)


@dataclass
class Store:                           # :890-901
    time: int
    genesis_time: int
    justified_checkpoint: Checkpoint
    finalized_checkpoint: Checkpoint
    best_justified_checkpoint: Checkpoint
    proposer_boost_root: bytes
    equivocating_indices: Set[int]
    blocks: Dict[bytes, BeaconBlock] = field(default_factory=dict)
    block_states: Dict[bytes, BeaconState] = field(default_factory=dict)
    checkpoint_states: Dict[Checkpoint, BeaconState] = field(default_factory=dict)
    latest_messages: Dict[int, LatestMessage] = field(default_factory=dict)


# ----------------------------------------------------------------------------- SSZ-lite
def sha256(b: bytes) -> bytes:
    return hashlib.sha256(b).digest()


def _u64_chunk(v: int) -> bytes:
    return int(v).to_bytes(8, "little") + bytes(24)


def _merkle(chunks):
    n = 1
    while n < len(chunks):
        n *= 2
    layer = list(chunks) + [ZERO32] * (n - len(chunks))
    while len(layer) > 1:
        layer = [sha256(layer[i] + layer[i + 1]) for i in range(0, len(layer), 2)]
    return layer[0]


def hash_tree_root_checkpoint(c: Checkpoint) -> bytes:
    return _merkle([_u64_chunk(c.epoch), bytes(c.root)])


def hash_tree_root_attestation_data(d: AttestationData) -> bytes:
    return _merkle([_u64_chunk(d.slot), _u64_chunk(d.index), bytes(d.beacon_block_root),
                    hash_tree_root_checkpoint(d.source), hash_tree_root_checkpoint(d.target)])


def serialize_attestation_data(d: AttestationData) -> bytes:
    """SSZ serialisation of the fixed-size container (128 bytes): the record format of Engine.signing_roots."""
    return (int(d.slot).to_bytes(8, "little") + int(d.index).to_bytes(8, "little") + bytes(d.beacon_block_root)
            + int(d.source.epoch).to_bytes(8, "little") + bytes(d.source.root) + int(d.target.epoch).to_bytes(8, "little") + bytes(d.target.root))


def deserialize_attestation_data(b: bytes) -> AttestationData:
    u = lambda o: int.from_bytes(b[o:o + 8], "little")       # noqa: E731
    return AttestationData(slot=u(0), index=u(8), beacon_block_root=bytes(b[16:48]), source=Checkpoint(u(48), bytes(b[56:88])),
                           target=Checkpoint(u(88), bytes(b[96:128])))


def serialize_attestation(att: Attestation) -> bytes:
    """SSZ wire form (:714-717): [offset of aggregation_bits = 228][data][signature][Bitlist with its delimiter bit]."""
    n = len(att.aggregation_bits)
    bl = bytearray(pack_bits([list(att.aggregation_bits)], stride=n // 8 + 1)[0].tobytes())
    bl[n >> 3] |= 1 << (n & 7)
    return (228).to_bytes(4, "little") + serialize_attestation_data(att.data) + bytes(att.signature) + bytes(bl)


def compute_domain(domain_type: bytes, fork_version: bytes, genesis_validators_root: bytes) -> bytes:
    fork_data_root = _merkle([bytes(fork_version) + bytes(28), bytes(genesis_validators_root)])
    return bytes(domain_type) + fork_data_root[:28]


def integer_squareroot(n: int) -> int:
    x, y = n, (n + 1) // 2
    while y < x:
        x, y = y, (y + n // y) // 2
    return x


def shuffle_permutation(n: int, seed: bytes, rounds: int) -> np.ndarray:
    """perm[i] = compute_shuffled_index(i, n, seed) (:513-534) for every i at once: one pivot hash and
    ceil(n/256) source hashes per round instead of two hashes per index per round."""
    idx = np.arange(n, dtype=np.int64)
    if n == 0:
        return idx.astype(np.uint32)
    nblk = (n + 255) // 256
    for rnd in range(rounds):
        r = bytes([rnd])
        pivot = int.from_bytes(sha256(seed + r)[:8], "little") % n
        src = b"".join(sha256(seed + r + blk.to_bytes(4, "little")) for blk in range(nblk))
        bits = np.unpackbits(np.frombuffer(src, dtype=np.uint8), bitorder="little")
        flip = (pivot + n - idx) % n
        pos = np.maximum(idx, flip)
        idx = np.where(bits[pos] == 1, flip, idx)
    return idx.astype(np.uint32)


def pack_bits(rows, stride=None) -> np.ndarray:
    """list of bool lists -> uint8[n, stride], bit j of row a at byte j>>3, bit j&7 (SSZ Bitlist order)."""
    n = max((len(r) for r in rows), default=0)
    stride = stride or max(1, (n + 7) // 8)
    out = np.zeros((len(rows), stride), dtype=np.uint8)
    for a, r in enumerate(rows):
        if len(r):
            packed = np.packbits(np.asarray(r, dtype=np.uint8), bitorder="little")
            out[a, :packed.shape[0]] = packed
    return out


class Spec:
    """pyspec functions bound to a preset and a GPU engine."""

    def __init__(self, preset: Preset = MAINNET, engine: Engine = None):
        self.p = preset
        self.engine = engine if engine is not None else Engine(0)
        self._registry_key = None          # (length, fingerprint of the pubkey bytes) of the registry on the device
        self._balances_key = None          # (state object, slot) whose effective balances / flags are on the device
        self._balances_state = None
        self._committee_cache = {}
        self._mirror = None                # device mirror of a Store (block tree + latest messages), see _sync_store
        self.registry_valid = None         # KeyValidate verdict per validator of the registry on the device (uint8[n])
        self.stats = {"registry_uploads": 0, "balance_refreshes": 0, "store_uploads": 0, "lmd_device_updates": 0}
        # full: fingerprint = hash of ALL pubkeys on every call (O(n) host work, catches any in-place replacement);
        # sampled: first / last / 62 strided pubkeys (O(1)); auto: full up to 65 536 validators, sampled above
        self.registry_check = "auto"

    # ------------------------------------------------------------------ epochs / registry
    def compute_epoch_at_slot(self, slot):
        return slot // self.p.SLOTS_PER_EPOCH

    def compute_start_slot_at_epoch(self, epoch):
        return epoch * self.p.SLOTS_PER_EPOCH

    def get_current_epoch(self, state):
        return self.compute_epoch_at_slot(state.slot)

    def get_previous_epoch(self, state):
        cur = self.get_current_epoch(state)
        return GENESIS_EPOCH if cur == GENESIS_EPOCH else cur - 1

    @staticmethod
    def is_active_validator(v, epoch):
        return v.activation_epoch <= epoch < v.exit_epoch

    def get_active_validator_indices(self, state, epoch):
        return [i for i, v in enumerate(state.validators) if v.activation_epoch <= epoch < v.exit_epoch]

    def get_total_active_balance(self, state):
        e = self.get_current_epoch(state)
        tot = sum(v.effective_balance for v in state.validators if v.activation_epoch <= e < v.exit_epoch)
        return max(self.p.EFFECTIVE_BALANCE_INCREMENT, tot)

    def get_total_balance(self, state, indices):                             # described at :811
        return max(self.p.EFFECTIVE_BALANCE_INCREMENT, sum(state.validators[i].effective_balance for i in indices))

    @staticmethod
    def has_flag(flags: int, flag_index: int) -> bool:                          # ParticipationFlags helper used at :747
        return (int(flags) >> flag_index) & 1 == 1

    @staticmethod
    def add_flag(flags: int, flag_index: int) -> int:                           # :748
        return int(flags) | (1 << flag_index)

    def get_unslashed_participating_indices(self, state, flag_index, epoch):   # described at :805-807 (host set form)
        cur = self.get_current_epoch(state)
        assert epoch in (self.get_previous_epoch(state), cur)
        table = state.current_epoch_participation if epoch == cur else state.previous_epoch_participation
        return set(i for i in self.get_active_validator_indices(state, epoch)
                   if self.has_flag(table[i], flag_index) and not state.validators[i].slashed)

    # ------------------------------------------------------------------ FFG accounting (:793-803, :817-852)
    def process_justification_and_finalization(self, state):
        """:793-803.  The three balance sums (total active, previous / current epoch TIMELY_TARGET balance of unslashed
        validators) are one pass over the device registry and participation tables (b2_ffg_balances); the 2/3 tests and the
        four finalization rules are scalar host logic."""
        cur = self.get_current_epoch(state)
        if cur <= GENESIS_EPOCH + 1:
            return
        self.sync_registry(state)
        eng = self.engine
        eng.participation_load(0, np.asarray(state.current_epoch_participation, dtype=np.uint8))
        eng.participation_load(1, np.asarray(state.previous_epoch_participation, dtype=np.uint8))
        total, cur_target, prev_target, _ = eng.ffg_balances(TIMELY_TARGET_FLAG_INDEX)
        inc = self.p.EFFECTIVE_BALANCE_INCREMENT                                # get_total_balance's floor
        self.weigh_justification_and_finalization(state, max(inc, total), max(inc, prev_target), max(inc, cur_target))

    def weigh_justification_and_finalization(self, state, total_active_balance, previous_epoch_target_balance,
                                             current_epoch_target_balance):    # :817-852
        prev_epoch, cur_epoch = self.get_previous_epoch(state), self.get_current_epoch(state)
        old_prev, old_cur = state.previous_justified_checkpoint, state.current_justified_checkpoint
        state.previous_justified_checkpoint = old_cur
        bits = [0] + [int(b) for b in state.justification_bits[:JUSTIFICATION_BITS_LENGTH - 1]]
        for pos, epoch, balance in ((1, prev_epoch, previous_epoch_target_balance), (0, cur_epoch, current_epoch_target_balance)):
            if 3 * balance >= 2 * total_active_balance:                        # :830, :834
                state.current_justified_checkpoint = Checkpoint(epoch=epoch, root=self.get_block_root(state, epoch))
                bits[pos] = 1
        state.justification_bits = bits
        # :842-852 -- (bits [lo, hi) all set, source checkpoint, its distance from the current epoch), in the reference's order
        for lo, hi, source, dist in ((1, 4, old_prev, 3), (1, 3, old_prev, 2), (0, 3, old_cur, 2), (0, 2, old_cur, 1)):
            if all(bits[lo:hi]) and source.epoch + dist == cur_epoch:
                state.finalized_checkpoint = source

    def get_randao_mix(self, state, epoch):
        return state.randao_mixes[epoch % self.p.EPOCHS_PER_HISTORICAL_VECTOR]

    def get_block_root_at_slot(self, state, slot):
        assert slot < state.slot <= slot + self.p.SLOTS_PER_HISTORICAL_ROOT
        return state.block_roots[slot % self.p.SLOTS_PER_HISTORICAL_ROOT]

    def get_block_root(self, state, epoch):
        return self.get_block_root_at_slot(state, self.compute_start_slot_at_epoch(epoch))

    # ------------------------------------------------------------------ committees
    def get_committee_count_per_slot(self, state, epoch):                      # :461-468
        n_active = len(self.get_active_validator_indices(state, epoch))
        return max(1, min(self.p.MAX_COMMITTEES_PER_SLOT, n_active // self.p.SLOTS_PER_EPOCH // self.p.TARGET_COMMITTEE_SIZE))

    def get_seed(self, state, epoch, domain_type):                             # :481-486
        mix = self.get_randao_mix(state, epoch + self.p.EPOCHS_PER_HISTORICAL_VECTOR - self.p.MIN_SEED_LOOKAHEAD - 1)
        return sha256(domain_type + int(epoch).to_bytes(8, "little") + mix)

    def compute_shuffled_index(self, index, index_count, seed):                # :513-534 (per-index form)
        assert index < index_count
        for rnd in range(self.p.SHUFFLE_ROUND_COUNT):
            r = bytes([rnd])
            pivot = int.from_bytes(sha256(seed + r)[0:8], "little") % index_count
            flip = (pivot + index_count - index) % index_count
            position = max(index, flip)
            source = sha256(seed + r + (position // 256).to_bytes(4, "little"))
            if (source[(position % 256) // 8] >> (position % 8)) & 1:
                index = flip
        return index

    def compute_committee(self, indices, seed, index, count):                  # :495-504
        n = len(indices)
        start, end = (n * index) // count, (n * (index + 1)) // count
        perm = self._permutation(n, seed)
        return [indices[int(perm[i])] for i in range(start, end)]

    def _permutation(self, n, seed):
        key = (n, seed, self.p.SHUFFLE_ROUND_COUNT)
        if key not in self._committee_cache:
            if len(self._committee_cache) > 8:
                self._committee_cache.clear()
            if hasattr(self.engine, "shuffle_committees"):      # SHA-256 + swap-or-not rounds on the GPU (k_shuffle_*)
                self._committee_cache[key] = self.engine.shuffle_committees(seed, n, self.p.SHUFFLE_ROUND_COUNT)
            else:                                               # engine stand-ins of the CPU tests
                self._committee_cache[key] = shuffle_permutation(n, seed, self.p.SHUFFLE_ROUND_COUNT)
        return self._committee_cache[key]

    def epoch_committees(self, state, epoch):
        """All committees of ``epoch`` in array form: (members u32[n_active], off u32[count+1], committees_per_slot)."""
        active = np.asarray(self.get_active_validator_indices(state, epoch), dtype=np.uint32)
        cps = self.get_committee_count_per_slot(state, epoch)
        seed = self.get_seed(state, epoch, DOMAIN_BEACON_ATTESTER)
        n = len(active)
        members = active[self._permutation(n, seed)]
        count = cps * self.p.SLOTS_PER_EPOCH
        off = np.array([(n * k) // count for k in range(count + 1)], dtype=np.uint32)
        return members, off, cps

    def get_beacon_committee(self, state, slot, index):                        # called at :729
        epoch = self.compute_epoch_at_slot(slot)
        members, off, cps = self.epoch_committees(state, epoch)
        k = (slot % self.p.SLOTS_PER_EPOCH) * cps + index
        return members[off[k]:off[k + 1]].tolist()

    def compute_proposer_index(self, state, indices, seed):                    # :604-618
        assert len(indices) > 0
        i, total = 0, len(indices)
        while True:
            cand = indices[self.compute_shuffled_index(i % total, total, seed)]
            random_byte = sha256(seed + (i // 32).to_bytes(8, "little"))[i % 32]
            if state.validators[cand].effective_balance * 255 >= self.p.MAX_EFFECTIVE_BALANCE * random_byte:
                return cand
            i += 1

    def get_beacon_proposer_index(self, state):                                # called at :754
        epoch = self.get_current_epoch(state)
        seed = sha256(self.get_seed(state, epoch, DOMAIN_BEACON_PROPOSER) + int(state.slot).to_bytes(8, "little"))
        return self.compute_proposer_index(state, self.get_active_validator_indices(state, epoch), seed)

    # ------------------------------------------------------------------ signing
    def get_domain(self, state, domain_type, epoch=None):
        epoch = self.get_current_epoch(state) if epoch is None else epoch
        fv = state.fork.previous_version if epoch < state.fork.epoch else state.fork.current_version
        return compute_domain(domain_type, fv, state.genesis_validators_root)

    @staticmethod
    def compute_signing_root(data: AttestationData, domain):
        return _merkle([hash_tree_root_attestation_data(data), bytes(domain)])

    # ------------------------------------------------------------------ SSZ wire decode on the device (:714-717)
    def decode_attestations(self, encodings):
        """List of SSZ-encoded Attestations -> list of Attestation (None where the encoding is malformed: no delimiter bit, more
        than MAX_VALIDATORS_PER_COMMITTEE bits, truncated container).  One GPU call for the batch (b2_attestations_decode); the
        same call yields the bit rows and the 128-byte data records the verification kernels take."""
        if not encodings:
            return []
        limit = self.p.MAX_VALIDATORS_PER_COMMITTEE
        off = np.zeros(len(encodings) + 1, dtype=np.uint32)
        off[1:] = np.cumsum([len(e) for e in encodings])
        bits, blen, data, sig, st = self.engine.attestations_decode(b"".join(bytes(e) for e in encodings), off, (limit + 7) // 8, limit)
        out = []
        for a in range(len(encodings)):
            if st[a] != 0:
                out.append(None)
                continue
            row = np.unpackbits(bits[a], bitorder="little")[:int(blen[a])]
            out.append(Attestation([bool(x) for x in row], deserialize_attestation_data(data[a].tobytes()), sig[a].tobytes()))
        return out

    # ------------------------------------------------------------------ registry on the device
    def _pubkey_fingerprint(self, validators):
        n = len(validators)
        mode = self.registry_check
        if mode == "auto":
            mode = "full" if n <= 65536 else "sampled"
        h = hashlib.blake2b(digest_size=16)
        if mode == "full":
            h.update(b"".join(bytes(v.pubkey) for v in validators))
        else:
            step = max(1, n // 62)
            for i in list(range(0, n, step)) + [n - 1]:
                h.update(i.to_bytes(8, "little") + bytes(validators[i].pubkey))
        return h.digest()

    def invalidate_registry(self):
        """Forget what is on the device: the next call re-uploads the registry and the store (call it after replacing pubkeys in
        place in a registry larger than 65 536 validators, where the content check is sampled -- see registry_check)."""
        self._registry_key = self._balances_key = self._balances_state = self._mirror = None

    def sync_pubkeys(self, state):
        """state.validators[*].pubkey on the device.  Decompressed + KeyValidated on the GPU ONCE per registry CONTENT: the cache key is
        (length, fingerprint of the pubkey bytes), never an address -- so neither a recycled id() nor an in-place replacement can
        alias a stale registry, and two states that share the same validator set (the justified state and a target checkpoint
        state) share one upload.  Signature verification needs nothing else.  Returns True when it uploaded."""
        validators = state.validators
        n = len(validators)
        key = (n, self._pubkey_fingerprint(validators))
        if key == self._registry_key:
            return False
        pk = np.frombuffer(b"".join(bytes(v.pubkey) for v in validators), dtype=np.uint8).reshape(n, 48)
        zero = np.zeros(n, dtype=np.uint64)
        self.registry_valid = self.engine.registry_load(pk, zero, np.zeros(n, dtype=np.uint8))    # resets the device LMD table too
        self._registry_key = key
        self._balances_key = self._balances_state = self._mirror = None
        self.stats["registry_uploads"] += 1
        return True

    def sync_registry(self, state):
        """sync_pubkeys + the per-validator effective balance / activity / slashed table of THIS state (what the fork-choice weights
        and the FFG sums read).  Refreshed once per (state object, slot): the spec changes these fields only at epoch processing
        (:122-133) and in slashings; a different state (or slot) re-reads them (an O(n) pass over the Python objects)."""
        self.sync_pubkeys(state)
        bkey = (id(state), state.slot)
        if bkey == self._balances_key and state is self._balances_state:
            return
        validators = state.validators
        n = len(validators)
        epoch = self.get_current_epoch(state)
        eff = np.fromiter((v.effective_balance for v in validators), dtype=np.uint64, count=n)
        prev = self.get_previous_epoch(state)
        # bit0 active in the current epoch, bit1 slashed, bit2 active in the previous epoch (the FFG sums need both epochs)
        flags = np.fromiter(((1 if v.activation_epoch <= epoch < v.exit_epoch else 0) | (2 if v.slashed else 0)
                             | (4 if v.activation_epoch <= prev < v.exit_epoch else 0)
                             for v in validators), dtype=np.uint8, count=n)
        self.engine.registry_update_balances(eff, flags)
        self.stats["balance_refreshes"] += 1
        self._balances_key, self._balances_state = bkey, state                  # the state is held: its id() cannot be recycled
        self._total_active = (int(eff[(flags & 1) != 0].astype(object).sum()) if n else 0, int(np.count_nonzero(flags & 1)))

    # ------------------------------------------------------------------ indexed attestations (called at :736, :745, :975-976)
    def get_attesting_indices(self, state, data, bits):
        committee = self.get_beacon_committee(state, data.slot, data.index)
        return set(v for i, v in enumerate(committee) if bits[i])

    def get_indexed_attestation(self, state, attestation):
        idx = self.get_attesting_indices(state, attestation.data, attestation.aggregation_bits)
        return IndexedAttestation(sorted(idx), attestation.data, attestation.signature)

    def is_valid_indexed_attestation(self, state, indexed):
        return bool(self.are_valid_indexed_attestations(state, [indexed])[0])

    def are_valid_indexed_attestations(self, state, indexed_list):
        """Batched is_valid_indexed_attestation: one GPU FastAggregateVerify call for the whole list."""
        if not indexed_list:
            return np.zeros(0, dtype=np.uint8)
        self.sync_pubkeys(state)
        n = len(state.validators)
        members, off, sigs, structurally_ok, data128, domains = [], [0], [], [], [], []
        for ia in indexed_list:
            idx = list(ia.attesting_indices)
            good = len(idx) > 0 and idx == sorted(set(idx)) and idx[-1] < n and len(bytes(ia.signature)) == 96
            structurally_ok.append(good)
            members += idx if good else []
            off.append(len(members))
            domains.append(self.get_domain(state, DOMAIN_BEACON_ATTESTER, ia.data.target.epoch))
            data128.append(serialize_attestation_data(ia.data))
            sigs.append(bytes(ia.signature) if len(bytes(ia.signature)) == 96 else bytes(96))
        if hasattr(self.engine, "signing_roots"):           # SSZ merkleization on the GPU (k_signing_roots)
            msgs_arr = self.engine.signing_roots(np.frombuffer(b"".join(data128), dtype=np.uint8), np.frombuffer(b"".join(domains), dtype=np.uint8))
        else:
            msgs_arr = np.frombuffer(b"".join(self.compute_signing_root(ia.data, dom) for ia, dom in zip(indexed_list, domains)), dtype=np.uint8)
        sizes = np.diff(np.asarray(off, dtype=np.int64))
        stride = max(1, (int(sizes.max()) + 7) // 8)
        bits = np.zeros((len(indexed_list), stride), dtype=np.uint8)
        for a, s in enumerate(sizes):
            full, rem = divmod(int(s), 8)
            bits[a, :full] = 0xFF
            if rem:
                bits[a, full] = (1 << rem) - 1
        ok = self.engine.fast_aggregate_verify(np.asarray(members, dtype=np.uint32), off, bits,
                                               msgs_arr, np.frombuffer(b"".join(sigs), dtype=np.uint8))
        return ok & np.asarray(structurally_ok, dtype=np.uint8)

    # ------------------------------------------------------------------ participation helpers (called at :733, :747-754)
    def get_attestation_participation_flag_indices(self, state, data, inclusion_delay):
        justified = state.current_justified_checkpoint if data.target.epoch == self.get_current_epoch(state) else state.previous_justified_checkpoint
        matching_source = data.source == justified
        matching_target = matching_source and data.target.root == self.get_block_root(state, data.target.epoch)
        matching_head = matching_target and data.beacon_block_root == self.get_block_root_at_slot(state, data.slot)
        assert matching_source
        flags = []
        if matching_source and inclusion_delay <= integer_squareroot(self.p.SLOTS_PER_EPOCH):
            flags.append(TIMELY_SOURCE_FLAG_INDEX)
        if matching_target and inclusion_delay <= self.p.SLOTS_PER_EPOCH:
            flags.append(TIMELY_TARGET_FLAG_INDEX)
        if matching_head and inclusion_delay == self.p.MIN_ATTESTATION_INCLUSION_DELAY:
            flags.append(TIMELY_HEAD_FLAG_INDEX)
        return flags

    def get_base_reward_per_increment(self, state):
        return self.p.EFFECTIVE_BALANCE_INCREMENT * BASE_REWARD_FACTOR // integer_squareroot(self.get_total_active_balance(state))

    def get_base_reward(self, state, index):
        return (state.validators[index].effective_balance // self.p.EFFECTIVE_BALANCE_INCREMENT) * self.get_base_reward_per_increment(state)

    # ------------------------------------------------------------------ process_attestation (:722-754)
    def _check_attestation(self, state, attestation):
        data = attestation.data
        assert data.target.epoch in (self.get_previous_epoch(state), self.get_current_epoch(state))     # :724
        assert data.target.epoch == self.compute_epoch_at_slot(data.slot)                                # :725
        assert data.slot + self.p.MIN_ATTESTATION_INCLUSION_DELAY <= state.slot <= data.slot + self.p.SLOTS_PER_EPOCH   # :726
        assert data.index < self.get_committee_count_per_slot(state, data.target.epoch)                 # :727
        committee = self.get_beacon_committee(state, data.slot, data.index)                              # :729
        assert len(attestation.aggregation_bits) == len(committee)                                       # :730
        flag_indices = self.get_attestation_participation_flag_indices(state, data, state.slot - data.slot)
        return committee, flag_indices

    def _apply_attestation(self, state, attestation, committee, flag_indices):
        data = attestation.data
        participation = state.current_epoch_participation if data.target.epoch == self.get_current_epoch(state) else state.previous_epoch_participation
        per_inc = self.get_base_reward_per_increment(state)
        numerator = 0
        for v, bit in zip(committee, attestation.aggregation_bits):                                      # :745-749
            if not bit:
                continue
            for flag_index, weight in enumerate(PARTICIPATION_FLAG_WEIGHTS):
                if flag_index in flag_indices and not (participation[v] >> flag_index) & 1:
                    participation[v] |= 1 << flag_index
                    numerator += (state.validators[v].effective_balance // self.p.EFFECTIVE_BALANCE_INCREMENT) * per_inc * weight
        denominator = (WEIGHT_DENOMINATOR - PROPOSER_WEIGHT) * WEIGHT_DENOMINATOR // PROPOSER_WEIGHT     # :752
        state.balances[self.get_beacon_proposer_index(state)] += numerator // denominator                # :753-754

    def process_attestation(self, state, attestation):
        committee, flag_indices = self._check_attestation(state, attestation)
        assert self.is_valid_indexed_attestation(state, self.get_indexed_attestation(state, attestation))   # :736 -> GPU
        self._apply_attestation(state, attestation, committee, flag_indices)

    def process_attestations(self, state, attestations):
        """process_operations' attestation loop with ONE GPU verification batch.  Same observable behaviour as calling
        process_attestation in order: the first invalid attestation raises AssertionError after its predecessors were applied."""
        pre = []
        for att in attestations:
            try:
                pre.append(self._check_attestation(state, att) + (self.get_indexed_attestation(state, att),))
            except AssertionError:
                pre.append(None)
                break
        verdicts = self.are_valid_indexed_attestations(state, [p[2] for p in pre if p is not None])
        n_ok = 0
        while n_ok < len(attestations) and n_ok < len(pre) and pre[n_ok] is not None and verdicts[n_ok]:
            n_ok += 1
        if n_ok and hasattr(self.engine, "participation_update"):
            self._apply_attestations_gpu(state, attestations[:n_ok], pre[:n_ok])
        else:
            for k in range(n_ok):
                self._apply_attestation(state, attestations[k], pre[k][0], pre[k][1])
        assert n_ok == len(attestations)

    def _apply_attestations_gpu(self, state, attestations, pre):
        """Flag scatter + proposer-reward numerators of :745-752 for a prefix of valid attestations on the GPU (order-exact),
        then the per-attestation integer division and the proposer credit of :752-754 on the host."""
        eng = self.engine
        self.sync_registry(state)              # the reward numerators read THIS state's effective balances on the device
        cur = self.get_current_epoch(state)
        tables = (state.current_epoch_participation, state.previous_epoch_participation)
        per_inc = self.get_base_reward_per_increment(state)
        denominator = (WEIGHT_DENOMINATOR - PROPOSER_WEIGHT) * WEIGHT_DENOMINATOR // PROPOSER_WEIGHT
        proposer = self.get_beacon_proposer_index(state)
        for which in (0, 1):
            sel = [k for k, att in enumerate(attestations) if (att.data.target.epoch == cur) == (which == 0)]
            if not sel:
                continue
            members, off, rows, masks = [], [0], [], []
            for k in sel:
                members += pre[k][0]
                off.append(len(members))
                rows.append(attestations[k].aggregation_bits)
                masks.append(sum(1 << f for f in pre[k][1]))
            eng.participation_load(which, np.asarray(tables[which], dtype=np.uint8))
            num = eng.participation_update(which, np.asarray(members, dtype=np.uint32), off, pack_bits(rows), np.asarray(masks, dtype=np.uint8), None,
                                           self.p.EFFECTIVE_BALANCE_INCREMENT, per_inc)
            tables[which][:] = eng.participation_read(which).tolist()
            for v in num:
                state.balances[proposer] += int(v) // denominator

    # ------------------------------------------------------------------ fork choice
    def update_latest_messages(self, store, attesting_indices, attestation):   # :1435-1441
        """The store's dict is updated as the reference does; when the store is mirrored on the device (after a get_head /
        get_weight call) the same update is applied to the device table by K7 (b2_latest_messages_update), so the next get_head
        is a kernel launch, not a re-upload."""
        target, root = attestation.data.target, attestation.data.beacon_block_root
        indices = [int(i) for i in attesting_indices]
        for i in indices:
            if i in store.equivocating_indices:
                continue
            if i not in store.latest_messages or target.epoch > store.latest_messages[i].epoch:
                store.latest_messages[i] = LatestMessage(epoch=target.epoch, root=root)
        m = self._mirror
        if m is not None and m["store"] is store:
            blk = m["index"].get(root)
            if blk is None or not indices or max(indices) >= m["n_val"] or target.epoch >= 0xFFFFFFFF:
                self._mirror = None                  # a vote the device form cannot hold: fall back to a full re-upload
            else:
                k = len(indices)
                bits = np.full((1, (k + 7) // 8), 0xFF, dtype=np.uint8)
                if k & 7:
                    bits[0, -1] = (1 << (k & 7)) - 1
                self.engine.latest_messages_update(np.asarray(indices, dtype=np.uint32), [0, k], bits, [target.epoch], [blk])
                m["n_msgs"] = len(store.latest_messages)
                self.stats["lmd_device_updates"] += 1

    def get_current_slot(self, store):
        return (store.time - store.genesis_time) // getattr(self.p, "SECONDS_PER_SLOT", 12)

    def get_ancestor(self, store, root, slot):
        block = store.blocks[root]
        while block.slot > slot:
            root = block.parent_root
            block = store.blocks[root]
        return root

    def validate_on_attestation(self, store, attestation, is_from_block):      # called at :970 (upstream fork-choice spec, v1.2.0)
        target = attestation.data.target
        if not is_from_block:                                                   # validate_target_epoch_against_current_time
            current_epoch = self.compute_epoch_at_slot(self.get_current_slot(store))
            previous_epoch = current_epoch - 1 if current_epoch > GENESIS_EPOCH else GENESIS_EPOCH
            assert target.epoch in (current_epoch, previous_epoch)
        assert target.epoch == self.compute_epoch_at_slot(attestation.data.slot)
        assert target.root in store.blocks
        assert attestation.data.beacon_block_root in store.blocks
        assert store.blocks[attestation.data.beacon_block_root].slot <= attestation.data.slot
        assert target.root == self.get_ancestor(store, attestation.data.beacon_block_root, self.compute_start_slot_at_epoch(target.epoch))
        assert self.get_current_slot(store) >= attestation.data.slot + 1

    def on_attestation(self, store, attestation, is_from_block=False):         # :963-979 / :1423-1428
        """AssertionError = invalid, as in the reference.  store_target_checkpoint_state (:971) needs process_slots, which is state
        transition and out of scope: the target checkpoint state must already be in store.checkpoint_states
        (the caller's on_block / on_tick put it there), otherwise the attestation is rejected."""
        self.validate_on_attestation(store, attestation, is_from_block)
        assert attestation.data.target in store.checkpoint_states, "target checkpoint state not in store (store_target_checkpoint_state is out of scope)"
        target_state = store.checkpoint_states[attestation.data.target]
        indexed = self.get_indexed_attestation(target_state, attestation)
        assert self.is_valid_indexed_attestation(target_state, indexed)
        self.update_latest_messages(store, indexed.attesting_indices, attestation)

    @staticmethod
    def is_slashable_attestation_data(data_1, data_2):                         # :1134-1143
        double_vote = data_1 != data_2 and data_1.target.epoch == data_2.target.epoch
        surround_vote = data_1.source.epoch < data_2.source.epoch and data_2.target.epoch < data_1.target.epoch
        return double_vote or surround_vote

    def on_attester_slashing(self, store, attester_slashing):                  # :1447-1461
        a1, a2 = attester_slashing.attestation_1, attester_slashing.attestation_2
        assert self.is_slashable_attestation_data(a1.data, a2.data)
        state = store.block_states[store.justified_checkpoint.root]
        ok = self.are_valid_indexed_attestations(state, [a1, a2])              # both signatures in one GPU batch
        assert ok[0] and ok[1]
        both = sorted(set(a1.attesting_indices).intersection(a2.attesting_indices))
        for index in both:
            store.equivocating_indices.add(index)
        m = self._mirror
        if m is not None and m["store"] is store:
            if hasattr(self.engine, "on_attester_slashing") and both and both[-1] < m["n_val"]:
                self.engine.on_attester_slashing(both, both)                    # mark them on the device table as well
                m["n_equiv"] = len(store.equivocating_indices)
            else:
                self._mirror = None

    def _store_arrays(self, store):
        """Store (dicts) -> the array form of include/b200pos.h: blocks below the justified root in topological order."""
        import collections
        jroot = store.justified_checkpoint.root
        children = {}
        for r, b in store.blocks.items():
            children.setdefault(b.parent_root, []).append(r)
        order, queue = [], collections.deque([jroot])
        while queue:
            r = queue.popleft()
            order.append(r)
            queue.extend(children.get(r, []))
        index = {r: i for i, r in enumerate(order)}
        nb = len(order)
        parent = np.zeros(nb, dtype=np.uint32)
        slot = np.zeros(nb, dtype=np.uint64)
        viable = np.ones(nb, dtype=np.uint8)
        for i, r in enumerate(order):
            b = store.blocks[r]
            slot[i] = b.slot
            if i:
                parent[i] = index[b.parent_root]
            if r not in children:
                hs = store.block_states[r]
                cj = store.justified_checkpoint.epoch == GENESIS_EPOCH or hs.current_justified_checkpoint == store.justified_checkpoint
                cf = store.finalized_checkpoint.epoch == GENESIS_EPOCH or hs.finalized_checkpoint == store.finalized_checkpoint
                viable[i] = 1 if (cj and cf) else 0
        roots = np.frombuffer(b"".join(order), dtype=np.uint8).reshape(nb, 32)
        return order, index, parent, slot, roots, viable

    def _sync_store(self, store):
        """Device mirror of the store: block tree (b2_tree_load) and latest messages (b2_latest_messages_load) are uploaded when the
        store's SHAPE changed -- another store object, justified / finalized checkpoint moved, a block was added (on_block), the
        registry changed, or latest_messages / equivocating_indices were changed behind this class's back (their sizes differ from
        what the mirror recorded).  Votes that arrive through on_attestation / update_latest_messages are applied to the device
        table incrementally, so between blocks get_head / get_weight cost one kernel sequence."""
        state = store.checkpoint_states[store.justified_checkpoint]
        self.sync_registry(state)
        m = self._mirror
        live = (m is not None and m["store"] is store and m["justified"] == store.justified_checkpoint and m["finalized"] == store.finalized_checkpoint
                and m["n_blocks"] == len(store.blocks) and m["n_msgs"] == len(store.latest_messages) and m["n_equiv"] == len(store.equivocating_indices)
                and m["n_val"] == len(state.validators))
        if not live:
            order, index, parent, slot, roots, viable = self._store_arrays(store)
            self.engine.tree_load(parent, slot, roots, viable)
            n = len(state.validators)
            epoch = np.zeros(n, dtype=np.uint64)
            blk = np.zeros(n, dtype=np.uint32)
            has = np.zeros(n, dtype=np.uint8)
            for v, lm in store.latest_messages.items():
                if v < n and lm.root in index:
                    epoch[v], blk[v], has[v] = lm.epoch, index[lm.root], 1
            eq = np.zeros(n, dtype=np.uint8)
            for v in store.equivocating_indices:
                if v < n:
                    eq[v] = 1
            self.engine.latest_messages_load(epoch, blk, has, eq)
            m = self._mirror = dict(store=store, justified=store.justified_checkpoint, finalized=store.finalized_checkpoint, n_blocks=len(store.blocks),
                                    n_msgs=len(store.latest_messages), n_equiv=len(store.equivocating_indices), n_val=n, order=order, index=index,
                                    boost=None)
            self.stats["store_uploads"] += 1
        bkey = (store.proposer_boost_root, self._balances_key)
        if m["boost"] is None or m["boost"][0] != bkey:
            boost_idx, boost_score = -1, 0
            if store.proposer_boost_root != ZERO32 and store.proposer_boost_root in m["index"]:
                total, num = self._total_active                                 # of the justified state, computed by sync_registry
                avg = max(self.p.EFFECTIVE_BALANCE_INCREMENT, total) // num
                boost_score = (num // self.p.SLOTS_PER_EPOCH) * avg * self.p.PROPOSER_SCORE_BOOST // 100
                boost_idx = m["index"][store.proposer_boost_root]
            m["boost"] = (bkey, boost_idx, boost_score)
        return m["order"], m["index"], m["boost"][1], m["boost"][2]

    def get_latest_attesting_balance(self, store, root):                       # called at :1116 (v1.2.0 form)
        order, index, boost_idx, boost_score = self._sync_store(store)
        return int(self.engine.get_weights(boost_idx, boost_score)[index[root]])

    get_weight = get_latest_attesting_balance

    def get_head(self, store):                                                 # :1102-1116
        order, index, boost_idx, boost_score = self._sync_store(store)
        return order[self.engine.get_head(0, boost_idx, boost_score)]
