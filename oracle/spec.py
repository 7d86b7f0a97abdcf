"""Executable restatement of the pyspec path quoted by the reference -- CPU oracle
(TEST INFRASTRUCTURE, see oracle/__init__.py).

Every function cites the line of /root/reference/pos-evolution.md it follows ("ref :N") or,
for helpers the reference calls but does not define, the call site plus "ext" (semantics
restated from consensus-specs ~v1.2.0).  Pure Python, literal,
slow: it is the bit-exact yardstick, not a product path.
"""
from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Set

from . import bls_sig as bls
from .ssz import (sha256, htr_attestation_data, compute_domain as _compute_domain,
                  compute_signing_root_from_object_root, ZERO32)

hash32 = sha256        # pyspec calls this ``hash``; not shadowing the builtin (dataclass __hash__ needs it)

FAR_FUTURE_EPOCH = 2**64 - 1
GENESIS_EPOCH = 0
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
    """SURVEY.md appendix A (values recalled from consensus-specs; the reference pins none)."""
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


# ----------------------------------------------------------------------------- containers (ref :36, :219, :287, :689, :714, :890)
@dataclass
class Validator:                       # ref :36-45
    pubkey: bytes
    effective_balance: int
    slashed: bool = False
    activation_epoch: int = 0
    exit_epoch: int = FAR_FUTURE_EPOCH


@dataclass(frozen=True)
class Checkpoint:                      # ref :219-221
    epoch: int = 0
    root: bytes = ZERO32


@dataclass(frozen=True)
class AttestationData:                 # ref :689-697
    slot: int
    index: int
    beacon_block_root: bytes
    source: Checkpoint
    target: Checkpoint


@dataclass
class Attestation:                     # ref :714-717
    aggregation_bits: List[bool]
    data: AttestationData
    signature: bytes


@dataclass
class IndexedAttestation:              # ext; used ref :1161-1162
    attesting_indices: List[int]
    data: AttestationData
    signature: bytes


@dataclass(frozen=True)
class LatestMessage:                   # ref :287-289
    epoch: int
    root: bytes


@dataclass
class Fork:
    previous_version: bytes = bytes(4)
    current_version: bytes = bytes(4)
    epoch: int = 0


@dataclass
class BeaconBlock:                     # ref :671-676 (only the fields fork choice reads)
    slot: int
    parent_root: bytes


@dataclass
class BeaconState:                     # ref :338-374 (only the fields the path reads/writes)
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
class Store:                           # ref :890-901
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


def uint_to_bytes(v: int, n: int) -> bytes:
    return int(v).to_bytes(n, "little")


def bytes_to_uint64(b: bytes) -> int:
    return int.from_bytes(b, "little")


def integer_squareroot(n: int) -> int:
    x, y = n, (n + 1) // 2
    while y < x:
        x, y = y, (y + n // y) // 2
    return x


def hash_tree_root_attestation_data(d: AttestationData) -> bytes:
    return htr_attestation_data(d.slot, d.index, d.beacon_block_root, d.source.epoch, d.source.root,
                                d.target.epoch, d.target.root)


class Spec:
    """pyspec functions bound to one preset; names and signatures follow the reference."""

    def __init__(self, preset: Preset = MAINNET, bls_module=bls):
        self.p = preset
        self.bls = bls_module

    # ------------------------------------------------------------------ epochs / registry (ext)
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
    def is_active_validator(v: Validator, epoch):
        return v.activation_epoch <= epoch < v.exit_epoch

    def get_active_validator_indices(self, state, epoch):
        return [i for i, v in enumerate(state.validators) if self.is_active_validator(v, epoch)]

    def get_total_active_balance(self, state):
        tot = sum(state.validators[i].effective_balance
                  for i in self.get_active_validator_indices(state, self.get_current_epoch(state)))
        return max(self.p.EFFECTIVE_BALANCE_INCREMENT, tot)

    def get_total_balance(self, state, indices):                             # ext; described at ref :811
        return max(self.p.EFFECTIVE_BALANCE_INCREMENT, sum(state.validators[i].effective_balance for i in indices))

    def get_unslashed_participating_indices(self, state, flag_index, epoch):   # ext; described at ref :805-807
        cur = self.get_current_epoch(state)
        assert epoch in (self.get_previous_epoch(state), cur)
        table = state.current_epoch_participation if epoch == cur else state.previous_epoch_participation
        return set(i for i in self.get_active_validator_indices(state, epoch)
                   if self.has_flag(table[i], flag_index) and not state.validators[i].slashed)

    # ------------------------------------------------------------------ FFG accounting (ref :793-803, :817-852)
    def process_justification_and_finalization(self, state):
        cur = self.get_current_epoch(state)
        if cur <= GENESIS_EPOCH + 1:
            return
        prev_idx = self.get_unslashed_participating_indices(state, TIMELY_TARGET_FLAG_INDEX, self.get_previous_epoch(state))
        cur_idx = self.get_unslashed_participating_indices(state, TIMELY_TARGET_FLAG_INDEX, cur)
        self.weigh_justification_and_finalization(state, self.get_total_active_balance(state),
                                                  self.get_total_balance(state, prev_idx), self.get_total_balance(state, cur_idx))

    def weigh_justification_and_finalization(self, state, total_active_balance, previous_epoch_target_balance,
                                             current_epoch_target_balance):
        prev_epoch, cur_epoch = self.get_previous_epoch(state), self.get_current_epoch(state)
        old_prev, old_cur = state.previous_justified_checkpoint, state.current_justified_checkpoint
        state.previous_justified_checkpoint = old_cur
        bits = [0] + [int(b) for b in state.justification_bits[:JUSTIFICATION_BITS_LENGTH - 1]]   # a new epoch enters at position 0
        # justification: 2/3 of the active stake voted for the target (ref :830, :834); the current epoch overrides the previous one
        for pos, epoch, balance in ((1, prev_epoch, previous_epoch_target_balance), (0, cur_epoch, current_epoch_target_balance)):
            if 3 * balance >= 2 * total_active_balance:
                state.current_justified_checkpoint = Checkpoint(epoch=epoch, root=self.get_block_root(state, epoch))
                bits[pos] = 1
        state.justification_bits = bits
        # finalization, the four rules of ref :842-852 in their order:
        # (bits [lo, hi) all justified, source checkpoint, distance of the source from the current epoch)
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
    def get_committee_count_per_slot(self, state, epoch):                      # ref :461-468
        n_active = len(self.get_active_validator_indices(state, epoch))
        return max(1, min(self.p.MAX_COMMITTEES_PER_SLOT,
                          n_active // self.p.SLOTS_PER_EPOCH // self.p.TARGET_COMMITTEE_SIZE))

    def get_seed(self, state, epoch, domain_type):                             # ref :481-486
        mix = self.get_randao_mix(state, epoch + self.p.EPOCHS_PER_HISTORICAL_VECTOR - self.p.MIN_SEED_LOOKAHEAD - 1)
        return hash32(domain_type + uint_to_bytes(epoch, 8) + mix)

    def compute_shuffled_index(self, index, index_count, seed):                # ref :513-534
        assert index < index_count
        for rnd in range(self.p.SHUFFLE_ROUND_COUNT):
            r = uint_to_bytes(rnd, 1)
            pivot = bytes_to_uint64(hash32(seed + r)[0:8]) % index_count
            flip = (pivot + index_count - index) % index_count
            position = max(index, flip)
            source = hash32(seed + r + uint_to_bytes(position // 256, 4))
            bit = (source[(position % 256) // 8] >> (position % 8)) & 1
            index = flip if bit else index
        return index

    def compute_committee(self, indices, seed, index, count):                  # ref :495-504
        start = (len(indices) * index) // count
        end = (len(indices) * (index + 1)) // count
        return [indices[self.compute_shuffled_index(i, len(indices), seed)] for i in range(start, end)]

    def get_beacon_committee(self, state, slot, index):                        # ext, called ref :729
        epoch = self.compute_epoch_at_slot(slot)
        cps = self.get_committee_count_per_slot(state, epoch)
        return self.compute_committee(
            indices=self.get_active_validator_indices(state, epoch),
            seed=self.get_seed(state, epoch, DOMAIN_BEACON_ATTESTER),
            index=(slot % self.p.SLOTS_PER_EPOCH) * cps + index,
            count=cps * self.p.SLOTS_PER_EPOCH)

    def compute_proposer_index(self, state, indices, seed):                    # ref :604-618
        assert len(indices) > 0
        i, total = 0, len(indices)
        while True:
            cand = indices[self.compute_shuffled_index(i % total, total, seed)]
            random_byte = hash32(seed + uint_to_bytes(i // 32, 8))[i % 32]
            if state.validators[cand].effective_balance * 255 >= self.p.MAX_EFFECTIVE_BALANCE * random_byte:
                return cand
            i += 1

    def get_beacon_proposer_index(self, state):                                # ext, called ref :754
        epoch = self.get_current_epoch(state)
        seed = hash32(self.get_seed(state, epoch, DOMAIN_BEACON_PROPOSER) + uint_to_bytes(state.slot, 8))
        return self.compute_proposer_index(state, self.get_active_validator_indices(state, epoch), seed)

    # ------------------------------------------------------------------ signing
    def get_domain(self, state, domain_type, epoch=None):                      # ext (pattern ref :162)
        epoch = self.get_current_epoch(state) if epoch is None else epoch
        fv = state.fork.previous_version if epoch < state.fork.epoch else state.fork.current_version
        return _compute_domain(domain_type, fv, state.genesis_validators_root)

    @staticmethod
    def compute_signing_root(data: AttestationData, domain):                   # ext (pattern ref :163)
        return compute_signing_root_from_object_root(hash_tree_root_attestation_data(data), domain)

    # ------------------------------------------------------------------ indexed attestations (ext; called ref :736, :745, :975-976)
    def get_attesting_indices(self, state, data, bits):
        committee = self.get_beacon_committee(state, data.slot, data.index)
        return set(idx for i, idx in enumerate(committee) if bits[i])

    def get_indexed_attestation(self, state, attestation):
        idx = self.get_attesting_indices(state, attestation.data, attestation.aggregation_bits)
        return IndexedAttestation(sorted(idx), attestation.data, attestation.signature)

    def is_valid_indexed_attestation(self, state, indexed):
        indices = list(indexed.attesting_indices)
        if len(indices) == 0 or indices != sorted(set(indices)):
            return False
        pubkeys = [state.validators[i].pubkey for i in indices]
        domain = self.get_domain(state, DOMAIN_BEACON_ATTESTER, indexed.data.target.epoch)
        signing_root = self.compute_signing_root(indexed.data, domain)
        return self.bls.FastAggregateVerify(pubkeys, signing_root, indexed.signature)

    # ------------------------------------------------------------------ altair participation helpers (ext; called ref :733, :747-754)
    def get_attestation_participation_flag_indices(self, state, data, inclusion_delay):
        if data.target.epoch == self.get_current_epoch(state):
            justified = state.current_justified_checkpoint
        else:
            justified = state.previous_justified_checkpoint
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
        return (self.p.EFFECTIVE_BALANCE_INCREMENT * BASE_REWARD_FACTOR
                // integer_squareroot(self.get_total_active_balance(state)))

    def get_base_reward(self, state, index):
        inc = state.validators[index].effective_balance // self.p.EFFECTIVE_BALANCE_INCREMENT
        return inc * self.get_base_reward_per_increment(state)

    @staticmethod
    def has_flag(flags, flag_index):
        return (flags >> flag_index) & 1 == 1

    @staticmethod
    def add_flag(flags, flag_index):
        return flags | (1 << flag_index)

    @staticmethod
    def increase_balance(state, index, delta):
        state.balances[index] += delta

    # ------------------------------------------------------------------ process_attestation (ref :722-754)
    def process_attestation(self, state, attestation):
        data = attestation.data
        assert data.target.epoch in (self.get_previous_epoch(state), self.get_current_epoch(state))   # ref :724
        assert data.target.epoch == self.compute_epoch_at_slot(data.slot)                              # ref :725
        assert (data.slot + self.p.MIN_ATTESTATION_INCLUSION_DELAY <= state.slot
                <= data.slot + self.p.SLOTS_PER_EPOCH)                                                 # ref :726
        assert data.index < self.get_committee_count_per_slot(state, data.target.epoch)               # ref :727
        committee = self.get_beacon_committee(state, data.slot, data.index)                            # ref :729
        assert len(attestation.aggregation_bits) == len(committee)                                     # ref :730
        flag_indices = self.get_attestation_participation_flag_indices(state, data, state.slot - data.slot)
        assert self.is_valid_indexed_attestation(state, self.get_indexed_attestation(state, attestation))  # ref :736
        if data.target.epoch == self.get_current_epoch(state):                                         # ref :739-742
            participation = state.current_epoch_participation
        else:
            participation = state.previous_epoch_participation
        numerator = 0
        for index in self.get_attesting_indices(state, data, attestation.aggregation_bits):           # ref :745-749
            for flag_index, weight in enumerate(PARTICIPATION_FLAG_WEIGHTS):
                if flag_index in flag_indices and not self.has_flag(participation[index], flag_index):
                    participation[index] = self.add_flag(participation[index], flag_index)
                    numerator += self.get_base_reward(state, index) * weight
        denominator = (WEIGHT_DENOMINATOR - PROPOSER_WEIGHT) * WEIGHT_DENOMINATOR // PROPOSER_WEIGHT   # ref :752
        self.increase_balance(state, self.get_beacon_proposer_index(state), numerator // denominator)  # ref :753-754

    # ------------------------------------------------------------------ fork choice
    def update_latest_messages(self, store, attesting_indices, attestation):   # ref :1435-1441
        target = attestation.data.target
        root = attestation.data.beacon_block_root
        for i in attesting_indices:
            if i in store.equivocating_indices:
                continue
            if i not in store.latest_messages or target.epoch > store.latest_messages[i].epoch:
                store.latest_messages[i] = LatestMessage(epoch=target.epoch, root=root)

    def on_attestation(self, store, attestation, is_from_block=False):         # ref :963-979 / :1423-1428
        """validate_on_attestation / store_target_checkpoint_state are store-maintenance (SURVEY.md
        section 2 row 10, out of scope): the caller provides checkpoint_states[target]."""
        target_state = store.checkpoint_states[attestation.data.target]
        indexed = self.get_indexed_attestation(target_state, attestation)
        assert self.is_valid_indexed_attestation(target_state, indexed)
        self.update_latest_messages(store, indexed.attesting_indices, attestation)

    def get_ancestor(self, store, root, slot):                                 # ext; called ref :953, :1005, :1058
        block = store.blocks[root]
        while block.slot > slot:
            root = block.parent_root
            block = store.blocks[root]
        return root

    def get_latest_attesting_balance(self, store, root):                       # ext; called ref :1116 (v1.2.0 form)
        state = store.checkpoint_states[store.justified_checkpoint]
        active = self.get_active_validator_indices(state, self.get_current_epoch(state))
        slot = store.blocks[root].slot
        score = sum(state.validators[i].effective_balance for i in active
                    if (i in store.latest_messages and i not in store.equivocating_indices
                        and self.get_ancestor(store, store.latest_messages[i].root, slot) == root))
        if store.proposer_boost_root == ZERO32:
            return score
        proposer_score = 0
        if self.get_ancestor(store, store.proposer_boost_root, slot) == root:
            num = len(active)
            avg = self.get_total_active_balance(state) // num
            committee_weight = (num // self.p.SLOTS_PER_EPOCH) * avg
            proposer_score = committee_weight * self.p.PROPOSER_SCORE_BOOST // 100
        return score + proposer_score

    get_weight = get_latest_attesting_balance                                  # v1.3+ name used by north_star

    def filter_block_tree(self, store, block_root, blocks):                    # ext (prose ref :874, :1121-1124)
        block = store.blocks[block_root]
        children = [r for r in store.blocks if store.blocks[r].parent_root == block_root]
        if children:
            ok = [self.filter_block_tree(store, c, blocks) for c in children]
            if any(ok):
                blocks[block_root] = block
                return True
            return False
        head_state = store.block_states[block_root]
        correct_justified = (store.justified_checkpoint.epoch == GENESIS_EPOCH
                             or head_state.current_justified_checkpoint == store.justified_checkpoint)
        correct_finalized = (store.finalized_checkpoint.epoch == GENESIS_EPOCH
                             or head_state.finalized_checkpoint == store.finalized_checkpoint)
        if correct_justified and correct_finalized:
            blocks[block_root] = block
            return True
        return False

    def get_filtered_block_tree(self, store):                                  # ext; called ref :1104
        blocks: Dict[bytes, BeaconBlock] = {}
        self.filter_block_tree(store, store.justified_checkpoint.root, blocks)
        return blocks

    def get_head(self, store):                                                 # ref :1102-1116
        blocks = self.get_filtered_block_tree(store)
        head = store.justified_checkpoint.root
        while True:
            children = [r for r in blocks if blocks[r].parent_root == head]
            if not children:
                return head
            head = max(children, key=lambda r: (self.get_latest_attesting_balance(store, r), r))
