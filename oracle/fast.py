"""Array-form (numpy) restatement of the non-BLS half of the path -- CPU oracle (TEST
INFRASTRUCTURE).  Same results as oracle/spec.py (tests/test_oracle_literal_spec.py proves it
on small cases) but able to run the BASELINE.json sizes (2^20 validators, 10^4 blocks) in
seconds, so it is the full-size checker for the CUDA kernels K7-K9 and the shuffle.

Array layout mirrors include/b200pos.h: block tree in topological order (parent index <
child index, block 0 = justified root), latest messages as (epoch u64[N], block_idx u32[N],
has_msg u8[N]).
"""
import hashlib

import numpy as np


def shuffle_permutation(n: int, seed: bytes, rounds: int) -> np.ndarray:
    """perm[i] == compute_shuffled_index(i, n, seed) for all i (ref /root/reference/pos-evolution.md:513-534),
    evaluated for the whole list at once: per round one pivot hash plus ceil(n/256) source hashes."""
    idx = np.arange(n, dtype=np.int64)
    if n == 0:
        return idx.astype(np.uint32)
    nblk = (n + 255) // 256
    for rnd in range(rounds):
        r = bytes([rnd])
        pivot = int.from_bytes(hashlib.sha256(seed + r).digest()[:8], "little") % n
        src = b"".join(hashlib.sha256(seed + r + blk.to_bytes(4, "little")).digest() for blk in range(nblk))
        bits = np.unpackbits(np.frombuffer(src, dtype=np.uint8), bitorder="little")
        flip = (pivot + n - idx) % n
        pos = np.maximum(idx, flip)
        idx = np.where(bits[pos] == 1, flip, idx)
    return idx.astype(np.uint32)


def committees_for_epoch(active: np.ndarray, seed: bytes, rounds: int, committees_per_slot: int, slots_per_epoch: int):
    """All committees of an epoch: returns (members u32[n_active] in committee order, offsets u32[count+1]).
    Committee k (= slot_in_epoch*cps + index) is members[off[k]:off[k+1]]   (ref :495-504 + get_beacon_committee ext)."""
    n = len(active)
    perm = shuffle_permutation(n, seed, rounds)
    members = np.asarray(active, dtype=np.uint32)[perm]
    count = committees_per_slot * slots_per_epoch
    off = np.array([(n * k) // count for k in range(count + 1)], dtype=np.uint32)
    return members, off


def lmd_update(msg_epoch, msg_block, has_msg, equivocating, indices, target_epoch: int, block_idx: int):
    """update_latest_messages (ref :1435-1441) on arrays, in place, for one attestation."""
    idx = np.asarray(indices, dtype=np.int64)
    idx = idx[equivocating[idx] == 0]
    upd = idx[(has_msg[idx] == 0) | (target_epoch > msg_epoch[idx])]
    msg_epoch[upd] = target_epoch
    msg_block[upd] = block_idx
    has_msg[upd] = 1


def ghost_weights(parent, msg_block, has_msg, eff_bal, active, equivocating, boost_idx: int, boost_score: int):
    """weight[b] = get_latest_attesting_balance(store, root_b) for every block b (ext; called ref :1116).
    Uses the subtree-sum equivalence of SURVEY.md appendix C.5; boost_idx < 0 means no boost."""
    nb = len(parent)
    w = np.zeros(nb, dtype=np.uint64)
    m = (has_msg != 0) & (active != 0) & (equivocating == 0)
    np.add.at(w, msg_block[m].astype(np.int64), eff_bal[m].astype(np.uint64))
    if boost_idx >= 0:
        w[boost_idx] += np.uint64(boost_score)
    for b in range(nb - 1, 0, -1):
        w[parent[b]] += w[b]
    return w


def ghost_viable(parent, leaf_viable):
    """get_filtered_block_tree (ext; called ref :1104): keep[b] iff some leaf below b is viable."""
    nb = len(parent)
    has_child = np.zeros(nb, dtype=bool)
    has_child[np.asarray(parent[1:], dtype=np.int64)] = True
    keep = np.where(has_child, False, np.asarray(leaf_viable) != 0)
    for b in range(nb - 1, 0, -1):
        if keep[b]:
            keep[parent[b]] = True
    return keep


def ghost_head(parent, roots, keep, weight, justified_idx: int = 0) -> int:
    """get_head walk (ref :1102-1116): argmax (weight, root) over kept children until a leaf."""
    nb = len(parent)
    children = [[] for _ in range(nb)]
    for b in range(1, nb):
        if keep[b]:
            children[parent[b]].append(b)
    head = justified_idx
    while children[head]:
        head = max(children[head], key=lambda c: (int(weight[c]), bytes(roots[c])))
    return head


def proposer_boost_score(eff_bal, active, slots_per_epoch: int, boost_pct: int, increment: int = 10**9) -> int:
    """v1.2.0 form."""
    num = int(np.count_nonzero(active))
    total = max(increment, int(eff_bal[active != 0].astype(object).sum()))
    avg = total // num
    return (num // slots_per_epoch) * avg * boost_pct // 100
