"""SSZ-lite: hash_tree_root of the few containers on the path -- CPU oracle (TEST INFRASTRUCTURE).

The reference only *calls* ``hash_tree_root`` / ``compute_signing_root`` / ``compute_domain``
(/root/reference/pos-evolution.md:142, :162-163, :423, :1016-1024); the merkleization
rules are restated from the SSZ spec: basic values are
little-endian, padded to 32-byte chunks; a container's root is the binary SHA-256 Merkle
root of its field roots padded with zero chunks to a power of two.
"""
import hashlib

ZERO32 = bytes(32)


def sha256(b: bytes) -> bytes:
    return hashlib.sha256(b).digest()


def uint_chunk(v: int) -> bytes:
    return int(v).to_bytes(8, "little") + bytes(24)


def merkleize(chunks):
    n = 1
    while n < len(chunks):
        n *= 2
    layer = list(chunks) + [ZERO32] * (n - len(chunks))
    while len(layer) > 1:
        layer = [sha256(layer[i] + layer[i + 1]) for i in range(0, len(layer), 2)]
    return layer[0]


def htr_checkpoint(epoch: int, root: bytes) -> bytes:
    return merkleize([uint_chunk(epoch), bytes(root)])


def htr_attestation_data(slot, index, beacon_block_root, source_epoch, source_root, target_epoch, target_root) -> bytes:
    return merkleize([uint_chunk(slot), uint_chunk(index), bytes(beacon_block_root),
                      htr_checkpoint(source_epoch, source_root), htr_checkpoint(target_epoch, target_root)])


def compute_fork_data_root(current_version: bytes, genesis_validators_root: bytes) -> bytes:
    return merkleize([bytes(current_version) + bytes(28), bytes(genesis_validators_root)])


def compute_domain(domain_type: bytes, fork_version: bytes, genesis_validators_root: bytes) -> bytes:
    return bytes(domain_type) + compute_fork_data_root(fork_version, genesis_validators_root)[:28]


def compute_signing_root_from_object_root(object_root: bytes, domain: bytes) -> bytes:
    return merkleize([bytes(object_root), bytes(domain)])
