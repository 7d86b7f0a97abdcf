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


# ---- SSZ wire form of Attestation (ref :714-717; AttestationData :689-697, Checkpoint :219-221)
ATTESTATION_FIXED_SIZE = 4 + 128 + 96          # offset of aggregation_bits, AttestationData, signature


def serialize_attestation_data(slot, index, beacon_block_root, source_epoch, source_root, target_epoch, target_root) -> bytes:
    u = lambda v: int(v).to_bytes(8, "little")     # noqa: E731
    return u(slot) + u(index) + bytes(beacon_block_root) + u(source_epoch) + bytes(source_root) + u(target_epoch) + bytes(target_root)


def serialize_bitlist(bits) -> bytes:
    """Bitlist: the bits, little-endian within bytes, followed by one delimiter bit (ref :715 discusses it)."""
    n = len(bits)
    out = bytearray(n // 8 + 1)
    for i, b in enumerate(bits):
        if b:
            out[i >> 3] |= 1 << (i & 7)
    out[n >> 3] |= 1 << (n & 7)
    return bytes(out)


def deserialize_bitlist(b: bytes, limit: int):
    if len(b) == 0 or b[-1] == 0:
        raise ValueError("bitlist without delimiter")
    n = 8 * (len(b) - 1) + b[-1].bit_length() - 1
    if n > limit:
        raise ValueError("bitlist longer than its limit")
    return [bool((b[i >> 3] >> (i & 7)) & 1) for i in range(n)]


def serialize_attestation(bits, data128: bytes, signature: bytes) -> bytes:
    assert len(data128) == 128 and len(signature) == 96
    return ATTESTATION_FIXED_SIZE.to_bytes(4, "little") + data128 + signature + serialize_bitlist(bits)


def deserialize_attestation(b: bytes, limit: int):
    """-> (bits, data128, signature); ValueError on a malformed encoding."""
    if len(b) < ATTESTATION_FIXED_SIZE or int.from_bytes(b[:4], "little") != ATTESTATION_FIXED_SIZE:
        raise ValueError("malformed Attestation container")
    return deserialize_bitlist(b[ATTESTATION_FIXED_SIZE:], limit), b[4:132], b[132:228]
