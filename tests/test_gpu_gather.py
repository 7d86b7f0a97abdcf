"""The TMA-staged pubkey gather (csrc/gather.cuh) and the device-side input guards of the *_dev entry points.

  * the gather stage alone (b2_gather_probe_dev): the TMA form (cp.async.bulk -> shared memory) and the plain-load form must
    fetch the same record words for every shape of committee row -- aligned, unaligned, partial last word, several 512-member
    chunks, empty rows, sparse bits;
  * K2 through the TMA kernel == K2 through the LDG kernel (B2_K2_TMA=0) == the oracle's G1 sums (tests/test_gpu_bls.py and
    tests/test_gpu_fullsize.py compare the same entry point with the oracle byte for byte);
  * guards: a member index outside the registry, a committee row whose offsets run backwards, an epoch that does not fit the
    32-bit key -- nothing is read out of bounds, the aggregate fails / is skipped, b2_guard_flags reports it."""
import os

import numpy as np
import pytest

import scenarios
from oracle.bls12_381 import E1, g1_compress, g1_decompress

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

N = 3000


@pytest.fixture(scope="module")
def world():
    from pos_evolution_b200.engine import Engine
    eng = Engine(0)
    pks = scenarios.pubkeys(64)
    # 3000 registry rows made of 64 distinct keys; rows 7 and 1999 invalid (undecodable / infinity)
    rows = [pks[(i * 37) % 64] for i in range(N)]
    rows[7] = bytes(48)
    rows[1999] = bytes([0xC0]) + bytes(47)
    pk = np.frombuffer(b"".join(rows), dtype=np.uint8).reshape(N, 48)
    valid = eng.registry_load(pk, np.full(N, 32 * 10**9, dtype=np.uint64))
    assert valid.sum() == N - 2
    yield eng, rows
    eng.close()


def _rows(rng, sizes, stride, lead=0):
    members = np.concatenate([rng.integers(0, N, size=lead)] + [rng.integers(0, N, size=s) for s in sizes]).astype(np.uint32)
    off = (lead + np.concatenate([[0], np.cumsum(sizes)])).astype(np.uint32)
    bits = rng.integers(0, 256, size=(len(sizes), stride), dtype=np.uint8)
    return members, off, bits


@pytest.mark.parametrize("sizes,stride,lead", [
    ([512] * 8, 64, 0),                         # the mainnet shape: 16-byte aligned rows, one chunk
    ([512, 511, 513, 1, 0, 33, 2048, 700], 256, 0),      # ragged; 2048 = four chunks; an empty row
    ([100, 37, 64], 13, 3),                     # odd stride and a 3-element lead: nothing is 16-byte aligned
    ([1200] * 3, 150, 1),
])
def test_gather_probe_tma_equals_ldg(world, sizes, stride, lead):
    eng, _ = world
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(len(sizes) * 1000 + stride)
    members, off, bits = _rows(rng, sizes, stride, lead)
    off[0] = lead
    n = len(sizes)
    d_m = torch.as_tensor(members.astype(np.int32), device=dev)
    d_o = torch.as_tensor(off.astype(np.int32), device=dev)
    d_b = torch.as_tensor(bits, device=dev)
    c1 = torch.zeros(n, dtype=torch.int32, device=dev)
    c0 = torch.zeros(n, dtype=torch.int32, device=dev)
    eng.gather_probe_dev(d_m, d_o, d_b, c1, tma=True)
    eng.gather_probe_dev(d_m, d_o, d_b, c0, tma=False)
    torch.cuda.synchronize()
    assert torch.equal(c0, c1)
    assert eng.guard_flags() == 0
    # sparse rows (few set bits) and all-clear rows
    bits[:] = 0
    bits[0, 0] = 1
    d_b = torch.as_tensor(bits, device=dev)
    eng.gather_probe_dev(d_m, d_o, d_b, c1, tma=True)
    eng.gather_probe_dev(d_m, d_o, d_b, c0, tma=False)
    torch.cuda.synchronize()
    assert torch.equal(c0, c1) and int(c1[1:].abs().sum()) == 0


def _oracle_sum(rows, members, off, bits, a):
    acc, bad, cnt = E1.INF, False, 0
    for j in range(int(off[a + 1] - off[a])):
        if (bits[a, j >> 3] >> (j & 7)) & 1:
            cnt += 1
            v = int(members[off[a] + j])
            if v in (7, 1999):
                bad = True
            else:
                acc = E1.add(acc, g1_decompress(rows[v]))
    return acc, bad, cnt


def test_k2_tma_vs_oracle_and_ldg_form(world):
    """b2_g1_aggregate (TMA gather by default) against the oracle's sums, incl. invalid keys and empty selections, and against a
    second context forced to the LDG kernel."""
    eng, rows = world
    rng = np.random.default_rng(5)
    sizes = [512, 130, 1, 40, 600, 0, 64]
    members, off, bits = _rows(rng, sizes, 80)
    members[off[3]:off[3] + 40] = rng.choice([i for i in range(N) if i not in (7, 1999)], size=40)      # a row without invalid keys
    bits[6] = 0                                                                                          # nothing selected
    out, status = eng.g1_aggregate(members, off, bits)
    for a in range(len(sizes)):
        acc, bad, cnt = _oracle_sum(rows, members, off, bits, a)
        want_status = (1 if bad else 0) | (2 if cnt == 0 else 0)
        if not bad:
            want_status |= 4 if E1.is_inf(acc) else 0
            assert bytes(out[a]) == g1_compress(acc), a
        assert int(status[a]) & 3 == want_status & 3, a
    from pos_evolution_b200.engine import Engine
    os.environ["B2_K2_TMA"] = "0"
    try:
        eng2 = Engine(0)
    finally:
        del os.environ["B2_K2_TMA"]
    pk = np.frombuffer(b"".join(rows), dtype=np.uint8).reshape(N, 48)
    eng2.registry_load(pk, np.full(N, 32 * 10**9, dtype=np.uint64))
    out2, status2 = eng2.g1_aggregate(members, off, bits)
    eng2.close()
    ok_rows = [a for a in range(len(sizes)) if not (int(status[a]) & 1)]
    assert np.array_equal(status, status2) and np.array_equal(out[ok_rows], out2[ok_rows])


def test_dev_entry_points_guard_their_inputs(world):
    eng, rows = world
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(9)
    members, off, bits = _rows(rng, [64, 64, 64, 64], 8)
    bits[:] = 0xFF
    members[70] = 0xFFFFFFF0                       # row 1: an index far outside the registry
    off_bad = off.copy()
    off_bad[3] = off[2] - 5                        # row 2 runs backwards (and row 3 becomes longer than the bit row)
    d_m = torch.as_tensor(members.astype(np.int64).astype(np.int32), device=dev)
    d_b = torch.as_tensor(bits, device=dev)
    msgs = torch.zeros((4, 32), dtype=torch.uint8, device=dev)
    sigs = torch.zeros((4, 96), dtype=torch.uint8, device=dev)
    sigs[:, 0] = 0xC0
    ok = torch.full((4,), 7, dtype=torch.uint8, device=dev)
    assert eng.guard_flags() == 0
    for o, want_flags in ((off, 1), (off_bad, 1 | 2)):
        d_o = torch.as_tensor(o.astype(np.int32), device=dev)
        eng.fast_aggregate_verify_dev(d_m, d_o, d_b, msgs, sigs, ok)
        torch.cuda.synchronize()
        assert ok.tolist() == [0, 0, 0, 0]         # (infinity signatures never verify against non-infinity keys)
        assert eng.guard_flags() == want_flags
    # LMD update: the bad index is skipped, the bad rows are skipped, a 2^32 epoch is skipped; everything else lands
    eng.latest_messages_reset()
    d_o = torch.as_tensor(off_bad.astype(np.int32), device=dev)
    te = torch.tensor([5, 6, 7, 8], dtype=torch.int64, device=dev)
    blk = torch.tensor([1, 2, 3, 4], dtype=torch.int32, device=dev)
    eng.tree_load(np.array([0, 0, 1, 2, 3], dtype=np.uint32), np.arange(5, dtype=np.uint64), np.arange(160, dtype=np.uint8).reshape(5, 32), np.ones(5, dtype=np.uint8))
    eng.latest_messages_update_dev(d_m, d_o, d_b, te, blk, None)
    assert eng.guard_flags() == (1 | 2)
    e, b, h = eng.latest_messages_read()
    touched = set(int(v) for v in members[:64]) | set(int(v) for v in members[64:128] if v < N)
    assert set(np.nonzero(h)[0].tolist()) == touched
    te[0] = 0xFFFFFFFF
    eng.latest_messages_reset()
    eng.latest_messages_update_dev(d_m, torch.as_tensor(off.astype(np.int32), device=dev), d_b, te, blk, None)
    assert eng.guard_flags() == (1 | 4)
    e, b, h = eng.latest_messages_read()
    assert not h[[int(v) for v in members[:64] if int(v) not in set(int(x) for x in members[64:]) ]].any()
