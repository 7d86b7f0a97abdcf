"""GPU parity of the participation-flag / proposer-reward kernels (process_attestation :745-752)
against the sequential rule, on a batch whose attestations overlap heavily (the same validator earns the same flag in
several attestations: only the first in list order may be credited)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

WEIGHTS = (14, 26, 14)


def test_participation_update_order_exact():
    from oracle.bls12_381 import G1, g1_compress
    from pos_evolution_b200.engine import Engine
    rng = np.random.default_rng(31)
    n_val, n_agg, csize = 2048, 200, 64
    eng = Engine(0)
    eff = (rng.integers(1, 33, size=n_val).astype(np.uint64)) * np.uint64(10**9) + rng.integers(0, 10**9, size=n_val).astype(np.uint64)
    eng.registry_load(np.tile(np.frombuffer(g1_compress(G1), dtype=np.uint8), (n_val, 1)), eff)
    inc, per_inc = 10**9, 7777
    for which in (0, 1):
        part = rng.integers(0, 8, size=n_val).astype(np.uint8) * (rng.random(n_val) < 0.3)
        part = part.astype(np.uint8)
        members = np.stack([rng.permutation(n_val)[:csize] for _ in range(n_agg)]).astype(np.uint32)
        off = np.arange(0, (n_agg + 1) * csize, csize, dtype=np.uint32)
        bits = rng.integers(0, 256, size=(n_agg, csize // 8)).astype(np.uint8)
        masks = rng.integers(0, 8, size=n_agg).astype(np.uint8)
        accept = (rng.random(n_agg) < 0.85).astype(np.uint8)
        ref_part = part.copy()
        ref_num = np.zeros(n_agg, dtype=object)
        for a in range(n_agg):
            if not accept[a]:
                continue
            for j in range(csize):
                if (bits[a, j >> 3] >> (j & 7)) & 1:
                    v = int(members[a, j])
                    for f in range(3):
                        if (masks[a] >> f) & 1 and not (ref_part[v] >> f) & 1:
                            ref_part[v] |= 1 << f
                            ref_num[a] += (int(eff[v]) // inc) * per_inc * WEIGHTS[f]
        eng.participation_load(which, part)
        num = eng.participation_update(which, members.reshape(-1), off, bits, masks, accept, inc, per_inc)
        assert [int(x) for x in num] == [int(x) for x in ref_num]
        assert np.array_equal(eng.participation_read(which), ref_part)
        # a second, identical batch earns nothing and changes nothing (flags already set)
        num2 = eng.participation_update(which, members.reshape(-1), off, bits, masks, accept, inc, per_inc)
        assert int(num2.sum()) == 0 and np.array_equal(eng.participation_read(which), ref_part)
    eng.close()


@pytest.mark.parametrize("n_val,seed", [(64, 1), (5000, 2), (1 << 20, 3)])
def test_ffg_balances_match_numpy(n_val, seed):
    """b2_ffg_balances (process_justification_and_finalization's three sums, pos-evolution.md:793-803) against numpy, including a
    table that was never loaded (counts as empty) and all eight flag positions."""
    from oracle.bls12_381 import G1, g1_compress
    from pos_evolution_b200.engine import Engine
    rng = np.random.default_rng(seed)
    eng = Engine(0)
    eff = rng.integers(16, 33, size=n_val).astype(np.uint64) * np.uint64(10**9)
    flags = (rng.random(n_val) < 0.9).astype(np.uint8) | ((rng.random(n_val) < 0.05).astype(np.uint8) << 1) | ((rng.random(n_val) < 0.9).astype(np.uint8) << 2)
    eng.registry_load(np.tile(np.frombuffer(g1_compress(G1), dtype=np.uint8), (n_val, 1)), eff, flags)
    cur = rng.integers(0, 256, size=n_val).astype(np.uint8)
    prev = rng.integers(0, 256, size=n_val).astype(np.uint8)

    def want(flag, cur_t, prev_t):
        act, sl, actp = (flags & 1) != 0, (flags & 2) != 0, (flags & 4) != 0
        s = lambda m: int(eff[m].astype(object).sum()) if m.any() else 0     # noqa: E731
        return (s(act), s(act & ~sl & (((cur_t >> flag) & 1) != 0)), s(actp & ~sl & (((prev_t >> flag) & 1) != 0)), s(act & ~sl))

    zero = np.zeros(n_val, dtype=np.uint8)
    assert eng.ffg_balances(1) == want(1, zero, zero)                        # nothing loaded yet
    eng.participation_load(0, cur)
    assert eng.ffg_balances(1) == want(1, cur, zero)
    eng.participation_load(1, prev)
    for flag in range(8):
        assert eng.ffg_balances(flag) == want(flag, cur, prev)
    eng.close()
