"""CPU oracle for the hot path  process_attestation -> bls.Aggregate / FastAggregateVerify
-> get_head / get_weight  of /root/reference/pos-evolution.md.

THIS PACKAGE IS TEST INFRASTRUCTURE.  It is a pure-Python / numpy restatement of the
algorithms the reference quotes (pos-evolution.md, line numbers cited per function) and of
the third-party arithmetic the reference calls but does not contain (eth2spec.utils.bls ->
py_ecc G2ProofOfPossession; IETF draft-irtf-cfrg-bls-signature-05; RFC 9380).  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker or as the timed CPU
baseline.  The product package ``pos_evolution_b200`` never imports it and has no CPU
fallback.

Parity pinning: the reference ships no tests or vectors for this path.
The oracle is pinned instead by the public known-answer vectors of the standards it
restates (RFC 9380 K.1 / J.10.1, the eth2 ``bls/sign`` vectors, SkToPk vector, ZCash
generator encodings) -- tests/test_oracle_kat.py -- and by executing the reference's own
fenced python blocks (loaded from pos-evolution.md by line range) against it --
tests/test_oracle_literal_spec.py.  With respect to a *running* py_ecc/eth2spec the
parity is "unpinned": neither is installable in this sandbox.
"""
