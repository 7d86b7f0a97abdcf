#!/usr/bin/env python3
"""Generates pos_evolution_b200/csrc/consts.cuh: every BLS12-381 constant the CUDA library needs,
as 32-bit limbs (little-endian limb order), field elements in Montgomery form (R = 2^384).

Self-contained on purpose (plain Python ints, no import of oracle/): run at development time,
output committed.  tests/test_consts.py re-derives the table and checks it against the oracle.
"""
import os
import sys

P = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
R_ORDER = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
X_ABS = 0xd201000000010000
RMONT = 1 << 384


def f2mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def f2pow(a, e):
    r = (1, 0)
    while e:
        if e & 1:
            r = f2mul(r, a)
        a = f2mul(a, a)
        e >>= 1
    return r


def f2inv(a):
    d = pow(a[0] * a[0] + a[1] * a[1], -1, P)
    return (a[0] * d % P, -a[1] * d % P)


def f2conj(a):
    return (a[0], -a[1] % P)


def limbs(v, n=12):
    return [(v >> (32 * i)) & 0xFFFFFFFF for i in range(n)]


def mont(v):
    return v * RMONT % P


XI = (1, 1)
G1X = 0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb
G1Y = 0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1
G2X = (0x024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8,
       0x13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e)
G2Y = (0x0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c923ac9cc3baca289e193548608b82801,
       0x0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab3f370d275cec1da1aaa9075ff05f79be)

_K = 0x5c759507e8e333ebb5b7a9a47d7ed8532c52d39fd3a042a88b58423c50ae15d5c2638e343d9c71c6238aaaaaaaa97d6
_L = 0x1530477c7ab4113b59a4c18b076d11930f7da5d4a07f649bf54439d87d27e500fc8c25ebf8c92f6812cfc71c71c6d706
ISO_XNUM = [
    (_K, _K),
    (0, 0x11560bf17baa99bc32126fced787c88f984f87adf7ae0c7f9a208c6b4f20a4181472aaa9cb8d555526a9ffffffffc71a),
    (0x11560bf17baa99bc32126fced787c88f984f87adf7ae0c7f9a208c6b4f20a4181472aaa9cb8d555526a9ffffffffc71e,
     0x8ab05f8bdd54cde190937e76bc3e447cc27c3d6fbd7063fcd104635a790520c0a395554e5c6aaaa9354ffffffffe38d),
    (0x171d6541fa38ccfaed6dea691f5fb614cb14b4e7f4e810aa22d6108f142b85757098e38d0f671c7188e2aaaaaaaa5ed1, 0),
]
ISO_XDEN = [(0, P - 72), (12, P - 12)]                     # + monic x^2
ISO_YNUM = [
    (_L, _L),
    (0, 0x5c759507e8e333ebb5b7a9a47d7ed8532c52d39fd3a042a88b58423c50ae15d5c2638e343d9c71c6238aaaaaaaa97be),
    (0x11560bf17baa99bc32126fced787c88f984f87adf7ae0c7f9a208c6b4f20a4181472aaa9cb8d555526a9ffffffffc71c,
     0x8ab05f8bdd54cde190937e76bc3e447cc27c3d6fbd7063fcd104635a790520c0a395554e5c6aaaa9354ffffffffe38f),
    (0x124c9ad43b6cf79bfbf7043de3811ad0761b0f37a1e26286b0e977c69aa274524e79097a56dc4bd9e1b371c71c718b10, 0),
]
ISO_YDEN = [(P - 432, P - 432), (0, P - 216), (18, P - 18)]  # + monic x^3


def sliding_window_program(e, w=4):
    """Constant-exponent sliding-window schedule: list of (n_squarings, table_index) with table[k] = a^(2k+1);
    index 255 = squarings only.  First entry's squarings are 0 (the accumulator is initialised from the table)."""
    bits = bin(e)[2:]
    ops, i, pending, first = [], 0, 0, True
    n = len(bits)
    while i < n:
        if bits[i] == "0":
            pending += 1
            i += 1
            continue
        j = min(n, i + w)
        while bits[j - 1] == "0":
            j -= 1
        val = int(bits[i:j], 2)
        ops.append((0 if first else pending + (j - i), (val - 1) // 2))
        first, pending, i = False, 0, j
    if pending:
        ops.append((pending, 255))
    # self-check against pow()
    a = 0x1234567
    tbl = [pow(a, 2 * k + 1, P) for k in range(1 << (w - 1))]
    r = tbl[ops[0][1]]
    for nsq, idx in ops[1:]:
        for _ in range(nsq):
            r = r * r % P
        if idx != 255:
            r = r * tbl[idx] % P
    assert r == pow(a, e, P)
    return ops


RUN_TOKEN = 255         # a^(2^8 - 1): table entry 8 of fp_pow_prog (built from a^15 with 4 squarings + 1 multiplication)


def token_program(e, w=4):
    """Constant-exponent schedule over the token set {odd values < 2^w} + {RUN_TOKEN}, parsed optimally (fewest multiplications) by
    dynamic programming over the bit string.  (p-3)/4 has runs of 33, 19, 17, 10, 8 and 6 one-bits: a window-4 parse spends one
    multiplication per four of them, the a^255 entry one per eight -- 68 instead of 78 multiplications for 5 extra operations.
    Same output format as sliding_window_program; table index 8 = RUN_TOKEN."""
    bits = bin(e)[2:]
    n = len(bits)
    toks = {bin(v)[2:]: (v - 1) // 2 for v in range(1, 1 << w, 2)}
    toks[bin(RUN_TOKEN)[2:]] = 1 << (w - 1)
    INF = 10**9
    best, choice = [INF] * (n + 1), [None] * (n + 1)
    best[n] = 0
    for i in range(n - 1, -1, -1):
        if bits[i] == "0":
            best[i] = best[i + 1]
            continue
        for pat in toks:
            if bits.startswith(pat, i) and 1 + best[i + len(pat)] < best[i]:
                best[i], choice[i] = 1 + best[i + len(pat)], pat
    ops, i, pending, first = [], 0, 0, True
    while i < n:
        if choice[i] is None:
            pending += 1
            i += 1
            continue
        pat = choice[i]
        ops.append((0 if first else pending + len(pat), toks[pat]))
        first, pending, i = False, 0, i + len(pat)
    if pending:
        ops.append((pending, 255))
    a = 0x1234567
    tbl = [pow(a, 2 * k + 1, P) for k in range(1 << (w - 1))] + [pow(a, RUN_TOKEN, P)]
    r = tbl[ops[0][1]]
    for nsq, idx in ops[1:]:
        for _ in range(nsq):
            r = r * r % P
        if idx != 255:
            r = r * tbl[idx] % P
    assert r == pow(a, e, P)
    return ops


def build():
    """-> ordered list of (name, kind, value); kind in {'fp','fp2','raw'}; raw = not Montgomery."""
    c = []
    fp = lambda n, v: c.append((n, "fp", v % P))            # noqa: E731
    fp2 = lambda n, v: c.append((n, "fp2", (v[0] % P, v[1] % P)))  # noqa: E731
    raw = lambda n, v: c.append((n, "raw", v))              # noqa: E731

    raw("C_P", P)
    raw("C_R2", RMONT * RMONT % P)                          # to_mont(a) = mont_mul(a, R2)
    raw("C_HALF_P", (P - 1) // 2)                           # compared against canonical (non-Montgomery) values
    raw("C_EXP_PM3D4", (P - 3) // 4)                        # a^((p-3)/4): inverse-sqrt / sqrt building block
    raw("C_EXP_PM2", P - 2)                                 # Fermat inverse
    raw("C_R_ORDER", R_ORDER)                               # subgroup order, 255 bits (top limbs zero)
    for name, e in (("C_PROG_PM3D4", (P - 3) // 4), ("C_PROG_PM2", P - 2)):
        ops = token_program(e, 4)
        assert len(ops) <= len(sliding_window_program(e, 4))
        c.append((name, "words", [len(ops)] + [(nsq << 8) | idx for nsq, idx in ops]))
    fp("C_ONE", 1)
    fp("C_TWO_INV", pow(2, -1, P))
    fp("C_TWO256", 1 << 256)                                # hash_to_field: (hi*2^256 + lo) mod p
    fp("C_FOUR", 4)                                         # E1: b
    fp2("C_B2", (4, 4))                                     # E2: b' = 4*xi
    fp("C_G1X", G1X)
    fp("C_G1Y", G1Y)
    fp("C_G1Y_NEG", P - G1Y)
    fp2("C_G2X", G2X)
    fp2("C_G2Y", G2Y)
    # Frobenius on Fp12 = Fp2[w]/(w^6 - xi): coefficient k of w^k picks up gamma^k, gamma = xi^((p-1)/6)
    g = f2pow(XI, (P - 1) // 6)
    gk = (1, 0)
    for k in range(1, 6):
        gk = f2mul(gk, g)
        fp2("C_FROB1_%d" % k, gk)
    gk = (1, 0)
    for k in range(1, 6):
        gk = f2mul(gk, g)
        n = f2mul(gk, f2conj(gk))
        assert n[1] == 0
        fp("C_FROB2_%d" % k, n[0])                          # p^2-Frobenius coefficient (in Fp)
    # psi endomorphism on E2 (RFC 9380 appendix G.3)
    fp2("C_PSI_CX", f2inv(f2pow(XI, (P - 1) // 3)))
    fp2("C_PSI_CY", f2inv(f2pow(XI, (P - 1) // 2)))
    fp("C_PSI2_CX", pow(pow(2, (P - 1) // 3, P), -1, P))
    # SSWU on E2': y^2 = x^3 + A x + B
    A, B, Z = (0, 240), (1012, 1012), (P - 2, P - 1)
    fp2("C_SSWU_A", A)
    fp2("C_SSWU_B", B)
    fp2("C_SSWU_Z", Z)
    fp2("C_SSWU_MB_OVER_A", f2mul(((-B[0]) % P, (-B[1]) % P), f2inv(A)))      # -B/A
    fp2("C_SSWU_B_OVER_ZA", f2mul(B, f2inv(f2mul(Z, A))))                      # x1 when the denominator vanishes
    zn = (Z[0] * Z[0] + Z[1] * Z[1]) % P                                       # norm(Z) = 5, a non-residue
    assert pow(zn, (P - 1) // 2, P) == P - 1
    zeta = pow(-zn % P, (P + 1) // 4, P)
    assert zeta * zeta % P == -zn % P
    fp("C_SSWU_ZNORM", zn)
    fp("C_SSWU_ZETA", zeta)                                                    # sqrt(-norm(Z))
    for i, v in enumerate(ISO_XNUM):
        fp2("C_ISO_XNUM%d" % i, v)
    for i, v in enumerate(ISO_XDEN):
        fp2("C_ISO_XDEN%d" % i, v)
    for i, v in enumerate(ISO_YNUM):
        fp2("C_ISO_YNUM%d" % i, v)
    for i, v in enumerate(ISO_YDEN):
        fp2("C_ISO_YDEN%d" % i, v)
    return c


def table():
    words, offsets = [], []
    for name, kind, v in build():
        offsets.append((name, len(words)))
        if kind == "words":
            words += list(v)
        elif kind == "raw":
            words += limbs(v)
        elif kind == "fp":
            words += limbs(mont(v))
        else:
            words += limbs(mont(v[0])) + limbs(mont(v[1]))
    return words, offsets


def render():
    words, offsets = table()
    inv = (-pow(P, -1, 1 << 32)) % (1 << 32)
    out = []
    out.append("// consts.cuh -- GENERATED by tools/gen_consts.py; do not edit.")
    out.append("// BLS12-381 constants as little-endian 32-bit limbs; field elements in Montgomery form (R = 2^384).")
    out.append("#pragma once")
    out.append('#include "platform.cuh"')
    out.append("namespace b2 {")
    out.append("#define B2_MONT_INV 0x%08xu   /* -p^-1 mod 2^32 */" % inv)
    out.append("#define B2_X_ABS 0x%016xull   /* |x|, curve parameter x = -|x| */" % X_ABS)
    out.append("HD constexpr uint32_t P_LIMB(int i) {")
    out.append("    constexpr uint32_t t[12] = {%s};" % ", ".join("0x%08xu" % w for w in limbs(P)))
    out.append("    return t[i];")
    out.append("}")
    out.append("enum ConstOffset : int {")
    for name, off in offsets:
        out.append("    %s = %d," % (name, off))
    out.append("    C_TABLE_WORDS = %d" % len(words))
    out.append("};")
    body = ",\n    ".join(", ".join("0x%08xu" % w for w in words[i:i + 6]) for i in range(0, len(words), 6))
    out.append("#define B2_CONST_WORDS \\\n    " + body.replace("\n", " \\\n"))
    out.append("#if defined(__CUDACC__)")
    out.append("__device__ __constant__ uint32_t d_const_table[C_TABLE_WORDS] = { B2_CONST_WORDS };")
    out.append("#endif")
    out.append("static const uint32_t h_const_table[C_TABLE_WORDS] = { B2_CONST_WORDS };")
    out.append("HD const uint32_t* const_table() {")
    out.append("#if defined(__CUDA_ARCH__)")
    out.append("    return d_const_table;")
    out.append("#else")
    out.append("    return h_const_table;")
    out.append("#endif")
    out.append("}")
    out.append("}  // namespace b2")
    return "\n".join(out) + "\n"


if __name__ == "__main__":
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "pos_evolution_b200", "csrc", "consts.cuh")
    txt = render()
    if len(sys.argv) > 1 and sys.argv[1] == "--check":
        sys.exit(0 if open(dst).read() == txt else 1)
    with open(dst, "w") as f:
        f.write(txt)
    print("wrote", os.path.normpath(dst), len(txt), "bytes")
