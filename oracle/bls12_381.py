"""BLS12-381 arithmetic -- CPU oracle (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates, with Python big ints, the arithmetic that the reference's ``bls.*`` call sites
(/root/reference/pos-evolution.md:165 ``bls.Verify``; ``is_valid_indexed_attestation`` at
:736, :976, :1456-1457) delegate to eth2spec.utils.bls -> py_ecc (un-vendored, unpinned;
SURVEY.md section 8c).  Conventions follow py_ecc / ZCash / RFC 9380:

  Fp2  = Fp[i]/(i^2+1)                    element = (c0, c1)
  Fp6  = Fp2[v]/(v^3 - xi),  xi = 1+i     element = (a0, a1, a2)
  Fp12 = Fp6[w]/(w^2 - v)                 element = (b0, b1)
  E1: y^2 = x^3 + 4        E2: y^2 = x^3 + 4*xi  (M-type twist)
  points are Jacobian triples (X, Y, Z); infinity <=> Z == 0.

The pairing is the optimal ate pairing with loop parameter |x| = 0xd201000000010000 and a
final exponentiation to the power 3*(p^12-1)/r (Hayashida-Hayasaka-Teruya chain); the cube
does not change any ``== 1`` verdict because gcd(3, r) = 1.
"""

P = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
X_ABS = 0xd201000000010000          # the curve parameter is x = -X_ABS
H1 = 0x396c8c005555e1568c00aaab0000aaab
H_EFF_G2 = 0xbc69f08f2ee75b3584c6a0ea91b352888e2a8e9145ad7689986ff031508ffe1329c2f178731db956d82bf015d1212b02ec0ec69d7477c1ae954cbc06689f6a359894c0adebbf6b4e8020005aaa95551

G1_X = 0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb
G1_Y = 0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1
G2_X = (0x024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8,
        0x13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e)
G2_Y = (0x0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c923ac9cc3baca289e193548608b82801,
        0x0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab3f370d275cec1da1aaa9075ff05f79be)

# ----------------------------------------------------------------------------- Fp
def fp_inv(a):
    return pow(a, -1, P)


def fp_sqrt(a):
    """Square root in Fp (p = 3 mod 4) or None."""
    a %= P
    s = pow(a, (P + 1) // 4, P)
    return s if s * s % P == a else None


# ----------------------------------------------------------------------------- Fp2
F2_ZERO = (0, 0)
F2_ONE = (1, 0)
XI = (1, 1)


def f2_add(a, b):
    return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)


def f2_sub(a, b):
    return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)


def f2_neg(a):
    return (-a[0] % P, -a[1] % P)


def f2_mul(a, b):
    t0 = a[0] * b[0]
    t1 = a[1] * b[1]
    return ((t0 - t1) % P, ((a[0] + a[1]) * (b[0] + b[1]) - t0 - t1) % P)


def f2_sqr(a):
    return ((a[0] + a[1]) * (a[0] - a[1]) % P, 2 * a[0] * a[1] % P)


def f2_muls(a, k):
    """Multiply by an Fp scalar."""
    return (a[0] * k % P, a[1] * k % P)


def f2_conj(a):
    return (a[0], -a[1] % P)


def f2_inv(a):
    d = pow(a[0] * a[0] + a[1] * a[1], -1, P)
    return (a[0] * d % P, -a[1] * d % P)


def f2_mul_xi(a):
    return ((a[0] - a[1]) % P, (a[0] + a[1]) % P)


def f2_pow(a, e):
    r = F2_ONE
    for bit in bin(e)[2:]:
        r = f2_sqr(r)
        if bit == "1":
            r = f2_mul(r, a)
    return r


def f2_is_zero(a):
    return a[0] % P == 0 and a[1] % P == 0


def f2_sqrt(a):
    """Some square root in Fp2, or None.  ("complex method", p = 3 mod 4)."""
    a0, a1 = a[0] % P, a[1] % P
    if a1 == 0:
        s = fp_sqrt(a0)
        if s is not None:
            return (s, 0)
        s = fp_sqrt(-a0 % P)          # sqrt(a0) = i*sqrt(-a0)
        return (0, s)                 # -a0 is a QR whenever a0 is not (p = 3 mod 4)
    n = fp_sqrt((a0 * a0 + a1 * a1) % P)
    if n is None:
        return None
    inv2 = (P + 1) // 2
    t = (a0 + n) * inv2 % P
    x0 = fp_sqrt(t)
    if x0 is None:
        t = (a0 - n) * inv2 % P
        x0 = fp_sqrt(t)
        if x0 is None:
            return None
    x1 = a1 * fp_inv(2 * x0 % P) % P
    r = (x0, x1)
    return r if f2_sqr(r) == (a0, a1) else None


# ----------------------------------------------------------------------------- Fp6
F6_ZERO = (F2_ZERO, F2_ZERO, F2_ZERO)
F6_ONE = (F2_ONE, F2_ZERO, F2_ZERO)


def f6_add(a, b):
    return (f2_add(a[0], b[0]), f2_add(a[1], b[1]), f2_add(a[2], b[2]))


def f6_sub(a, b):
    return (f2_sub(a[0], b[0]), f2_sub(a[1], b[1]), f2_sub(a[2], b[2]))


def f6_neg(a):
    return (f2_neg(a[0]), f2_neg(a[1]), f2_neg(a[2]))


def f6_mul(a, b):
    a0, a1, a2 = a
    b0, b1, b2 = b
    c0 = f2_add(f2_mul(a0, b0), f2_mul_xi(f2_add(f2_mul(a1, b2), f2_mul(a2, b1))))
    c1 = f2_add(f2_add(f2_mul(a0, b1), f2_mul(a1, b0)), f2_mul_xi(f2_mul(a2, b2)))
    c2 = f2_add(f2_add(f2_mul(a0, b2), f2_mul(a1, b1)), f2_mul(a2, b0))
    return (c0, c1, c2)


def f6_mul_v(a):
    return (f2_mul_xi(a[2]), a[0], a[1])


def f6_inv(a):
    a0, a1, a2 = a
    t0 = f2_sub(f2_sqr(a0), f2_mul_xi(f2_mul(a1, a2)))
    t1 = f2_sub(f2_mul_xi(f2_sqr(a2)), f2_mul(a0, a1))
    t2 = f2_sub(f2_sqr(a1), f2_mul(a0, a2))
    d = f2_add(f2_mul(a0, t0), f2_mul_xi(f2_add(f2_mul(a2, t1), f2_mul(a1, t2))))
    di = f2_inv(d)
    return (f2_mul(t0, di), f2_mul(t1, di), f2_mul(t2, di))


# ----------------------------------------------------------------------------- Fp12
F12_ONE = (F6_ONE, F6_ZERO)


def f12_mul(a, b):
    t0 = f6_mul(a[0], b[0])
    t1 = f6_mul(a[1], b[1])
    c0 = f6_add(t0, f6_mul_v(t1))
    c1 = f6_add(f6_mul(a[0], b[1]), f6_mul(a[1], b[0]))
    return (c0, c1)


def f12_sqr(a):
    return f12_mul(a, a)


def f12_conj(a):
    return (a[0], f6_neg(a[1]))


def f12_inv(a):
    d = f6_sub(f6_mul(a[0], a[0]), f6_mul_v(f6_mul(a[1], a[1])))
    di = f6_inv(d)
    return (f6_mul(a[0], di), f6_neg(f6_mul(a[1], di)))


def f12_pow(a, e):
    r = F12_ONE
    for bit in bin(e)[2:]:
        r = f12_sqr(r)
        if bit == "1":
            r = f12_mul(r, a)
    return r


def f12_to_w(a):
    """Tower element -> coefficients of w^0..w^5 (each in Fp2)."""
    (c0, c1, c2), (d0, d1, d2) = a
    return [c0, d0, c1, d1, c2, d2]


def f12_from_w(g):
    return ((g[0], g[2], g[4]), (g[1], g[3], g[5]))


# w^(p-1) = xi^((p-1)/6); Frobenius acts on g_k*w^k as conj(g_k) * GAMMA1^k * w^k
GAMMA1 = f2_pow(XI, (P - 1) // 6)
GAMMA1_POW = [F2_ONE]
for _k in range(5):
    GAMMA1_POW.append(f2_mul(GAMMA1_POW[-1], GAMMA1))


def f12_frob(a):
    g = f12_to_w(a)
    return f12_from_w([f2_mul(f2_conj(g[k]), GAMMA1_POW[k]) for k in range(6)])


def f12_frob_n(a, n):
    for _ in range(n):
        a = f12_frob(a)
    return a


# ----------------------------------------------------------------------------- curves
class _Curve:
    """Short Weierstrass y^2 = x^3 + b, a = 0, Jacobian coordinates, generic field."""

    def __init__(self, add, sub, mul, sqr, neg, inv, is_zero, zero, one, b):
        self.fadd, self.fsub, self.fmul, self.fsqr = add, sub, mul, sqr
        self.fneg, self.finv, self.fis_zero = neg, inv, is_zero
        self.zero, self.one, self.b = zero, one, b
        self.INF = (one, one, zero)

    def is_inf(self, p):
        return self.fis_zero(p[2])

    def from_affine(self, x, y):
        return (x, y, self.one)

    def to_affine(self, p):
        """-> (x, y) or None for infinity."""
        if self.is_inf(p):
            return None
        zi = self.finv(p[2])
        zi2 = self.fsqr(zi)
        return (self.fmul(p[0], zi2), self.fmul(p[1], self.fmul(zi2, zi)))

    def on_curve_affine(self, x, y):
        return self.fsqr(y) == self.fadd(self.fmul(self.fsqr(x), x), self.b)

    def neg(self, p):
        return (p[0], self.fneg(p[1]), p[2])

    def dbl(self, p):
        X, Y, Z = p
        if self.is_inf(p):
            return p
        A = self.fsqr(X)
        B = self.fsqr(Y)
        C = self.fsqr(B)
        t = self.fsub(self.fsub(self.fsqr(self.fadd(X, B)), A), C)
        D = self.fadd(t, t)
        E = self.fadd(self.fadd(A, A), A)
        F = self.fsqr(E)
        X3 = self.fsub(F, self.fadd(D, D))
        C8 = self.fadd(C, C)
        C8 = self.fadd(C8, C8)
        C8 = self.fadd(C8, C8)
        Y3 = self.fsub(self.fmul(E, self.fsub(D, X3)), C8)
        YZ = self.fmul(Y, Z)
        Z3 = self.fadd(YZ, YZ)
        return (X3, Y3, Z3)

    def add(self, p, q):
        if self.is_inf(p):
            return q
        if self.is_inf(q):
            return p
        X1, Y1, Z1 = p
        X2, Y2, Z2 = q
        Z1Z1 = self.fsqr(Z1)
        Z2Z2 = self.fsqr(Z2)
        U1 = self.fmul(X1, Z2Z2)
        U2 = self.fmul(X2, Z1Z1)
        S1 = self.fmul(self.fmul(Y1, Z2), Z2Z2)
        S2 = self.fmul(self.fmul(Y2, Z1), Z1Z1)
        H = self.fsub(U2, U1)
        Rr = self.fsub(S2, S1)
        if self.fis_zero(H):
            if self.fis_zero(Rr):
                return self.dbl(p)
            return self.INF
        HH = self.fsqr(H)
        HHH = self.fmul(H, HH)
        V = self.fmul(U1, HH)
        X3 = self.fsub(self.fsub(self.fsqr(Rr), HHH), self.fadd(V, V))
        Y3 = self.fsub(self.fmul(Rr, self.fsub(V, X3)), self.fmul(S1, HHH))
        Z3 = self.fmul(self.fmul(Z1, Z2), H)
        return (X3, Y3, Z3)

    def mul(self, p, k):
        if k < 0:
            return self.mul(self.neg(p), -k)
        r = self.INF
        for bit in bin(k)[2:]:
            r = self.dbl(r)
            if bit == "1":
                r = self.add(r, p)
        return r

    def eq(self, p, q):
        if self.is_inf(p) or self.is_inf(q):
            return self.is_inf(p) and self.is_inf(q)
        Z1Z1 = self.fsqr(p[2])
        Z2Z2 = self.fsqr(q[2])
        if self.fmul(p[0], Z2Z2) != self.fmul(q[0], Z1Z1):
            return False
        return self.fmul(self.fmul(p[1], q[2]), Z2Z2) == self.fmul(self.fmul(q[1], p[2]), Z1Z1)

    def batch_to_affine(self, pts):
        """Montgomery batch inversion; infinity -> None."""
        acc = self.one
        pref = []
        for p in pts:
            pref.append(acc)
            if not self.is_inf(p):
                acc = self.fmul(acc, p[2])
        inv = self.finv(acc)
        out = [None] * len(pts)
        for i in range(len(pts) - 1, -1, -1):
            p = pts[i]
            if self.is_inf(p):
                continue
            zi = self.fmul(inv, pref[i])
            inv = self.fmul(inv, p[2])
            zi2 = self.fsqr(zi)
            out[i] = (self.fmul(p[0], zi2), self.fmul(p[1], self.fmul(zi2, zi)))
        return out


E1 = _Curve(lambda a, b: (a + b) % P, lambda a, b: (a - b) % P, lambda a, b: a * b % P,
            lambda a: a * a % P, lambda a: -a % P, fp_inv, lambda a: a % P == 0, 0, 1, 4)
E2 = _Curve(f2_add, f2_sub, f2_mul, f2_sqr, f2_neg, f2_inv, f2_is_zero, F2_ZERO, F2_ONE, (4, 4))

G1 = E1.from_affine(G1_X, G1_Y)
G2 = E2.from_affine(G2_X, G2_Y)


def g1_in_subgroup(p):
    return E1.is_inf(E1.mul(p, R))


def g2_in_subgroup(p):
    return E2.is_inf(E2.mul(p, R))


# ----------------------------------------------------------------------------- serialization (ZCash format)
class DeserializationError(ValueError):
    pass


HALF_P = (P - 1) // 2


def _f2_lex_larger(y):
    """'y is the lexicographically larger root' -- compare c1 first, then c0."""
    if y[1] != 0:
        return y[1] > HALF_P
    return y[0] > HALF_P


def g1_compress(p):
    aff = E1.to_affine(p)
    if aff is None:
        return bytes([0xC0]) + bytes(47)
    x, y = aff
    b = bytearray(x.to_bytes(48, "big"))
    b[0] |= 0x80 | (0x20 if y > HALF_P else 0)
    return bytes(b)


def g1_decompress(b):
    """48 bytes -> Jacobian point (on curve, subgroup NOT checked); raises DeserializationError."""
    if len(b) != 48:
        raise DeserializationError("G1: length")
    c, i, s = b[0] >> 7 & 1, b[0] >> 6 & 1, b[0] >> 5 & 1
    if not c:
        raise DeserializationError("G1: compression flag clear")
    x = int.from_bytes(b, "big") & ((1 << 381) - 1)
    if i:
        if s or x != 0:
            raise DeserializationError("G1: bad infinity encoding")
        return E1.INF
    if x >= P:
        raise DeserializationError("G1: x >= p")
    y = fp_sqrt((x * x * x + 4) % P)
    if y is None:
        raise DeserializationError("G1: not on curve")
    if (y > HALF_P) != bool(s):
        y = P - y
    return (x, y, 1)


def g2_compress(p):
    aff = E2.to_affine(p)
    if aff is None:
        return bytes([0xC0]) + bytes(95)
    x, y = aff
    b = bytearray(x[1].to_bytes(48, "big") + x[0].to_bytes(48, "big"))
    b[0] |= 0x80 | (0x20 if _f2_lex_larger(y) else 0)
    return bytes(b)


def g2_decompress(b):
    """96 bytes -> Jacobian point (on curve, subgroup NOT checked); raises DeserializationError."""
    if len(b) != 96:
        raise DeserializationError("G2: length")
    c, i, s = b[0] >> 7 & 1, b[0] >> 6 & 1, b[0] >> 5 & 1
    if not c:
        raise DeserializationError("G2: compression flag clear")
    x1 = int.from_bytes(b[:48], "big") & ((1 << 381) - 1)
    x0 = int.from_bytes(b[48:], "big")
    if i:
        if s or x1 != 0 or x0 != 0:
            raise DeserializationError("G2: bad infinity encoding")
        return E2.INF
    if x1 >= P or x0 >= P:
        raise DeserializationError("G2: x >= p")
    x = (x0, x1)
    y = f2_sqrt(f2_add(f2_mul(f2_sqr(x), x), (4, 4)))
    if y is None:
        raise DeserializationError("G2: not on curve")
    if _f2_lex_larger(y) != bool(s):
        y = f2_neg(y)
    return (x, y, F2_ONE)


# ----------------------------------------------------------------------------- pairing
def _line_dbl(T, Pa):
    """Affine doubling step on E2: returns (2T, line coefficients (c0, c1, d1)) for P = (xP, yP).

    Untwist (x', y') -> (x'/w^2, y'/w^3); the tangent at T evaluated at P and scaled by w^3 is
        (lam*xT - yT) + (-lam*xP) * w^2 + yP * w^3
    i.e. tower slots c0, c1 (coefficient of v) and d1 (coefficient of v*w).
    """
    xT, yT = T
    lam = f2_mul(f2_muls(f2_sqr(xT), 3), f2_inv(f2_add(yT, yT)))
    x3 = f2_sub(f2_sqr(lam), f2_add(xT, xT))
    y3 = f2_sub(f2_mul(lam, f2_sub(xT, x3)), yT)
    line = (f2_sub(f2_mul(lam, xT), yT), f2_muls(lam, -Pa[0] % P), (Pa[1], 0))
    return (x3, y3), line


def _line_add(T, Q, Pa):
    xT, yT = T
    xQ, yQ = Q
    lam = f2_mul(f2_sub(yQ, yT), f2_inv(f2_sub(xQ, xT)))
    x3 = f2_sub(f2_sub(f2_sqr(lam), xT), xQ)
    y3 = f2_sub(f2_mul(lam, f2_sub(xT, x3)), yT)
    line = (f2_sub(f2_mul(lam, xT), yT), f2_muls(lam, -Pa[0] % P), (Pa[1], 0))
    return (x3, y3), line


def _line_to_f12(line):
    c0, c1, d1 = line
    return ((c0, c1, F2_ZERO), (F2_ZERO, d1, F2_ZERO))


def miller_loop(Pa, Qa):
    """f_{|x|,Q}(P) conjugated (x < 0).  Pa: affine G1 (x, y) ints; Qa: affine G2.  None = infinity -> 1."""
    if Pa is None or Qa is None:
        return F12_ONE
    f = F12_ONE
    T = Qa
    for bit in bin(X_ABS)[3:]:
        T, line = _line_dbl(T, Pa)
        f = f12_mul(f12_sqr(f), _line_to_f12(line))
        if bit == "1":
            T, line = _line_add(T, Qa, Pa)
            f = f12_mul(f, _line_to_f12(line))
    return f12_conj(f)


def _exp_by_x_abs(a):
    return f12_pow(a, X_ABS)


def final_exponentiation(f):
    """f^(3*(p^12-1)/r).  Easy part (p^6-1)(p^2+1), then the HHT hard part
    3*Phi12(p)/r = (x-1)^2 (x+p) (x^2+p^2-1) + 3  with x = -X_ABS."""
    t = f12_mul(f12_conj(f), f12_inv(f))            # f^(p^6-1)
    m = f12_mul(f12_frob_n(t, 2), t)                # ^(p^2+1): now in the cyclotomic subgroup
    inv = f12_conj                                  # inverse == conjugate there

    def exp_x(a):                                   # a^x, x negative
        return inv(_exp_by_x_abs(a))

    a = f12_mul(exp_x(m), inv(m))                   # m^(x-1)
    a = f12_mul(exp_x(a), inv(a))                   # m^((x-1)^2)
    b = f12_mul(exp_x(a), f12_frob(a))              # ^(x+p)
    c = f12_mul(f12_mul(exp_x(exp_x(b)), f12_frob_n(b, 2)), inv(b))   # ^(x^2+p^2-1)
    return f12_mul(c, f12_mul(f12_sqr(m), m))


def pairing(Pj, Qj):
    """e(P, Q)^3 for Jacobian P in E1, Q in E2."""
    return final_exponentiation(miller_loop(E1.to_affine(Pj), E2.to_affine(Qj)))


def pairing_product_is_one(pairs):
    """prod e(P_i, Q_i) == 1 for a list of (P_jacobian, Q_jacobian)."""
    f = F12_ONE
    for Pj, Qj in pairs:
        f = f12_mul(f, miller_loop(E1.to_affine(Pj), E2.to_affine(Qj)))
    return final_exponentiation(f) == F12_ONE
