"""R's random number stream, restated (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

The reference draws everything from base R's default generators (R 3.4.0 pinned,
packrat/packrat.lock:3): Mersenne-Twister + "Inversion" normals + "Rounding"
`sample()` (R < 3.6).  The sources of those are NOT under /root/reference; this file
restates the published algorithms (R's src/main/RNG.c, src/nmath/{snorm,qnorm,sexp,
rgamma,rbinom}.c, src/main/{random,sort}.c as documented in ?RNGkind / ?set.seed):

* `set.seed(s)`: 50 rounds of the LCG `s = 69069 s + 1` (mod 2^32), then 625 more to
  fill dummy[0..624]; dummy[0] (= mti) is forced to 624 so the first draw regenerates.
* `unif_rand()`: MT19937 `genrand` * 2.3283064365386963e-10, clamped into (0, 1).
* `norm_rand()`: u = unif_rand(); u = (int)(2^27 u) + unif_rand(); qnorm(u / 2^27)
  with Wichura's AS 241 (PPND16) -- two uniforms per normal.
* `exp_rand()`: Ahrens & Dieter (1972) algorithm SA.
* `rgamma()`: Ahrens & Dieter GD (1982) for shape >= 1, GS (1974) for shape < 1.
* `rbinom()`: inversion for n p < 30 (the only regime the path uses: rbinom(1, 1, p)).
* `sample()`: `floor(n * unif_rand()) + 1` ("Rounding"); with `prob`: Walker is not
  used below 200 categories -- probabilities are sorted descending by `revsort` and one
  uniform is compared with their running sum.  `sample_kind='Rejection'` gives the
  R >= 3.6 bit-rejection sampler for pins made with a modern R.

Pins (tests/test_rrng.py): the widely published answers `set.seed(1); runif(3)` =
0.2655087 0.3721239 0.5728534, `set.seed(1); rnorm(3)` = -0.6264538 0.1836433
-0.8356286, `set.seed(42); rnorm(1)` = 1.37095845, `set.seed(123); runif(3)`,
`set.seed(1); rexp(1)` = 0.7551818, `set.seed(1); sample(1:10)` (R < 3.6) =
3 4 5 7 2 8 9 6 10 1.  rgamma has no published answer I can cite: it is pinned only
through its inputs (norm_rand / exp_rand / unif_rand) and by moment checks.
"""
from __future__ import annotations

import math

import numpy as np

_N, _M = 624, 397
_MASK = 0xFFFFFFFF
_I2_32M1 = 2.328306437080797e-10


class RRng:
    """Mersenne-Twister stream with R's seeding and R's derived generators."""

    def __init__(self, seed=None, sample_kind='Rounding'):
        self.mt = [0] * _N
        self.mti = _N + 1
        self.sample_kind = sample_kind
        self.n_unif = 0            # uniforms consumed so far (ledger diagnostics)
        if seed is not None:
            self.set_seed(seed)

    # ---- RNG.c: set.seed -> Randomize -> RNG_Init -> FixupSeeds
    def set_seed(self, seed):
        s = int(seed) & _MASK
        for _ in range(50):
            s = (69069 * s + 1) & _MASK
        dummy = [0] * (_N + 1)
        for j in range(_N + 1):
            s = (69069 * s + 1) & _MASK
            dummy[j] = s
        self.mt = dummy[1:]
        self.mti = _N          # dummy[0] = 624
        self.n_unif = 0

    def state(self):
        """.Random.seed[-1]: (mti, mt[0..623])"""
        return [self.mti] + list(self.mt)

    def set_state(self, st):
        self.mti = int(st[0])
        self.mt = [int(v) & _MASK for v in st[1:]]

    def _genrand(self):
        mt = self.mt
        if self.mti >= _N:
            for kk in range(_N - _M):
                y = (mt[kk] & 0x80000000) | (mt[kk + 1] & 0x7FFFFFFF)
                mt[kk] = mt[kk + _M] ^ (y >> 1) ^ (0x9908B0DF if (y & 1) else 0)
            for kk in range(_N - _M, _N - 1):
                y = (mt[kk] & 0x80000000) | (mt[kk + 1] & 0x7FFFFFFF)
                mt[kk] = mt[kk + (_M - _N)] ^ (y >> 1) ^ (0x9908B0DF if (y & 1) else 0)
            y = (mt[_N - 1] & 0x80000000) | (mt[0] & 0x7FFFFFFF)
            mt[_N - 1] = mt[_M - 1] ^ (y >> 1) ^ (0x9908B0DF if (y & 1) else 0)
            self.mti = 0
        y = mt[self.mti]
        self.mti += 1
        y ^= (y >> 11)
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= (y >> 18)
        return y & _MASK

    def unif_rand(self):
        v = self._genrand() * 2.3283064365386963e-10
        self.n_unif += 1
        if v <= 0.0:
            return 0.5 * _I2_32M1
        if 1.0 - v <= 0.0:
            return 1.0 - 0.5 * _I2_32M1
        return v

    def norm_rand(self):
        big = 134217728.0
        u = self.unif_rand()
        u = float(int(big * u)) + self.unif_rand()
        return qnorm_std(u / big)

    def exp_rand(self):
        q = _EXP_Q
        a = 0.0
        u = self.unif_rand()
        while u <= 0.0 or u >= 1.0:
            u = self.unif_rand()
        while True:
            u += u
            if u > 1.0:
                break
            a += q[0]
        u -= 1.0
        if u <= q[0]:
            return a + u
        i = 0
        ustar = self.unif_rand()
        umin = ustar
        while True:
            ustar = self.unif_rand()
            if umin > ustar:
                umin = ustar
            i += 1
            if not (u > q[i]):
                break
        return a + umin * q[0]

    # ---- vectorised R-level calls
    def runif(self, n):
        return np.array([self.unif_rand() for _ in range(int(n))], dtype=np.float64)

    def rnorm(self, n, mean=0.0, sd=1.0):
        n = int(n)
        mean = np.resize(np.asarray(mean, dtype=np.float64), n) if n else np.zeros(0)
        sd = np.resize(np.asarray(sd, dtype=np.float64), n) if n else np.zeros(0)
        out = np.empty(n)
        for i in range(n):
            out[i] = mean[i] + sd[i] * self.norm_rand()
        return out

    def rexp(self, n):
        return np.array([self.exp_rand() for _ in range(int(n))])

    def rbinom1(self, size, pp):
        """rbinom(1, size, pp), inversion branch (size * min(pp, 1 - pp) < 30)."""
        n = int(math.floor(size + 0.5))
        if n == 0 or pp == 0.0:
            return 0
        if pp == 1.0:
            return n
        p = min(pp, 1.0 - pp)
        q = 1.0 - p
        if n * p >= 30.0:
            raise NotImplementedError('rbinom BTPE branch is not on the path')
        r = p / q
        g = r * (n + 1)
        qn = q ** n
        while True:
            ix = 0
            f = qn
            u = self.unif_rand()
            done = False
            while True:
                if u < f:
                    done = True
                    break
                if ix > 110:
                    break
                u -= f
                ix += 1
                f *= (g / ix - r)
            if done:
                break
        if pp > 0.5:
            ix = n - ix
        return ix

    def rgamma1(self, a, scale):
        """rgamma(1, shape = a, scale = scale) (the R-level call passes rate: scale = 1 / rate)."""
        if math.isnan(a) or math.isnan(scale):
            return float('nan')
        if a <= 0.0 or scale <= 0.0:
            if scale == 0.0 or a == 0.0:
                return 0.0
            return float('nan')
        if math.isinf(a) or math.isinf(scale):
            return float('inf')
        if a < 1.0:     # GS
            e = 1.0 + 0.36787944117144233 * a
            while True:
                p = e * self.unif_rand()
                if p >= 1.0:
                    x = -math.log((e - p) / a)
                    if self.exp_rand() >= (1.0 - a) * math.log(x):
                        break
                else:
                    x = math.exp(math.log(p) / a)
                    if self.exp_rand() >= x:
                        break
            return scale * x
        # GD
        q1, q2, q3, q4, q5, q6, q7 = (0.04166669, 0.02083148, 0.00801191, 0.00144121, -7.388e-5,
                                      2.4511e-4, 2.424e-4)
        a1, a2, a3, a4, a5, a6, a7 = (0.3333333, -0.250003, 0.2000062, -0.1662921, 0.1423657,
                                      -0.1367177, 0.1233795)
        s2 = a - 0.5
        s = math.sqrt(s2)
        d = 5.656854 - s * 12.0
        t = self.norm_rand()
        x = s + 0.5 * t
        ret = x * x
        if t >= 0.0:
            return scale * ret
        u = self.unif_rand()
        if d * u <= t * t * t:
            return scale * ret
        r = 1.0 / a
        q0 = ((((((q7 * r + q6) * r + q5) * r + q4) * r + q3) * r + q2) * r + q1) * r
        if a <= 3.686:
            b = 0.463 + s + 0.178 * s2
            si = 1.235
            c = 0.195 / s - 0.079 + 0.16 * s
        elif a <= 13.022:
            b = 1.654 + 0.0076 * s2
            si = 1.68 / s + 0.275
            c = 0.062 / s + 0.024
        else:
            b = 1.77
            si = 0.75
            c = 0.1515 / s

        def quot(t):
            v = t / (s + s)
            if abs(v) <= 0.25:
                return q0 + 0.5 * t * t * ((((((a7 * v + a6) * v + a5) * v + a4) * v + a3) * v + a2) * v + a1) * v
            return q0 - s * t + 0.25 * t * t + (s2 + s2) * math.log(1.0 + v)
        if x > 0.0:
            if math.log(1.0 - u) <= quot(t):
                return scale * ret
        while True:
            e = self.exp_rand()
            u = self.unif_rand()
            u = u + u - 1.0
            t = b - si * e if u < 0.0 else b + si * e
            if t >= -0.71874483771719:
                q = quot(t)
                if q > 0.0:
                    w = math.expm1(q)
                    if c * abs(u) <= w * math.exp(e - 0.5 * t * t):
                        break
        x = s + 0.5 * t
        return scale * x * x

    def rgamma(self, n, shape, rate=1.0):
        return np.array([self.rgamma1(shape, 1.0 / rate) for _ in range(int(n))])

    # ---- sample()
    def unif_index(self, dn):
        if self.sample_kind == 'Rounding':
            return int(math.floor(dn * self.unif_rand()))
        if dn <= 0:
            return 0
        bits = int(math.ceil(math.log2(dn)))
        while True:
            v = 0
            n = 0
            while n <= bits:
                v1 = int(math.floor(self.unif_rand() * 65536))
                v = 65536 * v + v1
                n += 16
            if bits < 64:
                v &= (1 << bits) - 1
            if not (dn <= v):
                return v

    def sample_int1(self, n):
        """sample.int(n, 1): 1-based."""
        return self.unif_index(float(n)) + 1

    def sample_prob1(self, prob):
        """sample.int(length(prob), 1, prob = prob): 1-based index."""
        p = [float(v) for v in prob]
        tot = sum(p)
        p = [v / tot for v in p]
        perm = list(range(1, len(p) + 1))
        revsort(p, perm)
        for i in range(1, len(p)):
            p[i] += p[i - 1]
        ru = self.unif_rand()
        j = 0
        while j < len(p) - 1:
            if ru <= p[j]:
                break
            j += 1
        return perm[j]


def revsort(a, ib):
    """sort.c revsort(): heapsort into descending order, ib alongside (in place)."""
    n = len(a)
    if n <= 1:
        return
    A = [0.0] + a
    B = [0] + ib
    l = (n >> 1) + 1
    ir = n
    while True:
        if l > 1:
            l -= 1
            ra, ii = A[l], B[l]
        else:
            ra, ii = A[ir], B[ir]
            A[ir], B[ir] = A[1], B[1]
            ir -= 1
            if ir == 1:
                A[1], B[1] = ra, ii
                break
        i = l
        j = l << 1
        while j <= ir:
            if j < ir and A[j] > A[j + 1]:
                j += 1
            if ra > A[j]:
                A[i], B[i] = A[j], B[j]
                i = j
                j += i
            else:
                j = ir + 1
        A[i], B[i] = ra, ii
    a[:] = A[1:]
    ib[:] = B[1:]


# exp_rand table: q[k-1] = sum_{i=1..k} ln(2)^i / i!
_EXP_Q = [0.6931471805599453, 0.9333736875190459, 0.9888777961838675, 0.9984589039328340,
          0.9998292811061389, 0.9999833164100727, 0.9999985691438767, 0.9999998906925558,
          0.9999999924734159, 0.9999999995283275, 0.9999999999728814, 0.9999999999985598,
          0.9999999999999289, 0.9999999999999968, 0.9999999999999999, 1.0000000000000000]


def qnorm_std(p):
    """qnorm5(p, 0, 1, lower = TRUE, log = FALSE): Wichura (1988) AS 241, PPND16."""
    if p <= 0.0:
        return float('-inf')
    if p >= 1.0:
        return float('inf')
    q = p - 0.5
    if abs(q) <= 0.425:
        r = .180625 - q * q
        return q * (((((((r * 2509.0809287301226727 + 33430.575583588128105) * r + 67265.770927008700853) * r +
                        45921.953931549871457) * r + 13731.693765509461125) * r + 1971.5909503065514427) * r +
                     133.14166789178437745) * r + 3.387132872796366608) / \
            (((((((r * 5226.495278852545925 + 28729.085735721942674) * r + 39307.89580009271061) * r +
                 21213.794301586595867) * r + 5394.1960214247511077) * r + 687.1870074920579083) * r +
              42.313330701600911252) * r + 1.)
    r = p if q < 0 else 1.0 - p
    r = math.sqrt(-math.log(r))
    if r <= 5.:
        r += -1.6
        val = (((((((r * 7.7454501427834140764e-4 + .0227238449892691845833) * r + .24178072517745061177) * r +
                   1.27045825245236838258) * r + 3.64784832476320460504) * r + 5.7694972214606914055) * r +
                4.6303378461565452959) * r + 1.42343711074968357734) / \
            (((((((r * 1.05075007164441684324e-9 + 5.475938084995344946e-4) * r + .0151986665636164571966) * r +
                 .14810397642748007459) * r + .68976733498510000455) * r + 1.6763848301838038494) * r +
              2.05319162663775882187) * r + 1.)
    else:
        r += -5.
        val = (((((((r * 2.01033439929228813265e-7 + 2.71155556874348757815e-5) * r + .0012426609473880784386) * r +
                   .026532189526576123093) * r + .29656057182850489123) * r + 1.7848265399172913358) * r +
                5.4637849111641143699) * r + 1.3493881297270480396) / \
            (((((((r * 2.04426310338993978564e-15 + 1.4215117583164458887e-7) * r + 1.8463183175100546818e-5) * r +
                 7.868691311456132591e-4) * r + .0148753612908506148525) * r + .13692988092273580531) * r +
              .59983224390749539497) * r + 1.)
    if q < 0.0:
        val = -val
    return val
