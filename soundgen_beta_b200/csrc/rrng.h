// R's random number stream on the host: what `set.seed(s)` followed by runif / rnorm / rgamma /
// rbinom / sample gives in the reference's R (3.4.0 pinned, packrat/packrat.lock:3): Mersenne-Twister,
// "Inversion" normals, "Rounding" sample().  north_star: every stochastic component is drawn on the
// host from R's stream exactly as the reference draws it and handed to the device as buffers.
// The R sources (src/main/RNG.c, src/nmath/{snorm,qnorm,sexp,rgamma,rbinom}.c, src/main/{random,sort}.c)
// are not under the reference tree: the algorithms below are the published ones (Matsumoto & Nishimura
// 1998; Wichura 1988 AS 241; Ahrens & Dieter 1972, 1974, 1982) with R's constants.
// Pinned in tests/test_frontend_host.py by the widely published answers set.seed(1): runif -> 0.2655087
// 0.3721239 0.5728534; rnorm -> -0.6264538 0.1836433 -0.8356286; set.seed(42): rnorm -> 1.37095845;
// set.seed(1): rexp -> 0.7551818; set.seed(1): sample(1:10) -> 3 4 5 7 2 8 9 6 10 1.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

struct RRng {
  uint32_t mt[624];
  int mti = 625;
  int64_t n_unif = 0;
  bool rejection_sampling = false;   // R >= 3.6 sample.kind = "Rejection"

  // set.seed(seed): Randomize -> RNG_Init -> FixupSeeds (RNG.c)
  void set_seed(uint32_t seed) {
    for (int j = 0; j < 50; j++) seed = 69069u * seed + 1u;
    seed = 69069u * seed + 1u;          // dummy[0], overwritten by mti = 624
    for (int j = 0; j < 624; j++) { seed = 69069u * seed + 1u; mt[j] = seed; }
    mti = 624;
    n_unif = 0;
  }
  // .Random.seed[2:626]
  void get_state(int32_t *out) const { out[0] = mti; for (int i = 0; i < 624; i++) out[i + 1] = (int32_t)mt[i]; }
  void set_state(const int32_t *in) { mti = in[0]; for (int i = 0; i < 624; i++) mt[i] = (uint32_t)in[i + 1]; }

  uint32_t genrand() {
    static const uint32_t mag01[2] = {0x0u, 0x9908b0dfu};
    if (mti >= 624) {
      int kk;
      uint32_t y;
      for (kk = 0; kk < 624 - 397; kk++) {
        y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
        mt[kk] = mt[kk + 397] ^ (y >> 1) ^ mag01[y & 1u];
      }
      for (; kk < 623; kk++) {
        y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
        mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ mag01[y & 1u];
      }
      y = (mt[623] & 0x80000000u) | (mt[0] & 0x7fffffffu);
      mt[623] = mt[396] ^ (y >> 1) ^ mag01[y & 1u];
      mti = 0;
    }
    uint32_t y = mt[mti++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
  }
  double unif_rand() {
    const double i2_32m1 = 2.328306437080797e-10;
    double v = (double)genrand() * 2.3283064365386963e-10;
    n_unif++;
    if (v <= 0.0) return 0.5 * i2_32m1;
    if ((1.0 - v) <= 0.0) return 1.0 - 0.5 * i2_32m1;
    return v;
  }
  static double qnorm_std(double p) {   // qnorm5(p, 0, 1, TRUE, FALSE): AS 241 PPND16
    if (p <= 0.0) return -INFINITY;
    if (p >= 1.0) return INFINITY;
    double q = p - 0.5, r, val;
    if (std::fabs(q) <= 0.425) {
      r = .180625 - q * q;
      return q * (((((((r * 2509.0809287301226727 + 33430.575583588128105) * r + 67265.770927008700853) * r +
                      45921.953931549871457) * r + 13731.693765509461125) * r + 1971.5909503065514427) * r +
                   133.14166789178437745) * r + 3.387132872796366608) /
             (((((((r * 5226.495278852545925 + 28729.085735721942674) * r + 39307.89580009271061) * r +
                  21213.794301586595867) * r + 5394.1960214247511077) * r + 687.1870074920579083) * r +
               42.313330701600911252) * r + 1.);
    }
    r = (q < 0) ? p : 1.0 - p;
    r = std::sqrt(-std::log(r));
    if (r <= 5.) {
      r += -1.6;
      val = (((((((r * 7.7454501427834140764e-4 + .0227238449892691845833) * r + .24178072517745061177) * r +
                 1.27045825245236838258) * r + 3.64784832476320460504) * r + 5.7694972214606914055) * r +
              4.6303378461565452959) * r + 1.42343711074968357734) /
            (((((((r * 1.05075007164441684324e-9 + 5.475938084995344946e-4) * r + .0151986665636164571966) * r +
                 .14810397642748007459) * r + .68976733498510000455) * r + 1.6763848301838038494) * r +
              2.05319162663775882187) * r + 1.);
    } else {
      r += -5.;
      val = (((((((r * 2.01033439929228813265e-7 + 2.71155556874348757815e-5) * r + .0012426609473880784386) * r +
                 .026532189526576123093) * r + .29656057182850489123) * r + 1.7848265399172913358) * r +
              5.4637849111641143699) * r + 1.3493881297270480396) /
            (((((((r * 2.04426310338993978564e-15 + 1.4215117583164458887e-7) * r + 1.8463183175100546818e-5) * r +
                 7.868691311456132591e-4) * r + .0148753612908506148525) * r + .13692988092273580531) * r +
              .59983224390749539497) * r + 1.);
    }
    return (q < 0.0) ? -val : val;
  }
  double norm_rand() {   // INVERSION (snorm.c)
    const double big = 134217728.0;
    double u = unif_rand();
    u = (double)(int)(big * u) + unif_rand();
    return qnorm_std(u / big);
  }
  double exp_rand() {    // sexp.c (q[3] is R's literal, which differs from the series value in the 5th digit)
    static const double q[] = {0.6931471805599453, 0.9333736875190459, 0.9888777961838675, 0.9984589039328340,
                               0.9998292811061389, 0.9999833164100727, 0.9999985691438767, 0.9999998906925558,
                               0.9999999924734159, 0.9999999995283275, 0.9999999999728814, 0.9999999999985598,
                               0.9999999999999289, 0.9999999999999968, 0.9999999999999999, 1.0000000000000000};
    double a = 0.;
    double u = unif_rand();
    while (u <= 0. || u >= 1.) u = unif_rand();
    for (;;) {
      u += u;
      if (u > 1.) break;
      a += q[0];
    }
    u -= 1.;
    if (u <= q[0]) return a + u;
    int i = 0;
    double ustar = unif_rand(), umin = ustar;
    do {
      ustar = unif_rand();
      if (umin > ustar) umin = ustar;
      i++;
    } while (u > q[i]);
    return a + umin * q[0];
  }
  double rnorm(double mean, double sd) { return mean + sd * norm_rand(); }
  // rbinom(1, size, pp): inversion branch (n * min(p, 1 - p) < 30; the path only calls rbinom(1, 1, p))
  double rbinom(double size, double pp) {
    int n = (int)std::floor(size + 0.5);
    if (n == 0 || pp == 0.) return 0;
    if (pp == 1.) return n;
    double p = std::fmin(pp, 1. - pp), q = 1. - p, r = p / q, g = r * (n + 1);
    double qn = std::pow(q, (double)n);
    int ix;
    for (;;) {
      ix = 0;
      double f = qn, u = unif_rand();
      bool done = false;
      for (;;) {
        if (u < f) { done = true; break; }
        if (ix > 110) break;
        u -= f;
        ix++;
        f *= (g / ix - r);
      }
      if (done) break;
    }
    if (pp > 0.5) ix = n - ix;
    return (double)ix;
  }
  // rgamma(1, shape = a, scale): GD for a >= 1, GS for a < 1 (rgamma.c)
  double rgamma(double a, double scale) {
    const double sqrt32 = 5.656854, exp_m1 = 0.36787944117144233;
    const double q1 = 0.04166669, q2 = 0.02083148, q3 = 0.00801191, q4 = 0.00144121, q5 = -7.388e-5,
                 q6 = 2.4511e-4, q7 = 2.424e-4;
    const double a1 = 0.3333333, a2 = -0.250003, a3 = 0.2000062, a4 = -0.1662921, a5 = 0.1423657,
                 a6 = -0.1367177, a7 = 0.1233795;
    if (std::isnan(a) || std::isnan(scale)) return NAN;
    if (a <= 0.0 || scale <= 0.0) { if (scale == 0. || a == 0.) return 0.; return NAN; }
    if (!std::isfinite(a) || !std::isfinite(scale)) return INFINITY;
    double e, p, q, r, t, u, v, w, x, ret_val;
    if (a < 1) {
      e = 1.0 + exp_m1 * a;
      for (;;) {
        p = e * unif_rand();
        if (p >= 1.0) {
          x = -std::log((e - p) / a);
          if (exp_rand() >= (1.0 - a) * std::log(x)) break;
        } else {
          x = std::exp(std::log(p) / a);
          if (exp_rand() >= x) break;
        }
      }
      return scale * x;
    }
    const double s2 = a - 0.5, s = std::sqrt(s2), d = sqrt32 - s * 12;
    t = norm_rand();
    x = s + 0.5 * t;
    ret_val = x * x;
    if (t >= 0) return scale * ret_val;
    u = unif_rand();
    if (d * u <= t * t * t) return scale * ret_val;
    r = 1 / a;
    const double q0 = ((((((q7 * r + q6) * r + q5) * r + q4) * r + q3) * r + q2) * r + q1) * r;
    double b, si, c;
    if (a <= 3.686) { b = 0.463 + s + 0.178 * s2; si = 1.235; c = 0.195 / s - 0.079 + 0.16 * s; }
    else if (a <= 13.022) { b = 1.654 + 0.0076 * s2; si = 1.68 / s + 0.275; c = 0.062 / s + 0.024; }
    else { b = 1.77; si = 0.75; c = 0.1515 / s; }
    auto quot = [&](double tt) {
      double vv = tt / (s + s);
      if (std::fabs(vv) <= 0.25)
        return q0 + 0.5 * tt * tt * ((((((a7 * vv + a6) * vv + a5) * vv + a4) * vv + a3) * vv + a2) * vv + a1) * vv;
      return q0 - s * tt + 0.25 * tt * tt + (s2 + s2) * std::log(1.0 + vv);
    };
    if (x > 0.0) {
      q = quot(t);
      if (std::log(1.0 - u) <= q) return scale * ret_val;
    }
    for (;;) {
      e = exp_rand();
      u = unif_rand();
      u = u + u - 1.0;
      t = (u < 0.0) ? b - si * e : b + si * e;
      if (t >= -0.71874483771719) {
        q = quot(t);
        if (q > 0.0) {
          w = std::expm1(q);
          if (c * std::fabs(u) <= w * std::exp(e - 0.5 * t * t)) break;
        }
      }
    }
    (void)v;
    x = s + 0.5 * t;
    return scale * x * x;
  }
  // R_unif_index
  double unif_index(double dn) {
    if (!rejection_sampling) return std::floor(dn * unif_rand());
    if (dn <= 0) return 0.0;
    int bits = (int)std::ceil(std::log2(dn));
    double dv;
    do {
      int64_t vv = 0;
      for (int n = 0; n <= bits; n += 16) {
        int v1 = (int)std::floor(unif_rand() * 65536);
        vv = 65536 * vv + v1;
      }
      if (bits < 64) vv &= (((int64_t)1) << bits) - 1;
      dv = (double)vv;
    } while (dn <= dv);
    return dv;
  }
  int sample_int1(int n) { return (int)unif_index((double)n) + 1; }   // sample.int(n, 1)
  // sort.c revsort: heapsort into descending order, ib alongside
  static void revsort(double *a, int *ib, int n) {
    int l, j, ir, i, ii;
    double ra;
    if (n <= 1) return;
    a--; ib--;
    l = (n >> 1) + 1;
    ir = n;
    for (;;) {
      if (l > 1) { l = l - 1; ra = a[l]; ii = ib[l]; }
      else {
        ra = a[ir]; ii = ib[ir];
        a[ir] = a[1]; ib[ir] = ib[1];
        if (--ir == 1) { a[1] = ra; ib[1] = ii; return; }
      }
      i = l;
      j = l << 1;
      while (j <= ir) {
        if (j < ir && a[j] > a[j + 1]) ++j;
        if (ra > a[j]) { a[i] = a[j]; ib[i] = ib[j]; j += (i = j); }
        else j = ir + 1;
      }
      a[i] = ra; ib[i] = ii;
    }
  }
  // sample.int(n, 1, prob = p): ProbSampleReplace (n < 200 categories); 1-based result
  int sample_prob1(const double *prob, int n) {
    double p[16];
    int perm[16];
    double tot = 0;
    for (int i = 0; i < n; i++) tot += prob[i];
    for (int i = 0; i < n; i++) { p[i] = prob[i] / tot; perm[i] = i + 1; }
    revsort(p, perm, n);
    for (int i = 1; i < n; i++) p[i] += p[i - 1];
    double rU = unif_rand();
    int j;
    for (j = 0; j < n - 1; j++) if (rU <= p[j]) break;
    return perm[j];
  }
};
