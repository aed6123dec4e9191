// Wigner-d / spin-weighted Legendre rows by three-term recurrence in l.
//
// Replaces the Wigner recursion inside ssht (called by the reference through
// pyssht/pys2let: /root/reference/pxmcmc/measurements.py:223-239,
// pxmcmc/transforms.py:95-98).  Design is NOT ssht's (which expands d^l(theta)
// through d^l(pi/2) and FFTs): here every (m, ring) pair runs the l-recurrence
// of SURVEY.md A.7 directly at the ring's colatitude and the values are written
// once, at plan creation, into swizzled 32x16 tiles that the DMMA contraction
// kernel streams.
//
//   Lambda^{m,s}[t, l] = (-1)^s sqrt((2l+1)/4pi) d^l_{m,-s}(theta_t)
//
// Seeds are formed in log space and the pair (d^{l-1}, d^l) carries a binary
// exponent so that seeds far below DBL_MIN (polar rings, large |m|) still grow
// into the correct values; the coefficient l(l+1)cos(b) - mn is evaluated in the
// cancellation-free half-angle form.
#pragma once
#include <math.h>

#include "pxm_common.cuh"

struct PxmWigner {
  double dp, dc;  // scaled d^{l-1}, d^{l}
  double s2, c2;  // sin^2(theta/2), cos^2(theta/2)
  int ex;         // true value = d * 2^ex  (ex <= 0)
  int l, m, n;
};

PXM_HD double pxm_lgamma(double x) { return lgamma(x); }

// sh = sin(theta/2), ch = cos(theta/2)
PXM_HD void pxm_wigner_init(PxmWigner& w, int m, int n, double sh, double ch) {
  const int am = m < 0 ? -m : m, an = n < 0 ? -n : n;
  const int l0 = am > an ? am : an;
  int a, pc, ps;
  double sgn = 1.0;
  if (am >= an) {
    a = n;
    if (m >= 0) {  // d^{l0}_{l0,n}
      pc = l0 + n;
      ps = l0 - n;
      sgn = ((l0 - n) & 1) ? -1.0 : 1.0;
    } else {  // d^{l0}_{-l0,n}
      pc = l0 - n;
      ps = l0 + n;
    }
  } else {
    a = m;
    if (n > 0) {  // d^{l0}_{m,l0}
      pc = l0 + m;
      ps = l0 - m;
    } else {  // d^{l0}_{m,-l0}
      pc = l0 - m;
      ps = l0 + m;
      sgn = ((l0 + m) & 1) ? -1.0 : 1.0;
    }
  }
  w.l = l0;
  w.m = m;
  w.n = n;
  w.s2 = sh * sh;
  w.c2 = ch * ch;
  w.dp = 0.0;
  w.ex = 0;
  if ((pc > 0 && ch == 0.0) || (ps > 0 && sh == 0.0)) {
    w.dc = 0.0;
    return;
  }
  double ln = 0.5 * (pxm_lgamma(2.0 * l0 + 1.0) - pxm_lgamma((double)(l0 + a) + 1.0) -
                     pxm_lgamma((double)(l0 - a) + 1.0));
  if (pc > 0) ln += pc * log(ch);
  if (ps > 0) ln += ps * log(sh);
  if (ln > -600.0) {
    w.dc = sgn * exp(ln);
  } else {
    const double ln2 = 0.6931471805599453;
    int e = (int)(ln / ln2);  // negative
    w.ex = e;
    w.dc = sgn * exp(ln - (double)e * ln2);
  }
}

PXM_HD double pxm_wigner_value(const PxmWigner& w) {
  if (w.ex == 0) return w.dc;
  if (w.ex < -1200) return 0.0;
  return ldexp(w.dc, w.ex);
}

// advance l -> l+1
PXM_HD void pxm_wigner_step(PxmWigner& w) {
  const int l = w.l;
  double dn;
  if (l == 0) {
    dn = (1.0 - 2.0 * w.s2) * w.dc;  // d^1_00 = cos(theta)
  } else {
    const double dl = (double)l, dl1 = (double)(l + 1);
    const double dm = (double)w.m, dnn = (double)w.n;
    const double al = sqrt((dl - dm) * (dl + dm) * ((dl - dnn) * (dl + dnn)));
    const double al1 = sqrt((dl1 - dm) * (dl1 + dm) * ((dl1 - dnn) * (dl1 + dnn)));
    const double ll1 = dl * dl1, mn = dm * dnn;
    const double coef = (w.s2 <= 0.5) ? ((ll1 - mn) - 2.0 * ll1 * w.s2) : (-(ll1 + mn) + 2.0 * ll1 * w.c2);
    dn = ((2.0 * dl + 1.0) * coef * w.dc - dl1 * al * w.dp) / (dl * al1);
  }
  w.dp = w.dc;
  w.dc = dn;
  w.l = l + 1;
  if (fabs(dn) > 0x1p300) {
    w.dp *= 0x1p-300;
    w.dc *= 0x1p-300;
    w.ex += 300;
  }
}

// half-angle sine/cosine of MW ring t at bandlimit Lg: theta/2 = pi (2t+1) / (2 (2Lg-1))
PXM_HD void pxm_mw_half_angle(int t, int Lg, double* sh, double* ch) {
  const int num = 2 * t + 1, den = 2 * (2 * Lg - 1);
  if (2 * num == den) {  // theta = pi exactly (south pole ring)
    *sh = 1.0;
    *ch = 0.0;
    return;
  }
#ifdef __CUDA_ARCH__
  sincospi((double)num / (double)den, sh, ch);
#else
  const double x = 3.14159265358979323846 * (double)num / (double)den;
  *sh = sin(x);
  *ch = cos(x);
#endif
}
