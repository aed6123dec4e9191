// Small complex DFT codelets held in registers (natural order in and out), shared by every ring-FFT kernel.
//
// Replaces the FFTW codelets behind ssht's phi transforms (reference call sites:
// /root/reference/pxmcmc/transforms.py:95-98, pxmcmc/measurements.py:223-239).
//
// DADD, DMUL and DFMA cost the same FP64 issue slot, so the codelets are written in multiply-add form (Linzer &
// Feig): a butterfly a +- w b with w = c (1 -+ i t), t = s / c, is
//     u = b (1 -+ i t)        2 FMA        (|t| <= 1: the factor with the larger modulus is pulled out)
//     a +- c u                4 FMA
// = 6 instructions instead of 4 (complex product) + 4 (additions); a radix-4 butterfly whose inputs carry the
// geometric twiddles (1, a, a^2, a^3) -- every twiddled radix-4 stage of a Cooley-Tukey split has this form --
// costs 24 instead of 28, and 20 when a^2 = -+i: 7.8 % fewer FP64 instructions in the persistent ring FFT.
// Measured (B200, gpurun_out/r2l_lf.log): neutral on its own -- the kernel is bound by load latency at 8 warps per
// SM, not by FP64 issue (ablations in DESIGN.md) -- and 5 % faster than the product-then-add form once the
// inter-pass twiddles are computed (twiddle_tree) rather than loaded.  PXM_HD: the same code runs on the host in
// tests/test_host_cpu.py (compiled by nvcc for the host) against a direct O(N^2) DFT.
#pragma once
#include <cuda_runtime.h>

#ifndef PXM_HD
#define PXM_HD __host__ __device__ __forceinline__
#endif

typedef double2 cplx;

PXM_HD cplx cmul(cplx a, cplx b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
PXM_HD cplx cmulc(cplx a, cplx b) {  // a * conj(b)
  return make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
PXM_HD cplx cadd(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
PXM_HD cplx csub(cplx a, cplx b) { return make_double2(a.x - b.x, a.y - b.y); }
// multiply by -i (forward) or +i (inverse)
template <bool INV>
PXM_HD cplx rot90(cplx a) {
  return INV ? make_double2(-a.y, a.x) : make_double2(a.y, -a.x);
}
// multiply by exp(-+ i pi/4)
template <bool INV>
PXM_HD cplx rot45(cplx a) {
  const double h = 0.70710678118654752440;
  return INV ? make_double2(h * (a.x - a.y), h * (a.x + a.y)) : make_double2(h * (a.x + a.y), h * (a.y - a.x));
}
template <bool INV>
PXM_HD cplx twc(cplx a, double c, double s) {  // a * (c -+ i s)
  return INV ? make_double2(a.x * c - a.y * s, a.x * s + a.y * c) : make_double2(a.x * c + a.y * s, a.y * c - a.x * s);
}

// (cos, sin)(2 pi k / 32), k < 16; k is a compile-time constant at every use (unrolled loops)
PXM_HD void w32(int k, double* c, double* sn) {
  constexpr double CW32[16] = {1.0, 0.98078528040323044913, 0.92387953251128675613, 0.83146961230254523708,
                               0.70710678118654752440, 0.55557023301960222474, 0.38268343236508977173,
                               0.19509032201612826785, 0.0, -0.19509032201612826785, -0.38268343236508977173,
                               -0.55557023301960222474, -0.70710678118654752440, -0.83146961230254523708,
                               -0.92387953251128675613, -0.98078528040323044913};
  constexpr double SW32[16] = {0.0, 0.19509032201612826785, 0.38268343236508977173, 0.55557023301960222474,
                               0.70710678118654752440, 0.83146961230254523708, 0.92387953251128675613,
                               0.98078528040323044913, 1.0, 0.98078528040323044913, 0.92387953251128675613,
                               0.83146961230254523708, 0.70710678118654752440, 0.55557023301960222474,
                               0.38268343236508977173, 0.19509032201612826785};
  *c = CW32[k];
  *sn = SW32[k];
}
template <bool INV>
PXM_HD cplx tw32(cplx a, int k) {
  double c, sn;
  w32(k, &c, &sn);
  return twc<INV>(a, c, sn);
}

// ---- multiply-add butterflies -------------------------------------------------------------------------------
// u = b * (c -+ i s) / f with f = the larger of |c|, |s| (returned in *f): 2 FMA (c, s compile-time constants)
template <bool INV>
PXM_HD cplx tw_unit(cplx b, double c, double s, double* f) {
  const bool cbig = (c < 0 ? -c : c) >= (s < 0 ? -s : s);
  if (cbig) {
    const double t = s / c;
    *f = c;
    return INV ? make_double2(b.x - t * b.y, b.y + t * b.x) : make_double2(b.x + t * b.y, b.y - t * b.x);
  }
  const double r = c / s;
  *f = s;
  // b (c -+ i s) = s (b r -+ i b):  forward (r b.x + b.y, r b.y - b.x), inverse (r b.x - b.y, r b.y + b.x)
  return INV ? make_double2(r * b.x - b.y, r * b.y + b.x) : make_double2(r * b.x + b.y, r * b.y - b.x);
}
// (a, b) <- (a + w b, a - w b), w = c -+ i s
template <bool INV>
PXM_HD void bfly_tw(cplx& a, cplx& b, double c, double s) {
  if (s == 0.0 && c == 1.0) {
    const cplx t = a;
    a = cadd(t, b);
    b = csub(t, b);
    return;
  }
  if (c == 0.0 && s == 1.0) {  // w = -+i
    const cplx r = rot90<INV>(b), t = a;
    a = cadd(t, r);
    b = csub(t, r);
    return;
  }
#ifdef PXM_DFT_PLAIN_TWIDDLES  // development switch: complex product, then the additions
  const cplx p = twc<INV>(b, c, s), t0 = a;
  a = cadd(t0, p);
  b = csub(t0, p);
#else
  double f;
  const cplx u = tw_unit<INV>(b, c, s, &f);
  const cplx t = a;
  a = make_double2(t.x + f * u.x, t.y + f * u.y);
  b = make_double2(t.x - f * u.x, t.y - f * u.y);
#endif
}
// a + w b
template <bool INV>
PXM_HD cplx add_tw(cplx a, cplx b, double c, double s) {
  if (s == 0.0 && c == 1.0) return cadd(a, b);
  if (c == 0.0 && s == 1.0) return cadd(a, rot90<INV>(b));
#ifdef PXM_DFT_PLAIN_TWIDDLES
  return cadd(a, twc<INV>(b, c, s));
#else
  double f;
  const cplx u = tw_unit<INV>(b, c, s, &f);
  return make_double2(a.x + f * u.x, a.y + f * u.y);
#endif
}

// ---- small DFTs in registers: y_q = sum_r x_r exp(-+ 2 pi i r q / R), natural order in and out
template <bool INV>
PXM_HD void dft2(cplx& a, cplx& b) {
  const cplx t = a;
  a = cadd(t, b);
  b = csub(t, b);
}
template <bool INV>
PXM_HD void dft4(cplx& x0, cplx& x1, cplx& x2, cplx& x3) {
  const cplx a02 = cadd(x0, x2), s02 = csub(x0, x2), a13 = cadd(x1, x3), s13 = rot90<INV>(csub(x1, x3));
  x0 = cadd(a02, a13);
  x2 = csub(a02, a13);
  x1 = cadd(s02, s13);
  x3 = csub(s02, s13);
}
// radix-4 butterfly of the inputs (y0, a y1, a^2 y2, a^3 y3), a = ac -+ i as, a^2 = a2c -+ i a2s
template <bool INV>
PXM_HD void dft4_tw(cplx& y0, cplx& y1, cplx& y2, cplx& y3, double ac, double as, double a2c, double a2s) {
  bfly_tw<INV>(y0, y2, a2c, a2s);  // y0 = z0 + z2, y2 = z0 - z2
  bfly_tw<INV>(y1, y3, a2c, a2s);  // y1 = (z1 + z3) / a, y3 = (z1 - z3) / a
  cplx r = rot90<INV>(y3);
  bfly_tw<INV>(y0, y1, ac, as);    // X0, X2
  bfly_tw<INV>(y2, r, ac, as);     // X1, X3
  const cplx x2 = y1;
  y1 = y2;
  y2 = x2;
  y3 = r;
}
template <bool INV>
PXM_HD void dft8(cplx* x) {
  // 8 = 2 x 4: r = 4 r1 + r2 (r1<2, r2<4), q = q1 + 2 q2
  dft2<INV>(x[0], x[4]);
  dft2<INV>(x[1], x[5]);
  dft2<INV>(x[2], x[6]);
  dft2<INV>(x[3], x[7]);
  const double h = 0.70710678118654752440;
  dft4<INV>(x[0], x[1], x[2], x[3]);                        // q1 = 0 -> outputs q = 2 q2
  dft4_tw<INV>(x[4], x[5], x[6], x[7], h, h, 0.0, 1.0);     // q1 = 1: twiddles w8^(r2) -> outputs q = 1 + 2 q2
  // reorder to natural q: currently x[q2] = X[2 q2], x[4+q2] = X[1+2 q2]
  const cplx t1 = x[1], t2 = x[2], t3 = x[3], t4 = x[4], t5 = x[5], t6 = x[6];
  x[1] = t4;
  x[2] = t1;
  x[3] = t5;
  x[4] = t2;
  x[5] = t6;
  x[6] = t3;
}
PXM_HD void transpose4x4(cplx* x) {
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = a + 1; b < 4; ++b) {
      const cplx t = x[4 * a + b];
      x[4 * a + b] = x[4 * b + a];
      x[4 * b + a] = t;
    }
}
template <bool INV>
PXM_HD void dft16(cplx* x) {
  // 16 = 4 x 4: r = 4 r1 + r2, q = q1 + 4 q2; the twiddles w16^(r2 q1) ride on the second radix-4 stage
  const double c1 = 0.92387953251128675613, s1 = 0.38268343236508977173;  // cos, sin(pi/8)
  const double h = 0.70710678118654752440;
#pragma unroll
  for (int r2 = 0; r2 < 4; ++r2) dft4<INV>(x[r2], x[4 + r2], x[8 + r2], x[12 + r2]);
  dft4<INV>(x[0], x[1], x[2], x[3]);
  dft4_tw<INV>(x[4], x[5], x[6], x[7], c1, s1, h, h);          // a = w16
  dft4_tw<INV>(x[8], x[9], x[10], x[11], h, h, 0.0, 1.0);      // a = w16^2, a^2 = -+i
  dft4_tw<INV>(x[12], x[13], x[14], x[15], s1, c1, -h, h);     // a = w16^3, a^2 = w16^6
  transpose4x4(x);  // x[4 q1 + q2] = X[q1 + 4 q2] -> natural order
}
// forward DFT16 of x_j w32^j (the odd half of a radix-32 split whose upper inputs are zero): the input twiddles
// w32^(4 r1 + r2) = w8^r1 w32^r2 ride on the two radix-4 stages
PXM_HD void dft16_pre32(cplx* x) {
  const double h = 0.70710678118654752440;
#pragma unroll
  for (int r2 = 0; r2 < 4; ++r2) dft4_tw<false>(x[r2], x[4 + r2], x[8 + r2], x[12 + r2], h, h, 0.0, 1.0);
#pragma unroll
  for (int q1 = 0; q1 < 4; ++q1) {
    double bc, bs, b2c, b2s;
    w32(1 + 2 * q1, &bc, &bs);
    w32(2 + 4 * q1, &b2c, &b2s);
    dft4_tw<false>(x[4 * q1], x[4 * q1 + 1], x[4 * q1 + 2], x[4 * q1 + 3], bc, bs, b2c, b2s);
  }
  transpose4x4(x);
}
template <int R, bool INV>
PXM_HD void dftR(cplx* x) {
  if (R == 2) dft2<INV>(x[0], x[1]);
  if (R == 4) dft4<INV>(x[0], x[1], x[2], x[3]);
  if (R == 8) dft8<INV>(x);
  if (R == 16) dft16<INV>(x);
}
template <bool INV>
PXM_HD void dft32(cplx* x) {
  cplx e[16], o[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    e[k] = x[2 * k];
    o[k] = x[2 * k + 1];
  }
  dft16<INV>(e);
  dft16<INV>(o);
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    double c, s;
    w32(k, &c, &s);
    bfly_tw<INV>(e[k], o[k], c, s);
    x[k] = e[k];
    x[k + 16] = o[k];
  }
}
template <int R, bool INV>
PXM_HD void dftN(cplx* x) {
  if (R == 32)
    dft32<INV>(x);
  else
    dftR<R, INV>(x);
}

// forward DFT of length R whose inputs x[R/2..R) are zero:  X[2q] = DFT_{R/2}(x)[q],
// X[2q+1] = DFT_{R/2}(x_j W_R^j)[q]
template <int R>
PXM_HD void dft_half_in(cplx* x) {
  constexpr int H = R / 2, STEP = 32 / R;
  cplx e[H], o[H];
#pragma unroll
  for (int j = 0; j < H; ++j) {
    e[j] = x[j];
    o[j] = (R == 32 || j == 0) ? x[j] : tw32<false>(x[j], j * STEP);
  }
  dftR<H, false>(e);
  if (R == 32)
    dft16_pre32(o);
  else
    dftR<H, false>(o);
#pragma unroll
  for (int q = 0; q < H; ++q) {
    x[2 * q] = e[q];
    x[2 * q + 1] = o[q];
  }
}
// inverse DFT of length R of which only the outputs z[0..R/2) are needed (left in x[0..R/2)):
// z[j] = E[j] + W_R^{-j} O[j],  E / O = inverse DFT_{R/2} of the even / odd inputs
template <int R>
PXM_HD void dft_half_out(cplx* x) {
  constexpr int H = R / 2, STEP = 32 / R;
  cplx e[H], o[H];
#pragma unroll
  for (int k = 0; k < H; ++k) {
    e[k] = x[2 * k];
    o[k] = x[2 * k + 1];
  }
  dftR<H, true>(e);
  dftR<H, true>(o);
#pragma unroll
  for (int j = 0; j < H; ++j) {
    double c, s;
    w32(j * STEP, &c, &s);
    x[j] = add_tw<true>(e[j], o[j], c, s);
  }
}
