// HEALPix (RING scheme) side of the harmonic transforms: ring geometry and the
// per-ring azimuthal DFT with the ring's phi offset.
//
// Replaces healpy.alm2map / healpy.map2alm (healpix_cxx + libsharp) that the
// reference reaches through /root/reference/pxmcmc/utils.py:106-113 and its
// drivers (experiments/earthtopography/main.py:80-82, experiments/weaklensing/
// main.py:31-37).  These run ONCE per experiment (data preparation), not per
// iteration: the Legendre part reuses the DMMA contraction kernel with a table
// evaluated at the HEALPix ring colatitudes; the azimuthal part below is a plain
// O(n_phi * L) DFT per ring (rings have 4, 8, ..., 4 nside pixels, each with its
// own phi_0; 3e9 flops at nside 256 -- a fraction of a millisecond of FP64).
//
// Geometry (HEALPix primer; restated in oracle/healpix_ref.py): rings i = 1 ..
// 4 nside - 1; north cap i < nside: n = 4 i pixels, z = 1 - i^2 / (3 nside^2),
// phi_j = (j + 1/2) pi / (2 i); belt nside <= i <= 3 nside: n = 4 nside,
// z = 4/3 - 2 i / (3 nside), phi_j = (j + s/2) pi / (2 nside), s = (i - nside + 1) mod 2;
// south cap: mirror image of the north cap.
#include "pxm_common.cuh"

namespace {

typedef double2 cplx;

struct HpxRing {
  int n;            // pixels on the ring
  int shift;        // 1: phi_j = (j + 1/2) 2 pi / n, 0: phi_j = j 2 pi / n
  long long start;  // index of the ring's first pixel
};

__host__ __device__ inline HpxRing hpx_ring(int nside, int r /* 0-based */) {
  const int i = r + 1;
  HpxRing g;
  const long long ncap = 2LL * nside * (nside - 1);
  const long long npix = 12LL * nside * nside;
  if (i < nside) {
    g.n = 4 * i;
    g.shift = 1;
    g.start = 2LL * i * (i - 1);
  } else if (i <= 3 * nside) {
    g.n = 4 * nside;
    g.shift = (i - nside + 1) & 1;
    g.start = ncap + (long long)(i - nside) * 4 * nside;
  } else {
    const int ip = 4 * nside - i;
    g.n = 4 * ip;
    g.shift = 1;
    g.start = npix - 2LL * ip * (ip + 1);
  }
  return g;
}

// exp(+i m phi_j) with phi_j = (2 j + shift) pi / n, reduced exactly in integers
__device__ __forceinline__ cplx phase(int m, int j, const HpxRing& g) {
  const long long two_n = 2LL * g.n;
  long long k = ((long long)m * (2LL * j + g.shift)) % two_n;
  if (k < 0) k += two_n;
  double s, c;
  sincospi((double)k / (double)g.n, &s, &c);
  return make_double2(c, s);
}

// F (ring coefficients, k4-interleaved, paired +-m columns, one chain) -> map pixels
//   f[r, j] = sum_{|m| < L} F_m[r] e^{+i m phi_j}
// The ring side of the paired layout stores (-1)^m F_{-m} in the -m columns (see pxm_fft.cu).
__global__ void k_hpx_synth(int nside, int L, cplx* __restrict__ map, const double* __restrict__ F,
                            unsigned long long f_off, unsigned long long slot_stride, int nld) {
  const int r = blockIdx.x;
  const HpxRing g = hpx_ring(nside, r);
  const size_t frow = f_off + ((size_t)(r >> 2) * (size_t)nld) * 4 + (size_t)(r & 3);
  for (int j = threadIdx.x; j < g.n; j += blockDim.x) {
    double re = 0.0, im = 0.0;
    for (int am = 0; am < L; ++am) {
      const size_t base = frow + (size_t)am * slot_stride;
      const cplx e = phase(am, j, g);
      const double pr = F[base], pi = F[base + 4];
      re += pr * e.x - pi * e.y;
      im += pr * e.y + pi * e.x;
      if (am > 0) {
        const double sg = (am & 1) ? -1.0 : 1.0;
        const double qr = sg * F[base + 8], qi = sg * F[base + 12];
        // e^{-i m phi} = conj(e)
        re += qr * e.x + qi * e.y;
        im += qi * e.x - qr * e.y;
      }
    }
    map[g.start + j] = make_double2(re, im);
  }
}

// map pixels -> F:  F_m[r] = sum_j f[r, j] e^{-i m phi_j}   (unweighted adjoint of the above)
__global__ void k_hpx_anal(int nside, int L, const cplx* __restrict__ map, double* __restrict__ F,
                           unsigned long long f_off, unsigned long long slot_stride, int nld) {
  const int r = blockIdx.x;
  const HpxRing g = hpx_ring(nside, r);
  const size_t frow = f_off + ((size_t)(r >> 2) * (size_t)nld) * 4 + (size_t)(r & 3);
  for (int am = threadIdx.x; am < L; am += blockDim.x) {
    double pr = 0.0, pi = 0.0, qr = 0.0, qi = 0.0;
    for (int j = 0; j < g.n; ++j) {
      const cplx v = map[g.start + j];
      const cplx e = phase(am, j, g);
      // +m: v * conj(e);  -m: v * e
      pr += v.x * e.x + v.y * e.y;
      pi += v.y * e.x - v.x * e.y;
      qr += v.x * e.x - v.y * e.y;
      qi += v.y * e.x + v.x * e.y;
    }
    const size_t base = frow + (size_t)am * slot_stride;
    F[base] = pr;
    F[base + 4] = pi;
    const double sg = (am & 1) ? -1.0 : 1.0;
    F[base + 8] = am > 0 ? sg * qr : 0.0;
    F[base + 12] = am > 0 ? sg * qi : 0.0;
  }
}

}  // namespace

// host: (sin, cos) of theta/2 for every ring, from z = cos(theta) without cancellation
void pxm_hpx_half_angles(int nside, double* out /* [4 nside - 1][2] */) {
  const int nr = 4 * nside - 1;
  for (int r = 0; r < nr; ++r) {
    const int i = r + 1;
    double sh, ch;
    if (i < nside) {  // 1 - z = i^2 / (3 nside^2)
      sh = (double)i / ((double)nside * sqrt(6.0));
      ch = sqrt(1.0 - sh * sh);
    } else if (i <= 3 * nside) {
      const double z = (4.0 * nside - 2.0 * i) / (3.0 * nside);
      sh = sqrt(0.5 * (1.0 - z));
      ch = sqrt(0.5 * (1.0 + z));
    } else {  // 1 + z = i'^2 / (3 nside^2)
      const int ip = 4 * nside - i;
      ch = (double)ip / ((double)nside * sqrt(6.0));
      sh = sqrt(1.0 - ch * ch);
    }
    out[2 * r] = sh;
    out[2 * r + 1] = ch;
  }
}

// dir 0: map -> F (analysis side), dir 1: F -> map
int pxm_hpx_ring_dft_launch(int dir, int nside, int L, const void* map, double* F, unsigned long long f_off,
                            unsigned long long slot_stride, int nld, cudaStream_t stream) {
  const int nr = 4 * nside - 1;
  if (dir == 0)
    k_hpx_anal<<<nr, 128, 0, stream>>>(nside, L, (const cplx*)map, F, f_off, slot_stride, nld);
  else
    k_hpx_synth<<<nr, 256, 0, stream>>>(nside, L, (cplx*)const_cast<void*>(map), F, f_off, slot_stride, nld);
  PXM_LAUNCHED();
  return PXM_OK;
}
