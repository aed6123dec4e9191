// Batched phi-direction DFT of MW rings (odd length n = 2l-1) for every
// bandlimit of a transform in ONE launch, fused with the transposition between
// the pixel layout [chain][ring t][phi p] and the per-m "k4-interleaved" layout
// the Legendre contraction consumes.
//
// Replaces the FFTW calls inside ssht (reference call sites:
// /root/reference/pxmcmc/transforms.py:95-98, pxmcmc/measurements.py:223-239).
// Odd, often prime lengths (511, 389, 259, 173, ...) => Bluestein chirp-z with a
// power-of-two length M >= 2n-1 done entirely in shared memory: radix-4(/2)
// decimation-in-frequency forward, pointwise product with the precomputed filter
// spectrum (kept in the kernel's own digit-reversed order, so no reordering pass
// exists anywhere), decimation-in-time inverse.
#include "pxm_common.cuh"

namespace {

typedef double2 cplx;

__device__ __forceinline__ cplx cmul(cplx a, cplx b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ cplx cmulc(cplx a, cplx b) {  // a * conj(b)
  return make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cplx cmuli(cplx a) { return make_double2(-a.y, a.x); }    // * i
__device__ __forceinline__ cplx cmulni(cplx a) { return make_double2(a.y, -a.x); }   // * -i

// In-place forward FFT (e^{-i}) of `nr` rings of length M held at s[r*M + i];
// output in the digit-reversed order defined by this very pass sequence.
__device__ void fft_forward_dif(cplx* s, int nr, int M, int logM, const cplx* __restrict__ tw) {
  int Ls = M;
  if (logM & 1) {
    const int h = M >> 1;
    for (int idx = threadIdx.x; idx < nr * h; idx += blockDim.x) {
      const int r = idx / h, j = idx - r * h;
      cplx* p = s + r * M;
      const cplx x0 = p[j], x1 = p[j + h];
      p[j] = cadd(x0, x1);
      p[j + h] = cmul(csub(x0, x1), tw[j]);
    }
    Ls = h;
    __syncthreads();
  }
  for (; Ls >= 4; Ls >>= 2) {
    const int q4 = Ls >> 2, nb = M >> 2, tstep = M / Ls;
    for (int idx = threadIdx.x; idx < nr * nb; idx += blockDim.x) {
      const int r = idx / nb, b = idx - r * nb;
      const int blk = b / q4, j = b - blk * q4;
      cplx* p = s + r * M + blk * Ls + j;
      const cplx x0 = p[0], x1 = p[q4], x2 = p[2 * q4], x3 = p[3 * q4];
      const cplx a02 = cadd(x0, x2), s02 = csub(x0, x2), a13 = cadd(x1, x3), s13 = csub(x1, x3);
      const cplx y0 = cadd(a02, a13);
      const cplx y2 = csub(a02, a13);
      const cplx y1 = cadd(s02, cmulni(s13));  // x0 - i x1 - x2 + i x3
      const cplx y3 = cadd(s02, cmuli(s13));   // x0 + i x1 - x2 - i x3
      p[0] = y0;
      p[q4] = cmul(y1, tw[j * tstep]);
      p[2 * q4] = cmul(y2, tw[2 * j * tstep]);
      p[3 * q4] = cmul(y3, tw[3 * j * tstep]);
    }
    __syncthreads();
  }
}

// exact inverse of fft_forward_dif up to the factor M (unnormalised)
__device__ void fft_inverse_dit(cplx* s, int nr, int M, int logM, const cplx* __restrict__ tw) {
  for (int Ls = 4; Ls <= ((logM & 1) ? (M >> 1) : M); Ls <<= 2) {
    const int q4 = Ls >> 2, nb = M >> 2, tstep = M / Ls;
    for (int idx = threadIdx.x; idx < nr * nb; idx += blockDim.x) {
      const int r = idx / nb, b = idx - r * nb;
      const int blk = b / q4, j = b - blk * q4;
      cplx* p = s + r * M + blk * Ls + j;
      const cplx y0 = p[0];
      const cplx y1 = cmulc(p[q4], tw[j * tstep]);
      const cplx y2 = cmulc(p[2 * q4], tw[2 * j * tstep]);
      const cplx y3 = cmulc(p[3 * q4], tw[3 * j * tstep]);
      const cplx a02 = cadd(y0, y2), s02 = csub(y0, y2), a13 = cadd(y1, y3), s13 = csub(y1, y3);
      p[0] = cadd(a02, a13);
      p[2 * q4] = csub(a02, a13);
      p[q4] = cadd(s02, cmuli(s13));       // y0 + i y1 - y2 - i y3
      p[3 * q4] = cadd(s02, cmulni(s13));  // y0 - i y1 - y2 + i y3
    }
    __syncthreads();
  }
  if (logM & 1) {
    const int h = M >> 1;
    for (int idx = threadIdx.x; idx < nr * h; idx += blockDim.x) {
      const int r = idx / h, j = idx - r * h;
      cplx* p = s + r * M;
      const cplx y0 = p[j], y1 = cmulc(p[j + h], tw[j]);
      p[j] = cadd(y0, y1);
      p[j + h] = csub(y0, y1);
    }
    __syncthreads();
  }
}

// ---- table setup -------------------------------------------------------------
__global__ void fft_fill_tables_kernel(const PxmFftGroup* groups, int ngroups, cplx* arena) {
  const PxmFftGroup gr = groups[blockIdx.x];
  cplx* chirp = arena + gr.chirp_off;
  cplx* tw = arena + gr.tw_off;
  for (int j = threadIdx.x; j < gr.n; j += blockDim.x) {
    const long long r = ((long long)j * j) % (2LL * gr.n);
    double sn, cs;
    sincospi(-(double)r / (double)gr.n, &sn, &cs);
    chirp[j] = make_double2(cs, sn);
  }
  for (int k = threadIdx.x; k < gr.M; k += blockDim.x) {
    double sn, cs;
    sincospi(-2.0 * (double)k / (double)gr.M, &sn, &cs);
    tw[k] = make_double2(cs, sn);
  }
}

__global__ void fft_fill_bhat_kernel(const PxmFftGroup* groups, int ngroups, cplx* arena) {
  extern __shared__ __align__(16) unsigned char fsm[];
  cplx* s = reinterpret_cast<cplx*>(fsm);
  const PxmFftGroup gr = groups[blockIdx.x];
  const cplx* chirp = arena + gr.chirp_off;
  for (int i = threadIdx.x; i < gr.M; i += blockDim.x) s[i] = make_double2(0.0, 0.0);
  __syncthreads();
  for (int j = threadIdx.x; j < gr.n; j += blockDim.x) {
    const cplx c = chirp[j];
    const cplx b = make_double2(c.x, -c.y);
    s[j] = b;
    if (j > 0) s[gr.M - j] = b;
  }
  __syncthreads();
  fft_forward_dif(s, 1, gr.M, gr.logM, arena + gr.tw_off);
  cplx* bhat = arena + gr.bhat_off;
  for (int i = threadIdx.x; i < gr.M; i += blockDim.x) bhat[i] = s[i];
}

// ---- the ring transform ---------------------------------------------------------
// DIR 0: pixels -> ring coefficients  F_m[t] = scale * sum_p f[t,p] e^{-i m phi_p}
// DIR 1: ring coefficients -> pixels  f[t,p] = scale * sum_m F_m[t] e^{+i m phi_p}
template <int DIR>
__global__ void __launch_bounds__(256)
pxm_ring_fft_kernel(const PxmFftGroup* __restrict__ groups, int ngroups, cplx* __restrict__ pix,
                    size_t pix_chain_stride, double* __restrict__ F, int nld, const cplx* __restrict__ arena) {
  extern __shared__ __align__(16) unsigned char fsm[];
  cplx* s = reinterpret_cast<cplx*>(fsm);
  int gi = 0;
  while (gi + 1 < ngroups && (int)blockIdx.x >= groups[gi + 1].cta_begin) ++gi;
  const PxmFftGroup gr = groups[gi];
  const int chain = blockIdx.y;
  const int t0 = ((int)blockIdx.x - gr.cta_begin) * gr.rings_per_cta;
  const int nr = min(gr.rings_per_cta, gr.rings - t0);
  const int n = gr.n, M = gr.M, ell = gr.ell;
  const cplx* chirp = arena + gr.chirp_off;
  const cplx* bhat = arena + gr.bhat_off;
  const cplx* tw = arena + gr.tw_off;
  cplx* mypix = pix + (size_t)chain * pix_chain_stride + gr.pix_off;

  // 1. load, pre-multiply by the chirp, zero-pad
  for (int idx = threadIdx.x; idx < nr * M; idx += blockDim.x) {
    const int r = idx / M, j = idx - r * M;
    cplx v = make_double2(0.0, 0.0);
    if (j < n) {
      const int t = t0 + r;
      if (DIR == 0) {
        v = mypix[(size_t)t * n + j];
      } else {
        const int m = (j < ell) ? j : j - n;
        size_t base;
        double sg = 1.0;
        if (gr.paired) {
          const int am = m < 0 ? -m : m;
          base = gr.f_off + (size_t)am * gr.slot_stride + pxm_il_index(t, chain * 4 + (m < 0 ? 2 : 0), nld);
          if (m < 0 && (am & 1)) sg = -1.0;
        } else {
          base = gr.f_off + (size_t)(m + ell - 1) * gr.slot_stride + pxm_il_index(t, chain * 2, nld);
        }
        // conj on load: x_p = conj( DFT( conj(F) ) )
        v = make_double2(sg * F[base], -sg * F[base + 4]);
      }
      v = cmul(v, chirp[j]);
    }
    s[idx] = v;
  }
  __syncthreads();
  // 2. circular convolution with the chirp filter
  fft_forward_dif(s, nr, M, gr.logM, tw);
  for (int idx = threadIdx.x; idx < nr * M; idx += blockDim.x) {
    const int j = idx & (M - 1);
    s[idx] = cmul(s[idx], bhat[j]);
  }
  __syncthreads();
  fft_inverse_dit(s, nr, M, gr.logM, tw);
  // 3. post-multiply, scale, scatter
  const double sc = gr.scale / (double)M;
  for (int idx = threadIdx.x; idx < nr * n; idx += blockDim.x) {
    const int r = idx / n, k = idx - r * n;
    const int t = t0 + r;
    cplx v = cmul(s[r * M + k], chirp[k]);
    v.x *= sc;
    v.y *= sc;
    if (DIR == 0) {
      const int m = (k < ell) ? k : k - n;
      size_t base;
      if (gr.paired) {
        const int am = m < 0 ? -m : m;
        base = gr.f_off + (size_t)am * gr.slot_stride + pxm_il_index(t, chain * 4 + (m < 0 ? 2 : 0), nld);
        if (m < 0 && (am & 1)) {
          v.x = -v.x;
          v.y = -v.y;
        }
      } else {
        base = gr.f_off + (size_t)(m + ell - 1) * gr.slot_stride + pxm_il_index(t, chain * 2, nld);
      }
      F[base] = v.x;
      F[base + 4] = v.y;  // next column in the k4-interleaved layout
    } else {
      mypix[(size_t)t * n + k] = make_double2(v.x, -v.y);
    }
  }
}

}  // namespace

int pxm_fft_choose_M(int n, int* logM) {
  int M = 16, lg = 4;
  while (M < 2 * n - 1) {
    M <<= 1;
    ++lg;
  }
  *logM = lg;
  return M;
}

int pxm_fft_rings_per_cta(int M) {
  int r = 4096 / M;
  if (r < 1) r = 1;
  if (r > 4) r = 4;
  return r;
}

constexpr int PXM_FFT_SMEM = 4096 * 16;  // 64 KB: up to 4096 complex points per CTA

int pxm_fft_setup_tables(const PxmFftGroup* d_groups, const PxmFftGroup* h_groups, int ngroups, void* d_arena,
                         cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    PXM_CUDA(cudaFuncSetAttribute(fft_fill_bhat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PXM_FFT_SMEM));
    PXM_CUDA(cudaFuncSetAttribute(pxm_ring_fft_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, PXM_FFT_SMEM));
    PXM_CUDA(cudaFuncSetAttribute(pxm_ring_fft_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, PXM_FFT_SMEM));
    configured = true;
  }
  for (int i = 0; i < ngroups; ++i) {
    if (h_groups[i].M > 4096) {
      pxm_set_error("ring FFT: bandlimit too large (Bluestein length > 4096, i.e. L > 1024)");
      return PXM_ERR_UNSUPPORTED;
    }
  }
  fft_fill_tables_kernel<<<ngroups, 256, 0, stream>>>(d_groups, ngroups, (cplx*)d_arena);
  PXM_CUDA(cudaGetLastError());
  fft_fill_bhat_kernel<<<ngroups, 256, PXM_FFT_SMEM, stream>>>(d_groups, ngroups, (cplx*)d_arena);
  PXM_CUDA(cudaGetLastError());
  return PXM_OK;
}

// dir 0: pix -> F ; dir 1: F -> pix
int pxm_fft_launch(int dir, const PxmFftGroup* d_groups, int ngroups, int ctas_per_chain, void* pix,
                   size_t pix_chain_stride, double* F, int nld, const void* d_arena, int nchains,
                   cudaStream_t stream) {
  if (nchains <= 0 || ctas_per_chain <= 0) return PXM_OK;
  dim3 grid(ctas_per_chain, nchains);
  if (dir == 0)
    pxm_ring_fft_kernel<0><<<grid, 256, PXM_FFT_SMEM, stream>>>(d_groups, ngroups, (cplx*)pix, pix_chain_stride, F,
                                                                 nld, (const cplx*)d_arena);
  else
    pxm_ring_fft_kernel<1><<<grid, 256, PXM_FFT_SMEM, stream>>>(d_groups, ngroups, (cplx*)pix, pix_chain_stride, F,
                                                                 nld, (const cplx*)d_arena);
  PXM_CUDA(cudaGetLastError());
  return PXM_OK;
}
