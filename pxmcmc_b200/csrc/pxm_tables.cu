// Plan-time generation of the Legendre tables on the GPU.
//
//  family LAMBDA:  T^m[t, l] = g[l] * (-1)^s sqrt((2l+1)/4pi) d^l_{m,-s}(theta_t)
//                  -> pyssht.inverse / inverse_adjoint (ssht_core_mw_inverse_sov_sym
//                     and its adjoint; reference call sites measurements.py:225,237)
//  family W     :  T^m[t, l] = g[l] * W^{m,s}[l, t], the exact MW-quadrature analysis
//                  operator restricted to the L stored rings
//                  -> pyssht.forward / forward_adjoint (ssht_core_mw_forward_sov_conv_sym
//                     and its adjoint; measurements.py:223,239)
//
// W is built as in SURVEY.md A.2 but without assuming reflection parity of the
// interpolant (oracle/ssht_ref.py::forward_quadrature_matrix documents why):
//   W^m[l, t] = 2 pi sum_{k < 2L} Lambda_{2L}^m[k, l] * Q_par[k, t]
// where Q folds the quadrature weights of the theta-extended fine grid with the
// trigonometric interpolation kernel of the theta-extended coarse ring samples.
// The contraction over k is done by the product's own DMMA Legendre kernel.
#include <algorithm>
#include <vector>

#include "pxm_plan.h"
#include "pxm_wigner.cuh"

namespace {

__global__ void k_wigner_tiles(double* __restrict__ tab, const PxmWigSlot* __restrict__ slots, int nslots, int rings,
                               int grid_L, int lmax, int spin, const double* __restrict__ g,
                               const double* __restrict__ half_angles) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)nslots * rings) return;
  const int si = (int)(idx / rings), t = (int)(idx % rings);
  const PxmWigSlot sl = slots[si];
  if (sl.nlb == 0) return;
  const int am = sl.m < 0 ? -sl.m : sl.m, as = spin < 0 ? -spin : spin;
  const int l0 = am > as ? am : as;
  int lend = am + 16 * (sl.lb0 + sl.nlb);
  if (lend > lmax) lend = lmax;
  double sh, ch;
  if (half_angles) {
    sh = half_angles[2 * t];
    ch = half_angles[2 * t + 1];
  } else {
    pxm_mw_half_angle(t, grid_L, &sh, &ch);
  }
  PxmWigner w;
  pxm_wigner_init(w, sl.m, -spin, sh, ch);
  const double ssign = (as & 1) ? -1.0 : 1.0;
  double* base = tab + sl.tile_off + (size_t)(t >> 5) * (size_t)sl.nlb * PXM_TILE_DOUBLES;
  const int r = t & 31;
  for (int l = l0; l < lend; ++l) {
    const int lam = l - am;
    const int lb = (lam >> 4) - sl.lb0;
    if (lb >= 0) {
      const double v = ssign * sqrt((2.0 * l + 1.0) * 0.07957747154594767 /* 1/(4 pi) */) * pxm_wigner_value(w) *
                       (g ? g[l] : 1.0);
      base[(size_t)lb * PXM_TILE_DOUBLES + pxm_tile_word(r, lam & 15)] = v;
    }
    if (l + 1 < lend) pxm_wigner_step(w);
  }
}

// theta-quadrature weights over the theta-extended MW grid of bandlimit Lf:
//   wr[k] = (1/nf) ( sum_{m even, |m|<Lf} 2 cos(m theta_k)/(1-m^2) + pi sin(theta_k) ),  theta_k=(2k+1)pi/nf
// (real part of the reference's utils.weights_theta, /root/reference/pxmcmc/utils.py:262-267, without 2pi/nf)
__global__ void k_quad_weights(double* __restrict__ wr, int Lf) {
  const int nf = 2 * Lf - 1;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nf) return;
  const long long a = 2LL * k + 1;  // theta_k = pi a / nf
  double acc = 2.0;                 // m = 0 term: w(0) = 2
  for (int m = 2; m < Lf; m += 2) {
    const long long r = ((long long)m * a) % (2LL * nf);
    acc += 2.0 * (2.0 / (1.0 - (double)m * (double)m)) * cospi((double)r / (double)nf);  // +-m pair
  }
  acc += 3.14159265358979323846 * sinpi((double)a / (double)nf);
  wr[k] = acc / (double)nf;
}

// Dirichlet kernel of the coarse extended grid (n = 2*ell-1 points) evaluated at
// x = theta'_k - theta_tc, with x/2 = pi * p / q exactly rational.
__device__ __forceinline__ double dirichlet(long long p, long long q, int n) {
  if (p == 0) return 1.0;
  long long pr = p % (2 * q);
  const double den = sinpi((double)pr / (double)q);
  long long np_ = ((long long)n * (p % (2 * q))) % (2 * q);
  const double num = sinpi((double)np_ / (double)q);
  if (den == 0.0) return 1.0;  // x multiple of 2 pi
  return num / ((double)n * den);
}

// Q_par[k, t] (k < Lf fine rings, t < ell coarse rings), k4-interleaved with nldq columns
__global__ void k_build_Q(double* __restrict__ Q, const double* __restrict__ wr, int ell, int Lf, int par_odd, int nldq) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)Lf * ell) return;
  const int k = (int)(idx / ell), t = (int)(idx % ell);
  const int n = 2 * ell - 1, nf = 2 * Lf - 1;
  const double par = par_odd ? -1.0 : 1.0;
  const long long q = 2LL * n * nf;
  auto interp = [&](int kk) {  // I_t(theta'_kk)
    const long long a = (2LL * kk + 1) * n;
    double v = dirichlet(a - (2LL * t + 1) * nf, q, n);
    if (t < ell - 1) v += par * dirichlet(a - (2LL * (2 * ell - 2 - t) + 1) * nf, q, n);
    return v;
  };
  double v = wr[k] * interp(k);
  if (k < Lf - 1) {
    const int kb = nf - 1 - k;
    v += par * wr[kb] * interp(kb);
  }
  Q[pxm_il_index(k, t, nldq)] = 6.283185307179586476925 * v;
}

// contraction output C[slot][lambda][t] (k4-interleaved) -> final swizzled tiles, x g[l]
__global__ void k_c_to_tiles(const double* __restrict__ C, const unsigned long long* __restrict__ c_slot_off,
                             const PxmWigSlot* __restrict__ slots, int nslots, int ell, int nldq,
                             const double* __restrict__ g, double* __restrict__ tab) {
  const int si = blockIdx.y;
  const PxmWigSlot sl = slots[si];
  if (sl.nlb == 0) return;
  const int am = sl.m < 0 ? -sl.m : sl.m;
  const int nlam = ell - am;
  const long long total = (long long)nlam * ell;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(idx / nlam), lam = (int)(idx % nlam);
    const int lb = (lam >> 4) - sl.lb0;
    if (lb < 0 || lb >= sl.nlb) continue;
    const double v = C[c_slot_off[si] + pxm_il_index(lam, t, nldq)] * (g ? g[lam + am] : 1.0);
    tab[sl.tile_off + ((size_t)(t >> 5) * sl.nlb + lb) * PXM_TILE_DOUBLES + pxm_tile_word(t & 31, lam & 15)] = v;
  }
}

// Gram tiles of the inverse transform: G^m[lam', lam] = scale * sum_t w_t Lambda^m[t, m + lam'] Lambda^m[t, m + lam]
// (rows and columns relative to |m|; one CTA per 32 x 16 output tile, one thread per entry; w == nullptr: w_t = 1)
__global__ void k_gram_tiles(const double* __restrict__ lam_tab, double* __restrict__ g_tab, const PxmWigSlot* __restrict__ lslots,
                             const PxmWigSlot* __restrict__ gslots, int rings, int lmax, int ntb_g, double scale,
                             const double* __restrict__ w) {
  const int si = blockIdx.z;
  const PxmWigSlot ls = lslots[si], gs = gslots[si];
  const int tb = blockIdx.y, lb = blockIdx.x;
  if (gs.nlb == 0 || lb >= gs.nlb || tb >= ntb_g) return;
  const int am = ls.m < 0 ? -ls.m : ls.m;
  const int nrow = lmax - am;
  const int r = threadIdx.x >> 4, c = threadIdx.x & 15;
  const int lp = 32 * tb + r, l = 16 * lb + c;  // relative degrees
  double acc = 0.0;
  if (lp < nrow && l < nrow) {
    const int lbp = lp >> 4, cp = lp & 15;
    for (int t = 0; t < rings; ++t) {
      const double* tile_row = lam_tab + ls.tile_off + (size_t)(t >> 5) * (size_t)ls.nlb * PXM_TILE_DOUBLES;
      const double a = tile_row[(size_t)(lbp - ls.lb0) * PXM_TILE_DOUBLES + pxm_tile_word(t & 31, cp)];
      const double b = tile_row[(size_t)(lb - ls.lb0) * PXM_TILE_DOUBLES + pxm_tile_word(t & 31, c)];
      acc += (w ? w[t] : 1.0) * (a * b);
    }
  }
  g_tab[gs.tile_off + ((size_t)tb * gs.nlb + lb) * PXM_TILE_DOUBLES + pxm_tile_word(r, c)] = scale * acc;
}

}  // namespace

void pxm_make_table_layout(PxmTableLayout& T, int grid_L, int rings, int lmax, int spin, int l_lo, int l_hi,
                           unsigned long long base_off, int rank, int world) {
  T.grid_L = grid_L;
  T.rings = rings;
  T.lmax = lmax;
  T.spin = spin;
  T.paired = (spin == 0);
  T.l_lo = std::max(l_lo, 0);
  T.l_hi = std::min(l_hi, lmax);
  T.ntb = pxm_ceil_div(rings, PXM_TILE_T);
  T.slot_m.clear();
  if (T.paired)
    for (int m = 0; m < lmax; ++m) T.slot_m.push_back(m);
  else
    for (int m = -(lmax - 1); m < lmax; ++m) T.slot_m.push_back(m);
  T.nslots = (int)T.slot_m.size();
  T.lb0.assign(T.nslots, 0);
  T.nlb.assign(T.nslots, 0);
  T.tile_off.assign(T.nslots, 0);
  unsigned long long off = base_off;
  const int as = spin < 0 ? -spin : spin;
  for (int s = 0; s < T.nslots; ++s) {
    const int am = std::abs(T.slot_m[s]);
    const int first = std::max(std::max(am, as), T.l_lo), last = T.l_hi;
    T.tile_off[s] = off;
    if (first >= last || pxm_owner_of_m(am, world) != rank) continue;
    T.lb0[s] = (first - am) / PXM_TILE_L;
    T.nlb[s] = (last - 1 - am) / PXM_TILE_L + 1 - T.lb0[s];
    off += (unsigned long long)T.ntb * T.nlb[s] * PXM_TILE_DOUBLES;
  }
  T.doubles = off - base_off;
}

static std::vector<PxmWigSlot> wig_slots(const PxmTableLayout& T, int s0, int s1) {
  std::vector<PxmWigSlot> v;
  for (int s = s0; s < s1; ++s) {
    PxmWigSlot w;
    w.m = T.slot_m[s];
    w.lb0 = T.lb0[s];
    w.nlb = T.nlb[s];
    w.pad = 0;
    w.tile_off = T.tile_off[s];
    v.push_back(w);
  }
  return v;
}

int pxm_generate_lambda(const PxmTableLayout& T, double* d_tab, const double* d_g, cudaStream_t st) {
  std::vector<PxmWigSlot> hs = wig_slots(T, 0, T.nslots);
  PxmDevVec<PxmWigSlot> ds;
  PXM_TRY(ds.upload(hs));
  const long long total = (long long)T.nslots * T.rings;
  k_wigner_tiles<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(d_tab, ds.d, T.nslots, T.rings, T.grid_L, T.lmax,
                                                                   T.spin, d_g, T.d_half_angles);
  PXM_LAUNCHED();
  PXM_CUDA(cudaStreamSynchronize(st));
  ds.release();
  return PXM_OK;
}

int pxm_generate_w(const PxmTableLayout& T, double* d_tab, const double* d_g, cudaStream_t st) {
  const int ell = T.lmax;
  if (T.grid_L != ell || T.rings != ell) {
    pxm_set_error("W tables need grid_L == lmax");
    return PXM_ERR_ARG;
  }
  const int Lf = 2 * ell, nf = 2 * Lf - 1;
  const int nldq = pxm_legendre_pad_columns(ell);
  const int rows_q = pxm_round_up(Lf, 32);
  // quadrature x interpolation matrices for both parities
  double *d_wr = nullptr, *d_Q = nullptr;
  PXM_CUDA(cudaMalloc(&d_wr, sizeof(double) * nf));
  const size_t qsz = (size_t)rows_q * nldq;
  PXM_CUDA(cudaMalloc(&d_Q, sizeof(double) * qsz * 2));
  PXM_CUDA(cudaMemsetAsync(d_Q, 0, sizeof(double) * qsz * 2, st));
  k_quad_weights<<<(nf + 127) / 128, 128, 0, st>>>(d_wr, Lf);
  PXM_LAUNCHED();
  for (int par = 0; par < 2; ++par) {
    const long long tot = (long long)Lf * ell;
    k_build_Q<<<(unsigned)((tot + 127) / 128), 128, 0, st>>>(d_Q + par * qsz, d_wr, ell, Lf, par, nldq);
    PXM_LAUNCHED();
  }
  // process slots in chunks to bound the temporary fine-grid tables
  const size_t budget = (size_t)96 << 20;  // doubles (768 MB) for fine tables + C
  int s0 = 0;
  const int as = std::abs(T.spin);
  while (s0 < T.nslots) {
    // fine layout for slots [s0, s1)
    int s1 = s0;
    size_t need = 0;
    std::vector<unsigned long long> c_off;
    PxmTableLayout F;  // fine-grid Lambda for this chunk (only slots s0..s1)
    F.grid_L = Lf;
    F.rings = Lf;
    F.lmax = ell;
    F.spin = T.spin;
    F.paired = T.paired;
    F.ntb = pxm_ceil_div(Lf, PXM_TILE_T);
    F.slot_m.clear();
    F.lb0.clear();
    F.nlb.clear();
    F.tile_off.clear();
    size_t fine_doubles = 0, c_doubles = 0;
    while (s1 < T.nslots) {
      const int am = std::abs(T.slot_m[s1]);
      const int first = std::max(am, as);
      // slots without tiles (outside the kernel's l-support, or owned by another rank) are skipped
      const int nlb = (first < ell && T.nlb[s1] > 0) ? (ell - 1 - am) / PXM_TILE_L + 1 : 0;
      const size_t fd = (size_t)F.ntb * nlb * PXM_TILE_DOUBLES;
      const size_t cd = (size_t)pxm_round_up(std::max(ell - am, 1), 64) * nldq;
      if (s1 > s0 && need + fd + cd > budget) break;
      F.slot_m.push_back(T.slot_m[s1]);
      F.lb0.push_back(0);
      F.nlb.push_back(nlb);
      F.tile_off.push_back(fine_doubles);
      c_off.push_back(c_doubles);
      fine_doubles += fd;
      c_doubles += cd;
      need += fd + cd;
      ++s1;
    }
    F.nslots = s1 - s0;
    F.doubles = fine_doubles;
    double *d_fine = nullptr, *d_C = nullptr;
    PXM_CUDA(cudaMalloc(&d_fine, sizeof(double) * std::max<size_t>(fine_doubles, 1)));
    PXM_CUDA(cudaMalloc(&d_C, sizeof(double) * std::max<size_t>(c_doubles, 1)));
    PXM_CUDA(cudaMemsetAsync(d_fine, 0, sizeof(double) * std::max<size_t>(fine_doubles, 1), st));
    PXM_CUDA(cudaMemsetAsync(d_C, 0, sizeof(double) * std::max<size_t>(c_doubles, 1), st));
    PXM_TRY(pxm_generate_lambda(F, d_fine, nullptr, st));
    // contraction over the fine rings:  C[lambda, t] = sum_k Lambda'[k, lambda] Q[k, t]
    std::vector<PxmLegItem> items;
    std::vector<PxmLegSeg> segs;
    for (int i = 0; i < F.nslots; ++i) {
      if (F.nlb[i] == 0) continue;
      const int par = (std::abs(F.slot_m[i] + T.spin)) & 1;
      for (int lt = 0; lt * 4 < F.nlb[i]; ++lt) {
        PxmLegSeg sg;
        sg.a_off = F.tile_off[i] + (unsigned long long)(lt * 4) * PXM_TILE_DOUBLES;
        sg.b_off = (unsigned long long)par * qsz;
        sg.a_kstride = F.nlb[i] * PXM_TILE_DOUBLES;
        sg.a_mstride = PXM_TILE_DOUBLES;
        sg.mt0 = 0;
        sg.nmt = std::min(4, F.nlb[i] - lt * 4);
        sg.nk = F.ntb;
        sg.src = 0;
        PxmLegItem it;
        it.c_off = c_off[i] + (unsigned long long)(lt * 64) * nldq;
        it.seg_begin = (int)segs.size();
        it.seg_count = 1;
        it.nmt_out = sg.nmt;
        it.cost = sg.nk * sg.nmt;
        it.dst = 0;
        it.pad = 0;
        segs.push_back(sg);
        items.push_back(it);
      }
    }
    PxmDevVec<PxmLegItem> di;
    PxmDevVec<PxmLegSeg> dsg;
    PXM_TRY(di.upload(items));
    PXM_TRY(dsg.upload(segs));
    PXM_TRY(pxm_legendre_launch(1, d_fine, d_Q, d_C, di.d, dsg.d, (int)items.size(), nldq, st, pxm_debug_naive()));
    // scatter into the final tiles
    std::vector<PxmWigSlot> fs = wig_slots(T, s0, s1);
    PxmDevVec<PxmWigSlot> dfs;
    PxmDevVec<unsigned long long> dco;
    PXM_TRY(dfs.upload(fs));
    PXM_TRY(dco.upload(c_off));
    dim3 grid(64, F.nslots);
    k_c_to_tiles<<<grid, 256, 0, st>>>(d_C, dco.d, dfs.d, F.nslots, ell, nldq, d_g, d_tab);
    PXM_LAUNCHED();
    PXM_CUDA(cudaStreamSynchronize(st));
    di.release();
    dsg.release();
    dfs.release();
    dco.release();
    cudaFree(d_fine);
    cudaFree(d_C);
    s0 = s1;
  }
  cudaFree(d_wr);
  cudaFree(d_Q);
  return PXM_OK;
}

// host-callable evaluation of one table row with the same recurrence code
// (used by the CPU test-suite to check the recurrence without a GPU)
extern "C" int pxm_debug_wigner_row_host(int grid_L, int ring, int m, int spin, int lmax, double* out) {
  const int am = m < 0 ? -m : m, as = spin < 0 ? -spin : spin;
  const int l0 = am > as ? am : as;
  double sh, ch;
  pxm_mw_half_angle(ring, grid_L, &sh, &ch);
  for (int l = 0; l < lmax; ++l) out[l] = 0.0;
  if (l0 >= lmax) return PXM_OK;
  PxmWigner w;
  pxm_wigner_init(w, m, -spin, sh, ch);
  const double ssign = (as & 1) ? -1.0 : 1.0;
  for (int l = l0; l < lmax; ++l) {
    out[l] = ssign * sqrt((2.0 * l + 1.0) * 0.07957747154594767) * pxm_wigner_value(w);
    if (l + 1 < lmax) pxm_wigner_step(w);
  }
  return PXM_OK;
}

// G^m = scale * Lambda^T Lambda from the (already generated, unweighted, full-support) Lambda table `Tl` in `d_lam_tab`
// into the layout `Tg` (same slots; rows = relative degrees) in `d_g_tab`
int pxm_generate_gram(const PxmTableLayout& Tl, const double* d_lam_tab, const PxmTableLayout& Tg, double* d_g_tab, double scale,
                      cudaStream_t st, const double* d_ring_weights) {
  if (Tl.nslots != Tg.nslots || Tl.l_lo != 0 || Tg.l_lo != 0) {
    pxm_set_error("gram tables need full-support Lambda tables with the same slots");
    return PXM_ERR_ARG;
  }
  std::vector<PxmWigSlot> hl = wig_slots(Tl, 0, Tl.nslots), hg = wig_slots(Tg, 0, Tg.nslots);
  PxmDevVec<PxmWigSlot> dl, dg;
  PXM_TRY(dl.upload(hl));
  PXM_TRY(dg.upload(hg));
  int maxnlb = 0;
  for (int v : Tg.nlb) maxnlb = std::max(maxnlb, v);
  dim3 grid(std::max(maxnlb, 1), Tg.ntb, Tg.nslots);
  k_gram_tiles<<<grid, 512, 0, st>>>(d_lam_tab, d_g_tab, dl.d, dg.d, Tl.rings, Tl.lmax, Tg.ntb, scale, d_ring_weights);
  PXM_LAUNCHED();
  PXM_CUDA(cudaStreamSynchronize(st));
  dl.release();
  dg.release();
  return PXM_OK;
}
