// Host check of the register DFT codelets of the ring FFT (pxmcmc_b200/csrc/pxm_dft.cuh) against a direct O(N^2)
// DFT in long double.  Built and run by tests/test_host_cpu.py (nvcc compiles the PXM_HD functions for the host).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "pxm_dft.cuh"

static long double PI = acosl(-1.0L);

static void naive(const std::vector<cplx>& x, std::vector<cplx>& y, int n, bool inv, int nin, int stride_out, int nout) {
  // y[q] = sum_{r < nin} x[r] exp(-+ 2 pi i r q / n), q = 0, stride_out, ...
  y.assign(nout, make_double2(0, 0));
  for (int qi = 0; qi < nout; ++qi) {
    const int q = qi * stride_out;
    long double re = 0, im = 0;
    for (int r = 0; r < nin; ++r) {
      const long double a = (inv ? 2.0L : -2.0L) * PI * (long double)((r * q) % n) / n;
      const long double c = cosl(a), s = sinl(a);
      re += x[r].x * c - x[r].y * s;
      im += x[r].x * s + x[r].y * c;
    }
    y[qi] = make_double2((double)re, (double)im);
  }
}

static double err(const cplx* a, const std::vector<cplx>& b, int n) {
  double e = 0, nb = 0;
  for (int i = 0; i < n; ++i) {
    e += (a[i].x - b[i].x) * (a[i].x - b[i].x) + (a[i].y - b[i].y) * (a[i].y - b[i].y);
    nb += b[i].x * b[i].x + b[i].y * b[i].y;
  }
  return sqrt(e / nb);
}

template <int R, bool INV>
static double check_full() {
  std::vector<cplx> x(R), y;
  for (int i = 0; i < R; ++i) x[i] = make_double2(drand48() - 0.5, drand48() - 0.5);
  naive(x, y, R, INV, R, 1, R);
  cplx w[R];
  for (int i = 0; i < R; ++i) w[i] = x[i];
  dftN<R, INV>(w);
  return err(w, y, R);
}
template <int R>
static double check_half_in() {
  std::vector<cplx> x(R), y;
  for (int i = 0; i < R; ++i) x[i] = i < R / 2 ? make_double2(drand48() - 0.5, drand48() - 0.5) : make_double2(0, 0);
  naive(x, y, R, false, R / 2, 1, R);
  cplx w[R];
  for (int i = 0; i < R; ++i) w[i] = x[i];
  dft_half_in<R>(w);
  return err(w, y, R);
}
template <int R>
static double check_half_out() {
  std::vector<cplx> x(R), y;
  for (int i = 0; i < R; ++i) x[i] = make_double2(drand48() - 0.5, drand48() - 0.5);
  naive(x, y, R, true, R, 1, R / 2);
  cplx w[R];
  for (int i = 0; i < R; ++i) w[i] = x[i];
  dft_half_out<R>(w);
  return err(w, y, R / 2);
}

int main() {
  srand48(7);
  double worst = 0;
  auto rec = [&](const char* name, double e) {
    printf("%-22s %.3e\n", name, e);
    if (!(e <= worst)) worst = e;
  };
  for (int rep = 0; rep < 20; ++rep) {
    rec("dft2 fwd", check_full<2, false>());
    rec("dft4 fwd", check_full<4, false>());
    rec("dft4 inv", check_full<4, true>());
    rec("dft8 fwd", check_full<8, false>());
    rec("dft8 inv", check_full<8, true>());
    rec("dft16 fwd", check_full<16, false>());
    rec("dft16 inv", check_full<16, true>());
    rec("dft32 fwd", check_full<32, false>());
    rec("dft32 inv", check_full<32, true>());
    rec("half_in 16", check_half_in<16>());
    rec("half_in 32", check_half_in<32>());
    rec("half_out 16", check_half_out<16>());
    rec("half_out 32", check_half_out<32>());
  }
  printf("worst %.3e\n", worst);
  return worst < 2e-15 ? 0 : 1;
}
