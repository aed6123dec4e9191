// Great-circle path rasteriser: the rows of the path matrix of the phase-velocity experiment.
//
// Replaces greatcirclepaths.GreatCirclePath(start, stop, "MW", L=L, weighting="average", latlon=True)
// .get_points(points_per_rad).fill() -> .map as called by /root/reference/experiments/phasevel/main.py:40-47
// (one Python object and one dense L x (2L-1) map per path there, 16 worker processes).  Here one CTA
// rasterises one path: points evenly spaced on the minor arc (spherical linear interpolation, end points
// included), each binned to its nearest MW pixel (pyssht.theta_to_index / phi_to_index), pixels sorted and
// run-length encoded in shared memory, weight = the pixel's share of the path's points (rows sum to one).
// Output: padded rows (sorted unique columns + weights) and a compaction kernel that packs them into CSR.
#include <math_constants.h>

#include "pxm_common.cuh"

namespace {

constexpr int GC_THREADS = 128;

struct GcPath {
  double a[3], b[3];  // unit vectors of the end points
  double d;           // epicentral distance
  int n;              // number of points
};

__device__ __forceinline__ void gc_unit(double lat_deg, double lon_deg, double* v) {
  const double lat = lat_deg * (CUDART_PI / 180.0), lon = lon_deg * (CUDART_PI / 180.0);
  double sl, cl, so, co;
  sincos(lat, &sl, &cl);
  sincos(lon, &so, &co);
  v[0] = cl * co;
  v[1] = cl * so;
  v[2] = sl;
}

__device__ __forceinline__ GcPath gc_path(const double* start, const double* stop, long long row, double ppr) {
  GcPath p;
  gc_unit(start[2 * row], start[2 * row + 1], p.a);
  gc_unit(stop[2 * row], stop[2 * row + 1], p.b);
  double dot = p.a[0] * p.b[0] + p.a[1] * p.b[1] + p.a[2] * p.b[2];
  dot = fmin(1.0, fmax(-1.0, dot));
  p.d = acos(dot);
  const int n = (int)ceil(ppr * p.d);
  p.n = n < 2 ? 2 : n;
  return p;
}

// flat MW pixel of point i of the path
__device__ __forceinline__ int gc_pixel(const GcPath& p, int i, int L) {
  const double f = (double)i / ((double)p.n - 1.0);
  double x, y, z;
  if (p.d < 1e-12) {
    x = p.a[0];
    y = p.a[1];
    z = p.a[2];
  } else {
    const double sa = sin((1.0 - f) * p.d), sb = sin(f * p.d), sd = sin(p.d);
    x = (sa * p.a[0] + sb * p.b[0]) / sd;
    y = (sa * p.a[1] + sb * p.b[1]) / sd;
    z = (sa * p.a[2] + sb * p.b[2]) / sd;
  }
  const double theta = acos(fmin(1.0, fmax(-1.0, z)));
  double phi = atan2(y, x);
  if (phi < 0.0) phi += 2.0 * CUDART_PI;
  if (phi >= 2.0 * CUDART_PI) phi -= 2.0 * CUDART_PI;
  const int n = 2 * L - 1;
  int t = (int)floor((theta * n / CUDART_PI - 1.0) / 2.0 + 0.5);
  t = t < 0 ? 0 : (t > L - 1 ? L - 1 : t);
  int q = (int)floor(phi * n / (2.0 * CUDART_PI) + 0.5) % n;
  return t * n + q;
}

__global__ void k_gc_count(const double* __restrict__ start, const double* __restrict__ stop, long long npaths,
                           double ppr, int* __restrict__ npoints) {
  const long long row = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (row < npaths) npoints[row] = gc_path(start, stop, row, ppr).n;
}

// one CTA per path; `cap` (a power of two >= the largest point count) slots of dynamic shared memory
__global__ void __launch_bounds__(GC_THREADS)
k_gc_rasterise(const double* __restrict__ start, const double* __restrict__ stop, long long npaths, int L, double ppr,
               int cap, int* __restrict__ cols, double* __restrict__ w, int* __restrict__ nnz) {
  extern __shared__ int gsm[];
  int* v = gsm;          // [cap] pixel of every point, then sorted
  int* pos = gsm + cap;  // [cap] rank of every run start
  __shared__ int s_runs;
  const long long row = blockIdx.x;
  if (row >= npaths) return;
  const GcPath p = gc_path(start, stop, row, ppr);
  for (int i = threadIdx.x; i < cap; i += blockDim.x) v[i] = i < p.n ? gc_pixel(p, i, L) : 0x7fffffff;
  __syncthreads();
  // bitonic sort, ascending (padding sorts to the end)
  for (int k = 2; k <= cap; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < cap; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const int a = v[i], b = v[ixj];
          const bool up = (i & k) == 0;
          if ((a > b) == up) {
            v[i] = b;
            v[ixj] = a;
          }
        }
      }
      __syncthreads();
    }
  }
  // run starts -> ranks (serial scan by one thread per chunk would do; cap <= 1024: a simple two-level scan)
  if (threadIdx.x == 0) {
    int r = 0;
    for (int i = 0; i < p.n; ++i) {
      const bool st = (i == 0) || (v[i] != v[i - 1]);
      pos[i] = st ? r : -1;
      r += st ? 1 : 0;
    }
    s_runs = r;
  }
  __syncthreads();
  const int runs = s_runs;
  int* crow = cols + row * (long long)cap;
  double* wrow = w + row * (long long)cap;
  for (int i = threadIdx.x; i < p.n; i += blockDim.x) {
    if (pos[i] >= 0) {
      int e = i + 1;
      while (e < p.n && v[e] == v[i]) ++e;
      crow[pos[i]] = v[i];
      wrow[pos[i]] = (double)(e - i) / (double)p.n;
    }
  }
  for (int i = runs + threadIdx.x; i < cap; i += blockDim.x) crow[i] = -1;
  if (threadIdx.x == 0) nnz[row] = runs;
}

__global__ void k_gc_compact(const long long* __restrict__ indptr, const int* __restrict__ cols,
                             const double* __restrict__ w, int cap, long long npaths, int* __restrict__ indices,
                             double* __restrict__ data) {
  const long long row = blockIdx.x;
  if (row >= npaths) return;
  const long long o = indptr[row];
  const int n = (int)(indptr[row + 1] - o);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    indices[o + i] = cols[row * (long long)cap + i];
    data[o + i] = w[row * (long long)cap + i];
  }
}

}  // namespace

extern "C" {

int pxm_gc_count_points(const double* d_start, const double* d_stop, long long npaths, double points_per_rad,
                        int* d_npoints, void* stream) {
  PXM_REQUIRE(npaths >= 0 && points_per_rad > 0, "pxm_gc_count_points");
  if (!npaths) return PXM_OK;
  k_gc_count<<<(unsigned)((npaths + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_start, d_stop, npaths, points_per_rad,
                                                                               d_npoints);
  PXM_LAUNCHED();
  return PXM_OK;
}

int pxm_gc_rasterise(const double* d_start, const double* d_stop, long long npaths, int L, double points_per_rad,
                     int cap, int* d_cols, double* d_w, int* d_nnz, void* stream) {
  PXM_REQUIRE(npaths >= 0 && L >= 1 && points_per_rad > 0, "pxm_gc_rasterise");
  PXM_REQUIRE(cap >= 2 && (cap & (cap - 1)) == 0 && cap <= 4096, "pxm_gc_rasterise: cap must be a power of two <= 4096");
  PXM_REQUIRE((double)cap >= ceil(points_per_rad * CUDART_PI), "pxm_gc_rasterise: cap smaller than the longest possible path");
  if (!npaths) return PXM_OK;
  k_gc_rasterise<<<(unsigned)npaths, GC_THREADS, 2 * cap * sizeof(int), (cudaStream_t)stream>>>(
      d_start, d_stop, npaths, L, points_per_rad, cap, d_cols, d_w, d_nnz);
  PXM_LAUNCHED();
  return PXM_OK;
}

int pxm_gc_compact(const long long* d_indptr, const int* d_cols, const double* d_w, int cap, long long npaths,
                   int* d_indices, double* d_data, void* stream) {
  if (!npaths) return PXM_OK;
  k_gc_compact<<<(unsigned)npaths, 128, 0, (cudaStream_t)stream>>>(d_indptr, d_cols, d_w, cap, npaths, d_indices, d_data);
  PXM_LAUNCHED();
  return PXM_OK;
}

}  // extern "C"
