// Reconstruction metrics of the evaluation path on the device.
//
// Reference: DeepGenerativeModelMixin.eval_x_reconstruction, src/DGMMixin.py:128-158 — numpy float64 RMSE,
// sklearn.metrics.r2_score(multioutput='variance_weighted'), a Python loop of scipy.stats.pearsonr over the rows, and
// the mean per-row Gaussian log-likelihood (src/blocks.py:230-234).  It runs on the full train and validation sets
// every epoch (src/DrVAE.py:784-821); once the training step takes microseconds that loop dominates an epoch.
//
// Here: three small kernels, all arithmetic in fp64 like the reference, every reduction in a fixed order
// (deterministic):
//   eval_rows_kernel   one warp per selected row: sums of x, r, x^2, r^2, x r, (x - r)^2 and of the log-density terms
//                      -> Pearson r, squared error and log-likelihood of the row
//   eval_cols_kernel   one thread per feature: sum and sum of squares of x over the selected rows (the variance-weighted
//                      R^2 needs the total sum of squares around the per-feature means)
//   eval_final_kernel  one block: fixed-order sums of the row and column statistics -> {rmse, r2, pearr, ll}
#include <cuda_runtime.h>
#include <math.h>

#include "drvae_b200.h"
#include "errors.h"

namespace {

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// rowstats[i] = {pearson r, sum (x - r)^2, log-likelihood, selected ? 1 : 0}
__global__ void __launch_bounds__(256) eval_rows_kernel(const float* __restrict__ x, const float* __restrict__ r, const float* __restrict__ sg,
                                                        const int* __restrict__ mask, int N, int X, double* __restrict__ rowstats) {
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= N) return;
  double* o = rowstats + 4 * (long long)i;
  if (mask && mask[i] == 0) {
    if (lane < 4) o[lane] = 0.0;
    return;
  }
  const float* xr = x + (long long)i * X;
  const float* rr = r + (long long)i * X;
  const float* sr = sg ? sg + (long long)i * X : nullptr;
  double sx = 0, sy = 0, sxx = 0, syy = 0, sxy = 0, se = 0, ll = 0;
  const double LOG2PI = 1.8378770664093454836;
  for (int f = lane; f < X; f += 32) {
    const double a = xr[f], b = rr[f];
    sx += a;
    sy += b;
    sxx += a * a;
    syy += b * b;
    sxy += a * b;
    const double d = a - b;
    se += d * d;
    if (sr) {
      const double s = sr[f];
      ll += -0.5 * (LOG2PI + log(s * s) + d * d / (s * s));
    }
  }
  sx = warp_sum_d(sx);
  sy = warp_sum_d(sy);
  sxx = warp_sum_d(sxx);
  syy = warp_sum_d(syy);
  sxy = warp_sum_d(sxy);
  se = warp_sum_d(se);
  ll = warp_sum_d(ll);
  if (lane == 0) {
    const double n = X;
    const double cov = sxy - sx * sy / n, vx = sxx - sx * sx / n, vy = syy - sy * sy / n;
    o[0] = cov / sqrt(vx * vy);  // zero variance -> nan, like scipy.stats.pearsonr
    o[1] = se;
    o[2] = sr ? ll : nan("");
    o[3] = 1.0;
  }
}

// colstats[f] = {sum_i x[i][f], sum_i x[i][f]^2} over the selected rows
__global__ void __launch_bounds__(128) eval_cols_kernel(const float* __restrict__ x, const int* __restrict__ mask, int N, int X,
                                                        double* __restrict__ colstats) {
  const int f = blockIdx.x * 128 + threadIdx.x;
  if (f >= X) return;
  double s = 0, ss = 0;
  for (int i = 0; i < N; ++i) {
    if (mask && mask[i] == 0) continue;
    const double a = x[(long long)i * X + f];
    s += a;
    ss += a * a;
  }
  colstats[2 * f] = s;
  colstats[2 * f + 1] = ss;
}

__device__ double block_sum_d(double v, double* sm) {
  const int t = threadIdx.x;
  sm[t] = v;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (t < o) sm[t] += sm[t + o];
    __syncthreads();
  }
  const double r = sm[0];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(256) eval_final_kernel(const double* __restrict__ rowstats, const double* __restrict__ colstats, int N, int X,
                                                         double* __restrict__ out) {
  __shared__ double sm[256];
  const int t = threadIdx.x;
  double pr = 0, se = 0, ll = 0, cnt = 0;
  for (int i = t; i < N; i += 256) {
    const double* o = rowstats + 4 * (long long)i;
    if (o[3] != 0.0) {
      pr += o[0];
      se += o[1];
      ll += o[2];
      cnt += 1.0;
    }
  }
  pr = block_sum_d(pr, sm);
  se = block_sum_d(se, sm);
  ll = block_sum_d(ll, sm);
  cnt = block_sum_d(cnt, sm);
  double tot = 0;
  for (int f = t; f < X; f += 256) tot += colstats[2 * f + 1] - colstats[2 * f] * colstats[2 * f] / cnt;
  tot = block_sum_d(tot, sm);
  if (t == 0) {
    out[0] = cnt > 0 ? sqrt(se / (cnt * X)) : nan("");
    out[1] = (cnt > 0 && tot > 0) ? 1.0 - se / tot : nan("");
    out[2] = cnt > 0 ? pr / cnt : nan("");
    out[3] = cnt > 0 ? ll / cnt : nan("");
  }
}

}  // namespace

extern "C" long long drvae_eval_workspace_bytes(int N, int X) {
  if (N < 1 || X < 1) return -1;
  return (long long)sizeof(double) * (4LL * N + 2LL * X);
}

extern "C" int drvae_eval_x_reconstruction(const float* x, const float* x_rec, const float* x_sigma, const int* mask, int N, int X,
                                           double* out, void* workspace, void* stream) {
  if (!x || !x_rec || !out || !workspace) return drvae::set_error("drvae_eval_x_reconstruction: null argument");
  if (N < 1 || X < 1) return drvae::set_error("drvae_eval_x_reconstruction: empty input");
  cudaStream_t st = (cudaStream_t)stream;
  double* rowstats = reinterpret_cast<double*>(workspace);
  double* colstats = rowstats + 4LL * N;
  eval_rows_kernel<<<(N + 7) / 8, 256, 0, st>>>(x, x_rec, x_sigma, mask, N, X, rowstats);
  eval_cols_kernel<<<(X + 127) / 128, 128, 0, st>>>(x, mask, N, X, colstats);
  eval_final_kernel<<<1, 256, 0, st>>>(rowstats, colstats, N, X, out);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return drvae::set_cuda_error("drvae_eval_x_reconstruction", err);
  return 0;
}
