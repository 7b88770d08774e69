// Two-component 1-D Gaussian mixture EM on the device.
// Replaces the scikit-learn fit inside divide_dataset / get_gmm_threshold
// ("#clean 분포와 noisy 분포가 만나는 지점의 loss보다 작은 데.py:290-292", "# 종합 loss.py:271-273":
//  GaussianMixture(n_components=2, max_iter=10, tol=1e-2, reg_covar=5e-4).fit(losses.reshape(-1, 1))).
//
// Same EM as sklearn.mixture (full covariance, 1 feature): E-step log-responsibilities through logsumexp,
// M-step nk = sum r + 10 eps, mu = sum r x / nk, var = sum r (x - mu)^2 / nk + reg_covar, w = nk / n, stop when
// the mean log-likelihood changes by less than tol.  DOCUMENTED DEVIATION: scikit-learn seeds EM with a k-means
// run drawn from the global numpy RNG; here the k-means (Lloyd) iterations start from the 25 % / 75 % order
// statistics, so the fit is deterministic and identical on every rank.  On a bimodal loss vector both reach the
// same optimum (tests compare against a seeded scikit-learn fit).
//
// One streaming pass per iteration: every CTA reduces its slice to 8 doubles (fixed order inside the CTA), the
// last CTA to finish adds the per-CTA partials in CTA order (reproducible) -> sums[8].  A one-thread update
// kernel turns the (all-reduced, when sharded) sums into the next parameters; iterations after convergence are
// no-ops, so the host enqueues max_iter rounds without reading anything back.
#include <float.h>

#include "common.cuh"

namespace sg {
namespace gmm {

// device state (doubles): [0..1] weights, [2..3] means, [4..5] variances, [6] lower bound, [7] n_iter,
// [8] converged flag, [9] phase (0 = k-means, 1 = EM, 2 = done), [10] k-means iterations left
constexpr int S_W = 0, S_MU = 2, S_VAR = 4, S_LB = 6, S_ITER = 7, S_CONV = 8, S_PHASE = 9, S_KLEFT = 10, S_WORDS = 16;
constexpr int kThreads = 256;
constexpr int kSums = 8;   // r0, r0 x, r0 x^2, r1, r1 x, r1 x^2, sum log p(x), (k-means: moved flag unused)

__global__ void __launch_bounds__(kThreads) accumulate_kernel(const float* __restrict__ v, int64_t n,
                                                              const double* __restrict__ state,
                                                              double* __restrict__ partials, double* __restrict__ sums,
                                                              unsigned int* __restrict__ ticket) {
  __shared__ double s_red[kThreads / 32][kSums];
  __shared__ int s_last;
  const int phase = (int)state[S_PHASE];
  if (phase == 2) return;
  const double mu0 = state[S_MU], mu1 = state[S_MU + 1];
  double acc[kSums];
#pragma unroll
  for (int j = 0; j < kSums; ++j) acc[j] = 0.0;
  if (phase == 0) {
    // Lloyd step: hard assignment to the nearer centre (ties -> component 0, as argmin does)
    for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
      const double x = (double)v[i];
      const bool c1 = fabs(x - mu1) < fabs(x - mu0);
      if (c1) { acc[3] += 1.0; acc[4] += x; acc[5] += x * x; }
      else { acc[0] += 1.0; acc[1] += x; acc[2] += x * x; }
    }
  } else {
    const double w0 = state[S_W], w1 = state[S_W + 1], v0 = state[S_VAR], v1 = state[S_VAR + 1];
    const double kLog2Pi = 1.8378770664093453;
    const double c0 = log(w0) - 0.5 * (kLog2Pi + log(v0)), c1 = log(w1) - 0.5 * (kLog2Pi + log(v1));
    const double i0 = 0.5 / v0, i1 = 0.5 / v1;
    for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
      const double x = (double)v[i];
      const double l0 = c0 - (x - mu0) * (x - mu0) * i0, l1 = c1 - (x - mu1) * (x - mu1) * i1;
      const double m = fmax(l0, l1);
      const double lse = m + log(exp(l0 - m) + exp(l1 - m));
      const double r0 = exp(l0 - lse), r1 = exp(l1 - lse);
      acc[0] += r0; acc[1] += r0 * x; acc[2] += r0 * x * x;
      acc[3] += r1; acc[4] += r1 * x; acc[5] += r1 * x * x;
      acc[6] += lse;
    }
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < kSums; ++j) {
    double x = acc[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) s_red[w][j] = x;
  }
  __syncthreads();
  if (threadIdx.x < kSums) {
    double x = 0.0;
    for (int ww = 0; ww < kThreads / 32; ++ww) x += s_red[ww][threadIdx.x];
    partials[(size_t)blockIdx.x * kSums + threadIdx.x] = x;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (s_last) {
    __threadfence();
    if (threadIdx.x < kSums) {
      double x = 0.0;
      for (unsigned b = 0; b < gridDim.x; ++b) x += __ldcg(partials + (size_t)b * kSums + threadIdx.x);   // CTA order
      sums[threadIdx.x] = x;
    }
    if (threadIdx.x == 0) *ticket = 0u;
  }
}

// sums -> next parameters (one thread).  n_total = global element count (all ranks).
__global__ void update_kernel(const double* __restrict__ sums, double n_total, double reg_covar, double tol, int max_iter,
                              double* __restrict__ state) {
  if (threadIdx.x != 0) return;
  const int phase = (int)state[S_PHASE];
  if (phase == 2) return;
  const double eps10 = 10.0 * DBL_EPSILON;
  if (phase == 0) {
    // k-means centres; once they stop moving (or the budget is spent) the one-hot responsibilities of the
    // final assignment seed the mixture exactly as sklearn's _initialize does
    const double n0 = sums[0], n1 = sums[3];
    const double m0 = n0 > 0 ? sums[1] / n0 : state[S_MU], m1 = n1 > 0 ? sums[4] / n1 : state[S_MU + 1];
    const bool moved = (m0 != state[S_MU]) || (m1 != state[S_MU + 1]);
    const double left = state[S_KLEFT] - 1.0;
    state[S_KLEFT] = left;
    if (moved && left > 0) {
      state[S_MU] = m0; state[S_MU + 1] = m1;
      return;
    }
    const double nk0 = n0 + eps10, nk1 = n1 + eps10;
    const double mu0 = sums[1] / nk0, mu1 = sums[4] / nk1;
    state[S_MU] = mu0; state[S_MU + 1] = mu1;
    state[S_VAR] = fmax(sums[2] / nk0 - mu0 * mu0, 0.0) + reg_covar;
    state[S_VAR + 1] = fmax(sums[5] / nk1 - mu1 * mu1, 0.0) + reg_covar;
    state[S_W] = nk0 / n_total; state[S_W + 1] = nk1 / n_total;
    state[S_LB] = -INFINITY;
    state[S_ITER] = 0.0;
    state[S_PHASE] = 1.0;
    return;
  }
  const double lb = sums[6] / n_total;
  const double nk0 = sums[0] + eps10, nk1 = sums[3] + eps10;
  const double mu0 = sums[1] / nk0, mu1 = sums[4] / nk1;
  state[S_MU] = mu0; state[S_MU + 1] = mu1;
  state[S_VAR] = fmax(sums[2] / nk0 - mu0 * mu0, 0.0) + reg_covar;
  state[S_VAR + 1] = fmax(sums[5] / nk1 - mu1 * mu1, 0.0) + reg_covar;
  state[S_W] = nk0 / n_total; state[S_W + 1] = nk1 / n_total;
  const double change = lb - state[S_LB];
  state[S_LB] = lb;
  const double it = state[S_ITER] + 1.0;
  state[S_ITER] = it;
  if (fabs(change) < tol) { state[S_CONV] = 1.0; state[S_PHASE] = 2.0; }
  else if (it >= (double)max_iter) state[S_PHASE] = 2.0;
}

__global__ void init_kernel(double* state, const float* q2, int kmeans_iters) {
  if (threadIdx.x != 0) return;
  for (int i = 0; i < S_WORDS; ++i) state[i] = 0.0;
  state[S_MU] = (double)q2[0];
  state[S_MU + 1] = (double)q2[1];
  state[S_KLEFT] = (double)kmeans_iters;
}

}  // namespace gmm
}  // namespace sg

extern "C" {

size_t sg_gmm1d_workspace_bytes(void) {
  // state[16] | sums[8] | ticket (8 B) | per-CTA partials
  return (16 + 8 + 1) * 8 + (size_t)4 * 160 * 8 * 8 + 256;
}

int sg_gmm1d_begin(const float* init_centers2, int kmeans_iters, void* workspace, void* stream) {
  SG_READY();
  SG_REQUIRE(init_centers2 && workspace && kmeans_iters >= 1, "arguments");
  double* state = static_cast<double*>(workspace);
  SG_CUDA(cudaMemsetAsync(state + 24, 0, 8, sg::as_stream(stream)));
  sg::gmm::init_kernel<<<1, 32, 0, sg::as_stream(stream)>>>(state, init_centers2, kmeans_iters);
  SG_LAUNCH_CHECK();
  return SG_OK;
}

int sg_gmm1d_accumulate(const float* v, int64_t n, void* workspace, void* stream) {
  using namespace sg::gmm;
  SG_READY();
  SG_REQUIRE(workspace && n >= 0 && (n == 0 || v), "arguments");
  double* state = static_cast<double*>(workspace);
  double* sums = state + 16;
  unsigned int* ticket = reinterpret_cast<unsigned int*>(state + 24);
  double* partials = state + 25;
  int64_t g = sg::ceil_div(n > 0 ? n : 1, (int64_t)kThreads * 8);
  const int64_t cap = (int64_t)sg::state().sm_count * 4;
  if (g > cap) g = cap;
  if (g > 640) g = 640;
  accumulate_kernel<<<(unsigned)g, kThreads, 0, sg::as_stream(stream)>>>(v, n, state, partials, sums, ticket);
  SG_LAUNCH_CHECK();
  return SG_OK;
}

int sg_gmm1d_update(int64_t n_total, double reg_covar, double tol, int max_iter, void* workspace, void* stream) {
  SG_READY();
  SG_REQUIRE(workspace && n_total >= 1 && max_iter >= 1, "arguments");
  double* state = static_cast<double*>(workspace);
  sg::gmm::update_kernel<<<1, 32, 0, sg::as_stream(stream)>>>(state + 16, (double)n_total, reg_covar, tol, max_iter, state);
  SG_LAUNCH_CHECK();
  return SG_OK;
}

}  // extern "C"
