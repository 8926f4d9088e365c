// Fused evaluation statistics (/root/reference/src/evaluation_vit.py:239-455).
//
// The reference's test loop turns every batch of predictions into ~200 host-synchronising reductions (`.item()` after
// each masked sum: 4 x 16 confusion cells, 12 skill counters, 4 x 3 x L per-lead TP/TN/FP/FN, RMSE / MAE sums per class
// threshold and lead, absolute / squared / relative error sums), for the model and for three baselines (persistence,
// the 21 h CMAQ run, the mean of the four CMAQ runs).  Everything in that list is a linear function of
//
//     counts[m][j][a][t]   #{ class(v_m) = a  and  truth class = t }          m: method, j: lead, a: 0..3, t: -1..3
//     sums  [m][j][t][k]   sum of |v_m - y| (k = 0) and (v_m - y)^2 (k = 1) over truth class t
//     glob                 sums over everything: (v_m - y)/y and |(v_m - y)/y| over y > 0; v_m, v_m^2, v_m*y; y, y^2
//                          (normalised bias / error and the Pearson correlation, for which the reference keeps every
//                          value of the whole evaluation in Python lists, :328-332, :507-523, :552-576)
//
// so one pass over the batch (24 bytes per grid cell and lead) accumulates those tables on the device and the host
// derives the reference's quantities once, at the end of the evaluation.  HBM-bound byte / integer work: coalesced
// loads along the cell index, class counts through shared-memory integer atomics, error sums in registers -> warp
// shuffles -> fixed-order block partials -> a one-block finalize, so the floating-point sums do not depend on the
// launch's scheduling.  Integer tables are exact.
#include "vg_host.h"

namespace vg {
namespace {

constexpr int EV_THREADS = 256;
constexpr int EV_METHODS = 4;           // model, persistence, sim 21h, sim average
constexpr int EV_TCLS = 5;              // truth class -1, 0, 1, 2, 3
constexpr int EV_NSUM = EV_METHODS * EV_TCLS * 2;      // 40 per lead
constexpr int EV_NGLOB = EV_METHODS * 2 + EV_METHODS * 3 + 2;   // 22: norm[m][2] | mom[m][3] = v, v^2, v*y | y, y^2
constexpr int EV_PART = EV_NSUM + EV_NGLOB;            // doubles per block partial

// evaluation_vit.py:31-32 with range_4class = [(-1,15],(15,35],(35,75],(75,inf)) (:194): default 0, so everything that is
// not above the first boundary (including values <= -1 and NaN) is class 0
__device__ __forceinline__ int pm_class(float v, float b1, float b2, float b3) {
  return v > b3 ? 3 : (v > b2 ? 2 : (v > b1 ? 1 : 0));
}

template <typename CT>
__global__ void __launch_bounds__(EV_THREADS)
eval_metrics_kernel(float* __restrict__ preds, const float* __restrict__ truth, const CT* __restrict__ tcls,
                    const float* __restrict__ persist, const float* __restrict__ sim21, const float* __restrict__ simavg,
                    int B, int L, int P, float b1, float b2, float b3, int clamp_preds,
                    unsigned long long* __restrict__ counts, unsigned long long* __restrict__ nonzero,
                    double* __restrict__ partial) {
  __shared__ unsigned int s_counts[EV_METHODS * 4 * EV_TCLS];
  __shared__ float s_red[EV_THREADS / 32][EV_PART];
  __shared__ unsigned int s_nz;
  const int j = blockIdx.y;
  for (int i = threadIdx.x; i < EV_METHODS * 4 * EV_TCLS; i += EV_THREADS) s_counts[i] = 0u;
  if (threadIdx.x == 0) s_nz = 0u;
  __syncthreads();
  float sabs[EV_METHODS][EV_TCLS], ssq[EV_METHODS][EV_TCLS], nrm[EV_METHODS][2], mom[EV_METHODS][3], ys[2] = {0.f, 0.f};
#pragma unroll
  for (int m = 0; m < EV_METHODS; ++m) {
    nrm[m][0] = nrm[m][1] = 0.f;
    mom[m][0] = mom[m][1] = mom[m][2] = 0.f;
#pragma unroll
    for (int t = 0; t < EV_TCLS; ++t) sabs[m][t] = ssq[m][t] = 0.f;
  }
  unsigned int nz = 0u;
  const long long cells = (long long)B * P;
  for (long long i = (long long)blockIdx.x * EV_THREADS + threadIdx.x; i < cells; i += (long long)gridDim.x * EV_THREADS) {
    const int b = (int)(i / P);
    const int p = (int)(i - (long long)b * P);
    const long long e = ((long long)b * L + j) * P + p;
    float v[EV_METHODS];
    v[0] = preds[e];
    if (clamp_preds && v[0] < 0.f) { v[0] = 0.f; preds[e] = 0.f; }       // preds[preds < 0.] = 0.   (:254)
    v[1] = persist[i];                                                  // last_PM repeated over the leads (:241-243)
    v[2] = sim21[e];
    v[3] = simavg[e];
    const float y = truth[e];
    int t = (int)tcls[e];
    t = t < -1 ? -1 : (t > 3 ? 3 : t);
    const int ti = t + 1;
    const bool pos = y > 0.f;                                           // nonzero_mask (:311)
    nz += pos ? 1u : 0u;
    const float inv_y = pos ? 1.0f / y : 0.f;
    ys[0] += y; ys[1] = fmaf(y, y, ys[1]);
#pragma unroll
    for (int m = 0; m < EV_METHODS; ++m) {
      mom[m][0] += v[m]; mom[m][1] = fmaf(v[m], v[m], mom[m][1]); mom[m][2] = fmaf(v[m], y, mom[m][2]);
      const float d = v[m] - y;
      const float ad = fabsf(d), sq = d * d;
      const int a = pm_class(v[m], b1, b2, b3);
      atomicAdd(&s_counts[(m * 4 + a) * EV_TCLS + ti], 1u);
#pragma unroll
      for (int tt = 0; tt < EV_TCLS; ++tt) {
        sabs[m][tt] += tt == ti ? ad : 0.f;
        ssq[m][tt] += tt == ti ? sq : 0.f;
      }
      if (pos) { const float q = d * inv_y; nrm[m][0] += q; nrm[m][1] += fabsf(q); }
    }
  }
  // block reduction: shuffles inside a warp, then warp rows summed in a fixed order
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int m = 0; m < EV_METHODS; ++m) {
#pragma unroll
    for (int t = 0; t < EV_TCLS; ++t) {
      float a = sabs[m][t], q = ssq[m][t];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
      if (lane == 0) { s_red[warp][(m * EV_TCLS + t) * 2] = a; s_red[warp][(m * EV_TCLS + t) * 2 + 1] = q; }
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      float a = nrm[m][k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (lane == 0) s_red[warp][EV_NSUM + m * 2 + k] = a;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float a = mom[m][k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (lane == 0) s_red[warp][EV_NSUM + EV_METHODS * 2 + m * 3 + k] = a;
    }
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    float a = ys[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) s_red[warp][EV_NSUM + EV_METHODS * 5 + k] = a;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nz += __shfl_xor_sync(0xffffffffu, nz, o);
  if (lane == 0 && nz) atomicAdd(&s_nz, nz);
  __syncthreads();
  if (threadIdx.x < EV_PART) {
    double a = 0.0;
#pragma unroll
    for (int w = 0; w < EV_THREADS / 32; ++w) a += (double)s_red[w][threadIdx.x];
    partial[((long long)j * gridDim.x + blockIdx.x) * EV_PART + threadIdx.x] = a;
  }
  for (int i = threadIdx.x; i < EV_METHODS * 4 * EV_TCLS; i += EV_THREADS) {
    const unsigned int c = s_counts[i];
    if (c) {
      const int m = i / (4 * EV_TCLS), r = i - m * 4 * EV_TCLS;
      atomicAdd(&counts[((long long)m * L + j) * 4 * EV_TCLS + r], (unsigned long long)c);      // integer: order-independent
    }
  }
  if (threadIdx.x == 0 && s_nz) atomicAdd(nonzero, (unsigned long long)s_nz);
}

// one block: folds the block partials into the running tables in a fixed order.  sums: [4][L][5][2]; glob: [22];
// loss_sum += (sum of squared model errors of THIS call) / (B*L*P)   (criterion = MSELoss, :140, :291)
__global__ void __launch_bounds__(256)
eval_finalize_kernel(const double* __restrict__ partial, int L, int nblk, double inv_numel, double* __restrict__ sums,
                     double* __restrict__ glob, double* __restrict__ loss_sum) {
  __shared__ double s_sq[256];
  double my_sq = 0.0;
  for (int it = threadIdx.x; it < L * EV_NSUM; it += blockDim.x) {
    const int j = it / EV_NSUM, r = it - j * EV_NSUM;
    double a = 0.0;
    for (int b = 0; b < nblk; ++b) a += partial[((long long)j * nblk + b) * EV_PART + r];
    const int m = r / (EV_TCLS * 2), q = r - m * EV_TCLS * 2;
    sums[((long long)m * L + j) * EV_TCLS * 2 + q] += a;
    if (m == 0 && (q & 1)) my_sq += a;
  }
  for (int r = threadIdx.x; r < EV_NGLOB; r += blockDim.x) {
    double a = 0.0;
    for (int j = 0; j < L; ++j)
      for (int b = 0; b < nblk; ++b) a += partial[((long long)j * nblk + b) * EV_PART + EV_NSUM + r];
    glob[r] += a;
  }
  s_sq[threadIdx.x] = my_sq;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
    for (int i = 0; i < (int)blockDim.x; ++i) a += s_sq[i];
    *loss_sum += a * inv_numel;
  }
}

int eval_blocks(long long cells) {
  long long nb = (cells + EV_THREADS * 4 - 1) / (EV_THREADS * 4);
  if (nb < 1) nb = 1;
  if (nb > 74) nb = 74;                     // x L leads: a multiple of half the 148 SMs per lead
  return (int)nb;
}

}  // namespace

long long eval_metrics_workspace_run(int B, int L, int P) {
  return (long long)L * eval_blocks((long long)B * P) * EV_PART;
}

int eval_metrics_run(float* preds, const float* truth, const void* tcls, int cls_i64, const float* persist, const float* sim21,
                     const float* simavg, int B, int L, int P, float b1, float b2, float b3, int clamp_preds,
                     unsigned long long* counts, double* sums, double* glob, unsigned long long* nonzero, double* loss_sum,
                     double* work, long long work_elems, cudaStream_t st) {
  if (B <= 0 || L <= 0 || P <= 0) return set_error("eval_metrics: empty batch (B=%d, L=%d, P=%d)", B, L, P);
  if (!(b1 <= b2 && b2 <= b3)) return set_error("eval_metrics: class boundaries must be ascending");
  const int nblk = eval_blocks((long long)B * P);
  if (work_elems < (long long)L * nblk * EV_PART) return set_error("eval_metrics: workspace too small (%lld < %lld doubles)", work_elems, (long long)L * nblk * EV_PART);
  dim3 grid((unsigned)nblk, (unsigned)L);
  if (cls_i64)
    eval_metrics_kernel<long long><<<grid, EV_THREADS, 0, st>>>(preds, truth, reinterpret_cast<const long long*>(tcls), persist, sim21, simavg,
                                                                 B, L, P, b1, b2, b3, clamp_preds, counts, nonzero, work);
  else
    eval_metrics_kernel<int><<<grid, EV_THREADS, 0, st>>>(preds, truth, reinterpret_cast<const int*>(tcls), persist, sim21, simavg,
                                                           B, L, P, b1, b2, b3, clamp_preds, counts, nonzero, work);
  int rc = check_launch("eval_metrics_kernel");
  if (rc) return rc;
  eval_finalize_kernel<<<1, 256, 0, st>>>(work, L, nblk, 1.0 / ((double)B * L * P), sums, glob, loss_sum);
  return check_launch("eval_finalize_kernel");
}

}  // namespace vg
