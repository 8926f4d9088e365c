// Window / grid attention pieces (maxvit.py:170-219, 289-341).
//
//  attn_gather : window (block) or dilated (grid) partition folded into the load addressing + register-token
//                concat + LayerNorm (no affine) + FiLM  ->  token matrix (Nw*S, C)          [HBM-bound]
//  attn_core   : per (window, head): QK-RMSNorm, QK^T + relative-position bias (index computed arithmetically,
//                maxvit.py:160-167), softmax, PV                                               [v1: SIMT fp32 math]
// The QKV and output projections are shifted-row GEMMs (vg_gemm.cu); the output projection's epilogue adds the
// residual and scatters through the inverse partition map, so no partitioned tensor is ever materialised in
// NCHW/NHWC form.
#include "vg_common.cuh"
#include "vg_host.h"
#include "vg_rng.cuh"

namespace vg {

// token row r = wdx*S + tok,  wdx = n*nwin + x*Y + y  (windows field-major, x-major: maxvit.py:306-307); the pixel of a
// window token is attn_token_pixel (vg_host.h)

// test hook: the global pixel index every token row gathers from / scatters to (-1 for register-token rows)
__global__ void attn_partition_debug_kernel(const AttnGeom g, long long* __restrict__ out, long long rows) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const int S = g.S(), nwin = g.nwin();
  const long long wdx = r / S;
  const int tok = (int)(r - wdx * S);
  const int n = (int)(wdx / nwin), wi = (int)(wdx - (long long)n * nwin);
  out[r] = tok < g.R ? -1 : (long long)n * g.Hl * g.Wl + attn_token_pixel(g, wi, tok - g.R);
}

template <typename T, typename TO>
__global__ void __launch_bounds__(256) attn_gather_kernel(const T* __restrict__ x, const float* __restrict__ reg, int reg_per_field,
                                                          const float* __restrict__ film, const AttnGeom g, float eps,
                                                          TO* __restrict__ tokens, long long rows) {
  const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const int S = g.S(), nwin = g.nwin(), C = g.C;
  const long long wdx = r / S;
  const int tok = (int)(r - wdx * S);
  const int n = (int)(wdx / nwin), wi = (int)(wdx - (long long)n * nwin);
  // C = 128*k: lane owns channels [lane*4 + 128*i, +4)
  float v[16];
  const int nv = C / 128;
  float s = 0.f;
  if (tok < g.R) {
    const float* src = reg + (reg_per_field ? (long long)n * g.R * C : 0) + (long long)tok * C;
    for (int i = 0; i < nv; ++i) {
      const float4 f = *reinterpret_cast<const float4*>(src + i * 128 + lane * 4);
      v[4 * i] = f.x; v[4 * i + 1] = f.y; v[4 * i + 2] = f.z; v[4 * i + 3] = f.w;
    }
  } else {
    const T* src = x + ((long long)n * g.Hl * g.Wl + attn_token_pixel(g, wi, tok - g.R)) * C;
    for (int i = 0; i < nv; ++i) {                                       // one 16- / 8-byte load per lane and 128 channels
      if constexpr (sizeof(T) == 4) {
        const float4 f = *reinterpret_cast<const float4*>(src + i * 128 + lane * 4);
        v[4 * i] = f.x; v[4 * i + 1] = f.y; v[4 * i + 2] = f.z; v[4 * i + 3] = f.w;
      } else {
        const uint2 u = *reinterpret_cast<const uint2*>(src + i * 128 + lane * 4);
        const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x)), b2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
        v[4 * i] = a.x; v[4 * i + 1] = a.y; v[4 * i + 2] = b2.x; v[4 * i + 3] = b2.y;
      }
    }
  }
  for (int i = 0; i < 4 * nv; ++i) s += v[i];
  const float mean = warp_sum(s) / (float)C;
  float ss = 0.f;
  for (int i = 0; i < 4 * nv; ++i) { v[i] -= mean; ss += v[i] * v[i]; }
  const float rstd = rsqrtf(warp_sum(ss) / (float)C + eps);             // nn.LayerNorm, no affine (maxvit.py:137)
  const float* gam = film + (long long)n * 2 * C;                        // [gamma | beta], used raw (maxvit.py:187)
  TO* dst = tokens + r * C;
  for (int i = 0; i < nv; ++i) {
    const int c = i * 128 + lane * 4;
    const float4 ga = *reinterpret_cast<const float4*>(gam + c), be = *reinterpret_cast<const float4*>(gam + C + c);
    const float o0 = v[4 * i] * rstd * ga.x + be.x, o1 = v[4 * i + 1] * rstd * ga.y + be.y;
    const float o2 = v[4 * i + 2] * rstd * ga.z + be.z, o3 = v[4 * i + 3] * rstd * ga.w + be.w;
    if constexpr (sizeof(TO) == 4) {
      *reinterpret_cast<float4*>(dst + c) = make_float4(o0, o1, o2, o3);
    } else {
      uint2 u;
      *reinterpret_cast<__nv_bfloat162*>(&u.x) = __floats2bfloat162_rn(o0, o1);
      *reinterpret_cast<__nv_bfloat162*>(&u.y) = __floats2bfloat162_rn(o2, o3);
      *reinterpret_cast<uint2*>(dst + c) = u;
    }
  }
}

// One warp per (window, head).  K-hat and V staged in shared memory as fp32, each lane owns query rows
// lane and lane+32; single pass over the keys with an online softmax.
template <typename T, int DH>
__global__ void __launch_bounds__(128) attn_core_kernel(const T* __restrict__ qkv, const float* __restrict__ qgamma,
                                                        const float* __restrict__ kgamma, const float* __restrict__ bias_table,
                                                        const AttnGeom g, int heads, T* __restrict__ out, long long pairs, const DropCfg drop) {
  extern __shared__ float sm[];
  const int S = g.S(), nb = (2 * g.win - 1) * (2 * g.win - 1) + 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* sk = sm + warp * (2 * S * (DH + 1) + nb);
  float* sv = sk + S * (DH + 1);
  float* sbias = sv + S * (DH + 1);
  const long long pair = (long long)blockIdx.x * 4 + warp;
  if (pair >= pairs) return;
  const long long wdx = pair / heads;
  const int hd = (int)(pair - wdx * heads);
  const int inner = heads * DH;
  const T* base = qkv + wdx * S * 3 * inner + hd * DH;
  const float rs = sqrtf((float)DH);

  for (int i = lane; i < nb; i += 32) sbias[i] = bias_table[i * heads + hd];
  // K-hat, V -> smem (row j handled by lane j, j+32)
  for (int j = lane; j < S; j += 32) {
    const T* kp = base + (long long)j * 3 * inner + inner;
    const T* vp = kp + inner;
    float kk[DH], nrm = 0.f;
#pragma unroll
    for (int d = 0; d < DH; d += 8) ld8(kp + d, kk + d);
#pragma unroll
    for (int d = 0; d < DH; ++d) nrm += kk[d] * kk[d];
    const float inv = rs / fmaxf(sqrtf(nrm), 1e-12f);                  // F.normalize eps (maxvit.py:30)
#pragma unroll
    for (int d = 0; d < DH; ++d) sk[j * (DH + 1) + d] = kk[d] * inv * kgamma[hd * DH + d];
#pragma unroll
    for (int d = 0; d < DH; d += 8) ld8(vp + d, kk + d);
#pragma unroll
    for (int d = 0; d < DH; ++d) sv[j * (DH + 1) + d] = kk[d];
  }
  __syncwarp();

  const int W2 = 2 * g.win - 1;
  for (int i = lane; i < S; i += 32) {
    float q[DH], o[DH], nrm = 0.f;
    const T* qp = base + (long long)i * 3 * inner;
#pragma unroll
    for (int d = 0; d < DH; d += 8) ld8(qp + d, q + d);
#pragma unroll
    for (int d = 0; d < DH; ++d) nrm += q[d] * q[d];
    const float inv = rs / fmaxf(sqrtf(nrm), 1e-12f);
#pragma unroll
    for (int d = 0; d < DH; ++d) { q[d] *= inv * qgamma[hd * DH + d]; o[d] = 0.f; }
    const int ti = i - g.R, ai = ti / g.win, bi = ti - ai * g.win;
    float m = -INFINITY, l = 0.f;
    uint32_t hsh = 0u;
    for (int j = 0; j < S; ++j) {
      // nn.Dropout on the probabilities (maxvit.py:146): the same counter-based mask as the fused kernels (vg_rng.cuh)
      if (drop.thresh && (j & 3) == 0) hsh = drop_hash(drop.seed, drop_row(wdx, i), drop_group_prob(drop.salt, hd, j >> 2));
      const float mk = !drop.thresh ? 1.0f : ((int)((hsh >> (8 * (j & 3))) & 255u) >= drop.thresh ? drop.scale : 0.f);
      float sc = 0.f;
#pragma unroll
      for (int d = 0; d < DH; ++d) sc = fmaf(q[d], sk[j * (DH + 1) + d], sc);
      int bidx = nb - 1;                                                 // register row/col -> shared last entry
      if (i >= g.R && j >= g.R) {
        const int tj = j - g.R, aj = tj / g.win, bj = tj - aj * g.win;
        bidx = (ai - aj + g.win - 1) * W2 + (bi - bj + g.win - 1);
      }
      sc += sbias[bidx];
      const float mn = fmaxf(m, sc);
      const float corr = __expf(m - mn), p = __expf(sc - mn);
      l = l * corr + p;
#pragma unroll
      for (int d = 0; d < DH; ++d) o[d] = fmaf(p * mk, sv[j * (DH + 1) + d], o[d] * corr);      // the normaliser l stays un-dropped
      m = mn;
    }
    const float il = 1.0f / l;
#pragma unroll
    for (int d = 0; d < DH; ++d) o[d] *= il;
    T* op = out + (wdx * S + i) * inner + hd * DH;
#pragma unroll
    for (int d = 0; d < DH; d += 8) st8(op + d, o + d);
  }
}

// out_bf16: fp32 residual stream in, bf16 tokens out (the mixed-precision training backward re-materialises tokens in 16 bits)
int attn_gather_run(int dtype, const void* x, const float* reg, int reg_per_field, const float* film, const AttnGeom& g,
                    float eps, void* tokens, int out_bf16, cudaStream_t st) {
  if (g.C % 128 || g.C > 512) return set_error("attn_gather: C=%d must be a multiple of 128 (<=512)", g.C);
  const long long rows = (long long)g.N * g.nwin() * g.S();
  const unsigned grid = (unsigned)((rows + 7) / 8);
  if (dtype == 0) attn_gather_kernel<bf16, bf16><<<grid, 256, 0, st>>>(reinterpret_cast<const bf16*>(x), reg, reg_per_field, film, g, eps, reinterpret_cast<bf16*>(tokens), rows);
  else if (out_bf16) attn_gather_kernel<float, bf16><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(x), reg, reg_per_field, film, g, eps, reinterpret_cast<bf16*>(tokens), rows);
  else attn_gather_kernel<float, float><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(x), reg, reg_per_field, film, g, eps, reinterpret_cast<float*>(tokens), rows);
  return check_launch("attn_gather_kernel");
}

int attn_partition_debug_run(const AttnGeom& g, long long* out, cudaStream_t st) {
  const long long rows = (long long)g.N * g.nwin() * g.S();
  attn_partition_debug_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(g, out, rows);
  return check_launch("attn_partition_debug_kernel");
}

template <typename T, int DH>
static int core_launch(const void* qkv, const float* qg, const float* kg, const float* bt, const AttnGeom& g, int heads, void* out, const DropCfg& drop, cudaStream_t st) {
  const int S = g.S(), nb = (2 * g.win - 1) * (2 * g.win - 1) + 1;
  const size_t smem = 4 * (size_t)(2 * S * (DH + 1) + nb) * sizeof(float);
  static PerDeviceSize attr_pd;                             // the size depends on the window geometry: raise the limit when it grows
  size_t& attr_bytes = attr_pd.cur();
  if (smem > attr_bytes) {
    cudaError_t e = cudaFuncSetAttribute(attn_core_kernel<T, DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error("attn_core smem attr: %s", cudaGetErrorString(e));
    attr_bytes = smem;
  }
  const long long pairs = (long long)g.N * g.nwin() * heads;
  attn_core_kernel<T, DH><<<(unsigned)((pairs + 3) / 4), 128, smem, st>>>(reinterpret_cast<const T*>(qkv), qg, kg, bt, g, heads, reinterpret_cast<T*>(out), pairs, drop);
  return check_launch("attn_core_kernel");
}

int attn_core_run(int dtype, const void* qkv, const float* qgamma, const float* kgamma, const float* bias_table,
                  const AttnGeom& g, int heads, int dh, void* out, unsigned seed, unsigned salt, int drop_thresh, cudaStream_t st) {
  if (g.S() > 128) return set_error("attn_core: sequence %d too long", g.S());
  if (drop_thresh < 0 || drop_thresh > 255 || (drop_thresh && g.S() > 64)) return set_error("attn_core: bad dropout threshold %d (or sequence > 64)", drop_thresh);
  DropCfg drop;
  drop.seed = seed; drop.salt = salt; drop.thresh = drop_thresh; drop.scale = 256.0f / (256.0f - (float)drop_thresh);
  if (dh == 32) return dtype == 0 ? core_launch<bf16, 32>(qkv, qgamma, kgamma, bias_table, g, heads, out, drop, st)
                                  : core_launch<float, 32>(qkv, qgamma, kgamma, bias_table, g, heads, out, drop, st);
  if (dh == 64) return dtype == 0 ? core_launch<bf16, 64>(qkv, qgamma, kgamma, bias_table, g, heads, out, drop, st)
                                  : core_launch<float, 64>(qkv, qgamma, kgamma, bias_table, g, heads, out, drop, st);
  return set_error("attn_core: dim_head %d not supported (32 or 64)", dh);
}

}  // namespace vg
