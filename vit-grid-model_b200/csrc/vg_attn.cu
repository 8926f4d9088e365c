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

__device__ __forceinline__ float4 ld4f(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld4f(const bf16* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x)), b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}

// One block per (window, head), one query row per thread (ceil(S / 32) warps).  K-hat and V are staged in shared memory as fp32
// rows of DH floats (pad rows up to a multiple of 4 are zeros): DH/4 lanes stage one row with 16-byte loads and stores, the
// row norm by shuffles.  Every lane of a warp then reads the same key, so the reads are 16-byte broadcasts.  Keys go four at a
// time through the online softmax: four scores (each four partial sums -- a single 64-long FMA chain was the latency bound of
// the first version of this kernel, which also kept one (K, V) copy per warp: 113 KB per block, 8 warps per SM), one rescale
// of the output row per four keys.  BASELINE configs[4] (32 heads x 64, 360 windows): 0.93 -> see profiles/r02_summary.md.
template <typename T, int DH>
__global__ void __launch_bounds__(128, DH == 64 ? 3 : 5) attn_core_kernel(const T* __restrict__ qkv, const float* __restrict__ qgamma,
                                                        const float* __restrict__ kgamma, const float* __restrict__ bias_table,
                                                        const AttnGeom g, int heads, T* __restrict__ out, const DropCfg drop) {
  extern __shared__ __align__(16) float sm[];
  const int S = g.S(), Sp = (S + 3) & ~3, nb = (2 * g.win - 1) * (2 * g.win - 1) + 1;
  const int lane = threadIdx.x & 31;
  float* sk = sm;
  float* sv = sk + Sp * DH;
  float* sbias = sv + Sp * DH;
  const long long pair = blockIdx.x;
  const long long wdx = pair / heads;
  const int hd = (int)(pair - wdx * heads);
  const int inner = heads * DH;
  const T* base = qkv + wdx * S * 3 * inner + hd * DH;
  const float rs = sqrtf((float)DH);

  for (int i = threadIdx.x; i < nb; i += blockDim.x) sbias[i] = bias_table[i * heads + hd];
  {
    constexpr int LPR = DH / 4, RPP = 32 / LPR;              // lanes per row, rows per warp pass
    const int sub = lane / LPR, ch = lane - sub * LPR;
    const float4 kg = *reinterpret_cast<const float4*>(kgamma + hd * DH + ch * 4);
    // all the loads of a batch of UNR passes are issued before the first norm: one global round trip per batch, not per pass
    constexpr int UNR = 7;
    const int step = (blockDim.x >> 5) * RPP;
    for (int jb = (threadIdx.x >> 5) * RPP; jb < Sp; jb += step * UNR) {
      float4 k[UNR], v[UNR];
#pragma unroll
      for (int t = 0; t < UNR; ++t) {
        const int j = jb + t * step + sub;
        k[t] = make_float4(0.f, 0.f, 0.f, 0.f); v[t] = k[t];
        if (j < S) { const T* kp = base + (long long)j * 3 * inner + inner + ch * 4; k[t] = ld4f(kp); v[t] = ld4f(kp + inner); }
      }
#pragma unroll
      for (int t = 0; t < UNR; ++t) {
        if (jb + t * step < Sp) {                                // warp-uniform; then j < Sp too (Sp is a multiple of 4, RPP divides 4)
          const int j = jb + t * step + sub;
          float nrm = k[t].x * k[t].x + k[t].y * k[t].y + k[t].z * k[t].z + k[t].w * k[t].w;
#pragma unroll
          for (int o = LPR / 2; o; o >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
          const float inv = rs / fmaxf(sqrtf(nrm), 1e-12f);      // F.normalize eps (maxvit.py:30)
          *reinterpret_cast<float4*>(sk + j * DH + ch * 4) = make_float4(k[t].x * inv * kg.x, k[t].y * inv * kg.y, k[t].z * inv * kg.z, k[t].w * inv * kg.w);
          *reinterpret_cast<float4*>(sv + j * DH + ch * 4) = v[t];
        }
      }
    }
  }
  __syncthreads();

  const int i = threadIdx.x;
  if (i >= S) return;
  const int W2 = 2 * g.win - 1;
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
  for (int j0 = 0; j0 < Sp; j0 += 4) {
    // nn.Dropout on the probabilities (maxvit.py:146): the same counter-based mask as the fused kernels (vg_rng.cuh)
    const uint32_t hsh = drop.thresh ? drop_hash(drop.seed, drop_row(wdx, i), drop_group_prob(drop.salt, hd, j0 >> 2)) : 0u;
    float sc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + u;
      const float4* kr = reinterpret_cast<const float4*>(sk + j * DH);
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
      for (int c = 0; c < DH / 4; ++c) {
        const float4 k = kr[c];
        s0 = fmaf(q[4 * c], k.x, s0); s1 = fmaf(q[4 * c + 1], k.y, s1); s2 = fmaf(q[4 * c + 2], k.z, s2); s3 = fmaf(q[4 * c + 3], k.w, s3);
      }
      int bidx = nb - 1;                                                 // register row/col -> shared last entry
      if (i >= g.R && j >= g.R && j < S) {
        const int tj = j - g.R, aj = tj / g.win, bj = tj - aj * g.win;
        bidx = (ai - aj + g.win - 1) * W2 + (bi - bj + g.win - 1);
      }
      sc[u] = j < S ? (s0 + s1) + (s2 + s3) + sbias[bidx] : -INFINITY;
    }
    const float mn = fmaxf(fmaxf(m, fmaxf(sc[0], sc[1])), fmaxf(sc[2], sc[3]));    // the first key of a group is always valid: finite
    const float corr = __expf(m - mn);
    float pm[4], ps = 0.f;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float p = __expf(sc[u] - mn);
      ps += p;                                                             // the normaliser l stays un-dropped
      pm[u] = !drop.thresh ? p : ((int)((hsh >> (8 * u)) & 255u) >= drop.thresh ? p * drop.scale : 0.f);
    }
    l = l * corr + ps;
    m = mn;
    const float4* v0 = reinterpret_cast<const float4*>(sv + j0 * DH);
#pragma unroll
    for (int c = 0; c < DH / 4; ++c) {
      const float4 a = v0[c], b = v0[c + DH / 4], cc = v0[c + 2 * (DH / 4)], dd = v0[c + 3 * (DH / 4)];
      o[4 * c]     = fmaf(pm[0], a.x, fmaf(pm[1], b.x, fmaf(pm[2], cc.x, fmaf(pm[3], dd.x, o[4 * c] * corr))));
      o[4 * c + 1] = fmaf(pm[0], a.y, fmaf(pm[1], b.y, fmaf(pm[2], cc.y, fmaf(pm[3], dd.y, o[4 * c + 1] * corr))));
      o[4 * c + 2] = fmaf(pm[0], a.z, fmaf(pm[1], b.z, fmaf(pm[2], cc.z, fmaf(pm[3], dd.z, o[4 * c + 2] * corr))));
      o[4 * c + 3] = fmaf(pm[0], a.w, fmaf(pm[1], b.w, fmaf(pm[2], cc.w, fmaf(pm[3], dd.w, o[4 * c + 3] * corr))));
    }
  }
  const float il = 1.0f / l;
#pragma unroll
  for (int d = 0; d < DH; ++d) o[d] *= il;
  T* op = out + (wdx * S + i) * inner + hd * DH;
#pragma unroll
  for (int d = 0; d < DH; d += 8) st8(op + d, o + d);
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
  const int S = g.S(), Sp = (S + 3) & ~3, nb = (2 * g.win - 1) * (2 * g.win - 1) + 1;
  const size_t smem = (size_t)(2 * Sp * DH + nb) * sizeof(float);
  static PerDeviceSize attr_pd;                             // the size depends on the window geometry: raise the limit when it grows
  size_t& attr_bytes = attr_pd.cur();
  if (smem > attr_bytes) {
    cudaError_t e = cudaFuncSetAttribute(attn_core_kernel<T, DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_core_kernel<T, DH>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return set_error("attn_core smem attr: %s", cudaGetErrorString(e));
    attr_bytes = smem;
  }
  const long long pairs = (long long)g.N * g.nwin() * heads;
  attn_core_kernel<T, DH><<<(unsigned)pairs, 32 * ((S + 31) / 32), smem, st>>>(reinterpret_cast<const T*>(qkv), qg, kg, bt, g, heads, reinterpret_cast<T*>(out), drop);
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
