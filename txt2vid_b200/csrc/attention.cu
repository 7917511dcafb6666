// Fused core of the SA-GAN non-local block (txt2vid/models/layers.py:23-36 2-D, :52-68 3-D):
//     phi_p = maxpool_{1x2x2}(phi),  g_p = maxpool_{1x2x2}(g)
//     beta  = softmax_k(theta . phi_p^T)            (P queries x Kp keys per map, Kp = P/4)
//     o     = beta . g_p
// in ONE kernel per direction: the (P x Kp) attention matrix (1 MB fp32 per 32x32 map, 1 GB per step at batch
// 1024) never exists in memory.  The 1x1 convolutions around it (theta, phi, g, o) run on the tcgen05 engine.
// HBM-bound: algorithmic bytes = theta + phi + g + o rows (bf16, padded channels).
//
// Forward : thread = one query; pooled keys/values of the map live in shared memory (fp32).
// Backward: phase A (thread = query): softmax statistics, D_q = sum_j beta_j dbeta_j, d theta;
//           phase B (thread = key): d phi_p, d g_p summed over the queries, routed to the arg-max voxel of each
//           2x2 pooling window (the other three voxels get zeros).
// First-order only: the discriminator's block sits on the gradient-penalty path (double backward) and keeps the
// composite formulation of ops.nonlocal_block.
#include "t2v_common.cuh"

namespace t2v {

static constexpr int kMaxC8 = 8, kMaxC2 = 16;

struct AttnParams {
  int N, D, H, W, c8, c2, C8p, C2p;   // real and padded (memory) channel counts
  int P, Kp, Hp, Wp;
  const __nv_bfloat16* theta; const __nv_bfloat16* phi; const __nv_bfloat16* g;
  __nv_bfloat16* o;
  const __nv_bfloat16* dout;
  __nv_bfloat16* dtheta; __nv_bfloat16* dphi; __nv_bfloat16* dg;
};

// pooled key k of map n: window voxels and the max / arg-max over them, channel c
__device__ __forceinline__ void pooled_window(const AttnParams& p, int k, int* pos4) {
  const int wp = k % p.Wp, hp = (k / p.Wp) % p.Hp, d = k / (p.Wp * p.Hp);
  const int base = (d * p.H + 2 * hp) * p.W + 2 * wp;
  pos4[0] = base; pos4[1] = base + 1; pos4[2] = base + p.W; pos4[3] = base + p.W + 1;
}

__device__ __forceinline__ void build_keys(const AttnParams& p, long long map0, float* s_phi, float* s_g,
                                           unsigned char* s_aphi, unsigned char* s_ag) {
  for (int k = threadIdx.x; k < p.Kp; k += blockDim.x) {
    int pos4[4];
    pooled_window(p, k, pos4);
    for (int c = 0; c < p.c8; ++c) {
      float best = -INFINITY; int arg = 0;
      for (int j = 0; j < 4; ++j) {
        const float v = bf2f(p.phi[(map0 + pos4[j]) * p.C8p + c]);
        if (v > best) { best = v; arg = j; }
      }
      s_phi[k * p.c8 + c] = best;
      if (s_aphi) s_aphi[k * p.c8 + c] = (unsigned char)arg;
    }
    for (int c = 0; c < p.c2; ++c) {
      float best = -INFINITY; int arg = 0;
      for (int j = 0; j < 4; ++j) {
        const float v = bf2f(p.g[(map0 + pos4[j]) * p.C2p + c]);
        if (v > best) { best = v; arg = j; }
      }
      s_g[k * p.c2 + c] = best;
      if (s_ag) s_ag[k * p.c2 + c] = (unsigned char)arg;
    }
  }
}

__global__ void __launch_bounds__(256) attn_fwd_kernel(const AttnParams p) {
  extern __shared__ float smem_f[];
  float* s_phi = smem_f;                       // [Kp][c8]
  float* s_g = s_phi + p.Kp * p.c8;            // [Kp][c2]
  const long long map0 = (long long)blockIdx.x * p.P;
  build_keys(p, map0, s_phi, s_g, nullptr, nullptr);
  __syncthreads();
  const int q = blockIdx.y * blockDim.x + threadIdx.x;
  if (q >= p.P) return;
  float th[kMaxC8];
#pragma unroll
  for (int c = 0; c < kMaxC8; ++c) th[c] = c < p.c8 ? bf2f(p.theta[(map0 + q) * p.C8p + c]) : 0.f;
  float m = -INFINITY;
  for (int k = 0; k < p.Kp; ++k) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < kMaxC8; ++c) if (c < p.c8) s = fmaf(th[c], s_phi[k * p.c8 + c], s);
    m = fmaxf(m, s);
  }
  float l = 0.f, acc[kMaxC2];
#pragma unroll
  for (int c = 0; c < kMaxC2; ++c) acc[c] = 0.f;
  for (int k = 0; k < p.Kp; ++k) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < kMaxC8; ++c) if (c < p.c8) s = fmaf(th[c], s_phi[k * p.c8 + c], s);
    const float e = __expf(s - m);
    l += e;
#pragma unroll
    for (int c = 0; c < kMaxC2; ++c) if (c < p.c2) acc[c] = fmaf(e, s_g[k * p.c2 + c], acc[c]);
  }
  const float inv = 1.f / l;
  __nv_bfloat16* orow = p.o + (map0 + q) * p.C2p;
  for (int c = 0; c < p.C2p; ++c) orow[c] = f2bf(c < p.c2 ? acc[c < kMaxC2 ? c : 0] * inv : 0.f);
}

// one CTA per map
__global__ void __launch_bounds__(256) attn_bwd_kernel(const AttnParams p) {
  extern __shared__ float smem_f[];
  float* s_phi = smem_f;                              // [Kp][c8]
  float* s_g = s_phi + p.Kp * p.c8;                   // [Kp][c2]
  float* s_th = s_g + p.Kp * p.c2;                    // [P][c8]
  float* s_do = s_th + p.P * p.c8;                    // [P][c2]
  float* s_m = s_do + p.P * p.c2;                     // [P] row max
  float* s_il = s_m + p.P;                            // [P] 1 / row sum
  float* s_dq = s_il + p.P;                           // [P] D_q
  unsigned char* s_aphi = reinterpret_cast<unsigned char*>(s_dq + p.P);   // [Kp][c8]
  unsigned char* s_ag = s_aphi + p.Kp * p.c8;                              // [Kp][c2]
  const long long map0 = (long long)blockIdx.x * p.P;
  build_keys(p, map0, s_phi, s_g, s_aphi, s_ag);
  for (int i = threadIdx.x; i < p.P * p.c8; i += blockDim.x)
    s_th[i] = bf2f(p.theta[(map0 + i / p.c8) * p.C8p + i % p.c8]);
  for (int i = threadIdx.x; i < p.P * p.c2; i += blockDim.x)
    s_do[i] = bf2f(p.dout[(map0 + i / p.c2) * p.C2p + i % p.c2]);
  __syncthreads();
  // ---- phase A: per query
  for (int q = threadIdx.x; q < p.P; q += blockDim.x) {
    float th[kMaxC8], dy[kMaxC2];
#pragma unroll
    for (int c = 0; c < kMaxC8; ++c) th[c] = c < p.c8 ? s_th[q * p.c8 + c] : 0.f;
#pragma unroll
    for (int c = 0; c < kMaxC2; ++c) dy[c] = c < p.c2 ? s_do[q * p.c2 + c] : 0.f;
    float m = -INFINITY;
    for (int k = 0; k < p.Kp; ++k) {
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < kMaxC8; ++c) if (c < p.c8) s = fmaf(th[c], s_phi[k * p.c8 + c], s);
      m = fmaxf(m, s);
    }
    float l = 0.f, dq = 0.f;
    for (int k = 0; k < p.Kp; ++k) {
      float s = 0.f, db = 0.f;
#pragma unroll
      for (int c = 0; c < kMaxC8; ++c) if (c < p.c8) s = fmaf(th[c], s_phi[k * p.c8 + c], s);
#pragma unroll
      for (int c = 0; c < kMaxC2; ++c) if (c < p.c2) db = fmaf(dy[c], s_g[k * p.c2 + c], db);
      const float e = __expf(s - m);
      l += e;
      dq = fmaf(e, db, dq);
    }
    const float il = 1.f / l;
    dq *= il;
    float dth[kMaxC8];
#pragma unroll
    for (int c = 0; c < kMaxC8; ++c) dth[c] = 0.f;
    for (int k = 0; k < p.Kp; ++k) {
      float s = 0.f, db = 0.f;
#pragma unroll
      for (int c = 0; c < kMaxC8; ++c) if (c < p.c8) s = fmaf(th[c], s_phi[k * p.c8 + c], s);
#pragma unroll
      for (int c = 0; c < kMaxC2; ++c) if (c < p.c2) db = fmaf(dy[c], s_g[k * p.c2 + c], db);
      const float ds = __expf(s - m) * il * (db - dq);
#pragma unroll
      for (int c = 0; c < kMaxC8; ++c) if (c < p.c8) dth[c] = fmaf(ds, s_phi[k * p.c8 + c], dth[c]);
    }
    s_m[q] = m; s_il[q] = il; s_dq[q] = dq;
    __nv_bfloat16* drow = p.dtheta + (map0 + q) * p.C8p;
    for (int c = 0; c < p.C8p; ++c) drow[c] = f2bf(c < p.c8 ? dth[c < kMaxC8 ? c : 0] : 0.f);
  }
  __syncthreads();
  // ---- phase B: per key
  for (int k = threadIdx.x; k < p.Kp; k += blockDim.x) {
    float ph[kMaxC8], gk[kMaxC2], dph[kMaxC8], dgk[kMaxC2];
#pragma unroll
    for (int c = 0; c < kMaxC8; ++c) { ph[c] = c < p.c8 ? s_phi[k * p.c8 + c] : 0.f; dph[c] = 0.f; }
#pragma unroll
    for (int c = 0; c < kMaxC2; ++c) { gk[c] = c < p.c2 ? s_g[k * p.c2 + c] : 0.f; dgk[c] = 0.f; }
    for (int q = 0; q < p.P; ++q) {
      float s = 0.f, db = 0.f;
#pragma unroll
      for (int c = 0; c < kMaxC8; ++c) if (c < p.c8) s = fmaf(s_th[q * p.c8 + c], ph[c], s);
#pragma unroll
      for (int c = 0; c < kMaxC2; ++c) if (c < p.c2) db = fmaf(s_do[q * p.c2 + c], gk[c], db);
      const float beta = __expf(s - s_m[q]) * s_il[q];
      const float ds = beta * (db - s_dq[q]);
#pragma unroll
      for (int c = 0; c < kMaxC8; ++c) if (c < p.c8) dph[c] = fmaf(ds, s_th[q * p.c8 + c], dph[c]);
#pragma unroll
      for (int c = 0; c < kMaxC2; ++c) if (c < p.c2) dgk[c] = fmaf(beta, s_do[q * p.c2 + c], dgk[c]);
    }
    int pos4[4];
    pooled_window(p, k, pos4);
    for (int j = 0; j < 4; ++j) {
      __nv_bfloat16* r1 = p.dphi + (map0 + pos4[j]) * p.C8p;
      for (int c = 0; c < p.C8p; ++c)
        r1[c] = f2bf((c < p.c8 && s_aphi[k * p.c8 + c] == j) ? dph[c < kMaxC8 ? c : 0] : 0.f);
      __nv_bfloat16* r2 = p.dg + (map0 + pos4[j]) * p.C2p;
      for (int c = 0; c < p.C2p; ++c)
        r2[c] = f2bf((c < p.c2 && s_ag[k * p.c2 + c] == j) ? dgk[c < kMaxC2 ? c : 0] : 0.f);
    }
  }
}

static int attn_fill(AttnParams& p, int64_t N, int D, int H, int W, int c8, int c2, int C8p, int C2p) {
  if (c8 < 1 || c8 > kMaxC8 || c2 < 1 || c2 > kMaxC2 || C8p < c8 || C2p < c2) return T2V_ERR_ARG;
  if ((H & 1) || (W & 1) || N <= 0 || N > 0x7fffffffLL) return T2V_ERR_ARG;
  p.N = (int)N; p.D = D; p.H = H; p.W = W; p.c8 = c8; p.c2 = c2; p.C8p = C8p; p.C2p = C2p;
  p.P = D * H * W; p.Hp = H / 2; p.Wp = W / 2; p.Kp = D * p.Hp * p.Wp;
  return T2V_OK;
}

}  // namespace t2v

using namespace t2v;

extern "C" {

int t2v_attention_fwd(const void* theta, const void* phi, const void* g, void* o, int64_t N, int32_t D, int32_t H,
                      int32_t W, int32_t c8, int32_t c2, int32_t C8p, int32_t C2p, void* stream) {
  AttnParams p{};
  int rc = attn_fill(p, N, D, H, W, c8, c2, C8p, C2p);
  if (rc) return rc;
  p.theta = reinterpret_cast<const __nv_bfloat16*>(theta);
  p.phi = reinterpret_cast<const __nv_bfloat16*>(phi);
  p.g = reinterpret_cast<const __nv_bfloat16*>(g);
  p.o = reinterpret_cast<__nv_bfloat16*>(o);
  const size_t smem = sizeof(float) * (size_t)p.Kp * (c8 + c2);
  if (smem > 200 * 1024) return T2V_ERR_ARG;
  cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid((unsigned)N, (unsigned)((p.P + 255) / 256), 1);
  attn_fwd_kernel<<<grid, 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  count_launch();
  return check_last("attention_fwd");
}

int t2v_attention_bwd(const void* theta, const void* phi, const void* g, const void* dout, void* dtheta, void* dphi,
                      void* dg, int64_t N, int32_t D, int32_t H, int32_t W, int32_t c8, int32_t c2, int32_t C8p,
                      int32_t C2p, void* stream) {
  AttnParams p{};
  int rc = attn_fill(p, N, D, H, W, c8, c2, C8p, C2p);
  if (rc) return rc;
  p.theta = reinterpret_cast<const __nv_bfloat16*>(theta);
  p.phi = reinterpret_cast<const __nv_bfloat16*>(phi);
  p.g = reinterpret_cast<const __nv_bfloat16*>(g);
  p.dout = reinterpret_cast<const __nv_bfloat16*>(dout);
  p.dtheta = reinterpret_cast<__nv_bfloat16*>(dtheta);
  p.dphi = reinterpret_cast<__nv_bfloat16*>(dphi);
  p.dg = reinterpret_cast<__nv_bfloat16*>(dg);
  const size_t smem = sizeof(float) * ((size_t)p.Kp * (c8 + c2) + (size_t)p.P * (c8 + c2 + 3)) + (size_t)p.Kp * (c8 + c2);
  if (smem > 200 * 1024) return T2V_ERR_ARG;
  cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  attn_bwd_kernel<<<(unsigned)N, 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  count_launch();
  return check_last("attention_bwd");
}

}  // extern "C"
