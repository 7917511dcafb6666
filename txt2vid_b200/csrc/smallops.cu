// Small fp32 kernels that take the last library calls off the training step:
//   * non-local block core as differentiable primitives (txt2vid/models/layers.py:23-36, 52-68): 1x2x2 max-pool with
//     recorded arg-max, batched small GEMM, row softmax with first- and second-order backward.  The discriminator's
//     block sits on the gradient-penalty path (gan/losses.py:169-178), so every primitive's backward is again one of
//     these primitives (ops.py) -- no ATen max_pool3d / cuBLAS bmm / softmax;
//   * the two Linear(.,1) heads of the discriminator (models/resnet3d.py:50-55) as row dot products with their
//     data / weight gradients (closed under the second differentiation of the penalty);
//   * relativistic / Wasserstein loss reduction over all pyramid levels and prediction pairs in ONE launch forward and
//     ONE backward (gan/losses.py:55-85, gan/cond_gan.py:51-61);
//   * gradient-penalty arithmetic: per-sample interpolation (gan/losses.py:140-145) and the fp32 dot / scale kernels
//     behind sum ||g||^2 and its derivative (losses.py:180-186);
//   * bf16x3 operand split of fp32 activations for the fp32 parity mode of the tcgen05 engine.
#include "t2v_common.cuh"

namespace t2v {

static inline unsigned nblocks(long long n, int per) {
  long long b = (n + per - 1) / per;
  return (unsigned)(b < 1 ? 1 : b);
}

// ------------------------------------------------------------------------------------ bf16 operand split
// x fp32 [rows][C] -> bf16 parts a = bf16(x), b = bf16(x - a), c = bf16(x - a - b) (x = a + b + c to 24 bits).
// terms = 3 (16 bits per operand): products a a' + b a' + a b'
//   layout 0: out [rows][3C]  = [a | b | a]         (K-concatenated A operand of fprop / dgrad; weights [a | a | b])
//   layout 1: out [3*rows][C] = [a ; b ; a]         (position-concatenated dy operand of wgrad)
//   layout 2: out [3*rows][C] = [a ; a ; b]         (position-concatenated x operand of wgrad)
// terms = 6 (24 bits per operand = fp32 operands): a a' + b a' + a b' + c a' + a c' + b b'
//   layout 0 / 1: [a | b | a | c | a | b],  layout 2 (and the weight packs): [a | a | b | a | c | b]
// so that sum_k a_k b_k over the concatenated axis is the fp32 product on the bf16 tensor pipe with fp32
// accumulation (dropped terms b c', c b', c c' <= 2^-24 relative).
__global__ void split_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, long long rows, int C,
                             int layout, int terms, long long total4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const int c4 = C / 4;
  const long long r = i / c4;
  const int c = (int)(i % c4) * 4;
  const float4 v = *reinterpret_cast<const float4*>(x + r * C + c);
  const float f[4] = {v.x, v.y, v.z, v.w};
  float pa[4], pb[4], pc[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    pa[j] = bf2f(f2bf(f[j]));
    const float r1 = f[j] - pa[j];                 // exact
    pb[j] = bf2f(f2bf(r1));
    pc[j] = r1 - pb[j];                            // exact; fits bf16 up to the last fp32 bits
  }
  uint2 P[3];
  P[0] = make_uint2(pack_bf16x2(pa[0], pa[1]), pack_bf16x2(pa[2], pa[3]));
  P[1] = make_uint2(pack_bf16x2(pb[0], pb[1]), pack_bf16x2(pb[2], pb[3]));
  P[2] = make_uint2(pack_bf16x2(pc[0], pc[1]), pack_bf16x2(pc[2], pc[3]));
  // part index per segment: x-side order (layouts 0, 1) and w-side order (layout 2)
  const int xs[6] = {0, 1, 0, 2, 0, 1}, ws[6] = {0, 0, 1, 0, 2, 1};
  const long long seg = layout == 0 ? (long long)C : rows * C;
  __nv_bfloat16* o = layout == 0 ? out + r * terms * C + c : out + r * C + c;
#pragma unroll
  for (int t = 0; t < 6; ++t)
    if (t < terms) *reinterpret_cast<uint2*>(o + t * seg) = P[layout == 2 ? ws[t] : xs[t]];
}

// ------------------------------------------------------------------------------------ 1x2x2 max-pool with arg-max
// x fp32 [maps][H][W][c] (maps = N*D) -> y [maps][H/2][W/2][c], idx u8 in 0..3 (first maximum, as ATen)
__global__ void maxpool122_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                      unsigned char* __restrict__ idx, int H, int W, int c, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int Hp = H / 2, Wp = W / 2;
  long long t = i;
  const int ch = (int)(t % c); t /= c;
  const int wp = (int)(t % Wp); t /= Wp;
  const int hp = (int)(t % Hp); t /= Hp;
  const long long m = t;
  const float* base = x + ((m * H + 2 * hp) * W + 2 * wp) * c + ch;
  const float v[4] = {base[0], base[c], base[(long long)W * c], base[(long long)W * c + c]};
  float best = v[0];
  int arg = 0;
#pragma unroll
  for (int j = 1; j < 4; ++j)
    if (v[j] > best || (v[j] != v[j] && best == best)) { best = v[j]; arg = j; }
  y[i] = best;
  idx[i] = (unsigned char)arg;
}
// gather: y[pooled] = x[window voxel idx]   (linear in x: the backward of scatter)
__global__ void pool122_gather_kernel(const float* __restrict__ x, const unsigned char* __restrict__ idx,
                                      float* __restrict__ y, int H, int W, int c, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int Hp = H / 2, Wp = W / 2;
  long long t = i;
  const int ch = (int)(t % c); t /= c;
  const int wp = (int)(t % Wp); t /= Wp;
  const int hp = (int)(t % Hp); t /= Hp;
  const int a = idx[i];
  y[i] = x[((t * H + 2 * hp + (a >> 1)) * W + 2 * wp + (a & 1)) * c + ch];
}
// scatter: dx[window voxel idx] = dy[pooled], zeros elsewhere (one thread per pooled element writes its 4 voxels)
__global__ void pool122_scatter_kernel(const float* __restrict__ dy, const unsigned char* __restrict__ idx,
                                       float* __restrict__ dx, int H, int W, int c, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int Hp = H / 2, Wp = W / 2;
  long long t = i;
  const int ch = (int)(t % c); t /= c;
  const int wp = (int)(t % Wp); t /= Wp;
  const int hp = (int)(t % Hp); t /= Hp;
  const int a = idx[i];
  const float g = dy[i];
  float* base = dx + ((t * H + 2 * hp) * W + 2 * wp) * c + ch;
  base[0] = a == 0 ? g : 0.f;
  base[c] = a == 1 ? g : 0.f;
  base[(long long)W * c] = a == 2 ? g : 0.f;
  base[(long long)W * c + c] = a == 3 ? g : 0.f;
}

// ------------------------------------------------------------------------------------ batched small GEMM (fp32)
// C[b] (M x N, row-major) = op(A[b]) * op(B[b]);  ta: A[b] stored (K x M);  tb: B[b] stored (N x K)
struct BmmParams {
  const float* A; const float* B; float* C;
  int M, N, K, ta, tb;
  long long sA, sB, sC;
};
__global__ void __launch_bounds__(256) bmm_kernel(const BmmParams p) {
  __shared__ float As[16][65], Bs[16][65];
  const int b = blockIdx.z;
  const float* A = p.A + b * p.sA;
  const float* B = p.B + b * p.sB;
  float* C = p.C + b * p.sC;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < p.K; k0 += 16) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = threadIdx.x + 256 * i;
      // A tile: element (k, m)
      {
        const int k = p.ta ? e / 64 : e % 16, m = p.ta ? e % 64 : e / 16;
        float v = 0.f;
        if (k0 + k < p.K && m0 + m < p.M)
          v = p.ta ? A[(long long)(k0 + k) * p.M + m0 + m] : A[(long long)(m0 + m) * p.K + k0 + k];
        As[k][m] = v;
      }
      {
        const int k = p.tb ? e % 16 : e / 64, n = p.tb ? e / 16 : e % 64;
        float v = 0.f;
        if (k0 + k < p.K && n0 + n < p.N)
          v = p.tb ? B[(long long)(n0 + n) * p.K + k0 + k] : B[(long long)(k0 + k) * p.N + n0 + n];
        Bs[k][n] = v;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[4], bb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[k][ty * 4 + i]; bb[i] = Bs[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < p.N) C[(long long)m * p.N + n] = acc[i][j];
    }
  }
}

// ------------------------------------------------------------------------------------ row softmax (+ 1st / 2nd order bwd)
// one warp per row
__global__ void softmax_fwd_kernel(const float* __restrict__ S, float* __restrict__ out, long long rows, int cols) {
  const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float* s = S + r * cols;
  float m = -INFINITY;
  for (int j = lane; j < cols; j += 32) m = fmaxf(m, s[j]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float l = 0.f;
  for (int j = lane; j < cols; j += 32) l += expf(s[j] - m);
  l = warp_sum(l);
  const float inv = 1.f / l;
  for (int j = lane; j < cols; j += 32) out[r * cols + j] = expf(s[j] - m) * inv;
}
// dS = beta * (dbeta - sum_j beta_j dbeta_j)
__global__ void softmax_bwd_kernel(const float* __restrict__ beta, const float* __restrict__ dbeta,
                                   float* __restrict__ dS, long long rows, int cols) {
  const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float* b = beta + r * cols;
  const float* d = dbeta + r * cols;
  float s = 0.f;
  for (int j = lane; j < cols; j += 32) s = fmaf(b[j], d[j], s);
  s = warp_sum(s);
  for (int j = lane; j < cols; j += 32) dS[r * cols + j] = b[j] * (d[j] - s);
}
// derivative of dS = softmax_bwd(beta, dbeta) against a cotangent u:
//   g_dbeta = beta * (u - t),  g_beta = u * (dbeta - s) - dbeta * t,   s = sum beta dbeta, t = sum beta u
__global__ void softmax_bwd_bwd_kernel(const float* __restrict__ beta, const float* __restrict__ dbeta,
                                       const float* __restrict__ u, float* __restrict__ g_beta,
                                       float* __restrict__ g_dbeta, long long rows, int cols) {
  const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float* b = beta + r * cols;
  const float* d = dbeta + r * cols;
  const float* uu = u + r * cols;
  float s = 0.f, t = 0.f;
  for (int j = lane; j < cols; j += 32) { s = fmaf(b[j], d[j], s); t = fmaf(b[j], uu[j], t); }
  s = warp_sum(s);
  t = warp_sum(t);
  for (int j = lane; j < cols; j += 32) {
    if (g_dbeta) g_dbeta[r * cols + j] = b[j] * (uu[j] - t);
    if (g_beta) g_beta[r * cols + j] = uu[j] * (d[j] - s) - d[j] * t;
  }
}

// ------------------------------------------------------------------------------------ Linear(., 1) heads
// out[b] = feat[b,:] . w[0:F] + cond[b,:] . w[F:F+E] + bias;  one warp per row
__global__ void head_fwd_kernel(const float* __restrict__ feat, const float* __restrict__ cond,
                                const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ out,
                                int B, int F, int E) {
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= B) return;
  float acc = 0.f;
  const float* f = feat + (long long)r * F;
  for (int j = lane; j < F; j += 32) acc = fmaf(f[j], w[j], acc);
  if (cond != nullptr) {
    const float* c = cond + (long long)r * E;
    for (int j = lane; j < E; j += 32) acc = fmaf(c[j], w[F + j], acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) out[r] = acc + (bias ? bias[0] : 0.f);
}
// dfeat[b, j] = dpred[b] * w[j],  dcond[b, j] = dpred[b] * w[F + j]
__global__ void head_bwd_data_kernel(const float* __restrict__ dpred, const float* __restrict__ w,
                                     float* __restrict__ dfeat, float* __restrict__ dcond, int B, int F, int E) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int FE = F + (dcond ? E : 0);
  if (i >= (long long)B * FE) return;
  const int b = (int)(i / FE), j = (int)(i % FE);
  const float v = dpred[b] * w[j];
  if (j < F) dfeat[(long long)b * F + j] = v;
  else dcond[(long long)b * E + j - F] = v;
}
// dw[j] (+)= sum_b dpred[b] * [feat | cond][b, j];  db (+)= sum_b dpred[b].  One thread per column j, rows in chunks
// of 64 per block (grid.y) combined with one atomicAdd per column and block; dw / db zeroed by the host when not
// accumulating.
__global__ void head_bwd_weight_kernel(const float* __restrict__ dpred, const float* __restrict__ feat,
                                       const float* __restrict__ cond, float* __restrict__ dw, float* __restrict__ db,
                                       int B, int F, int E) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int FE = F + (cond ? E : 0);
  const int b0 = blockIdx.y * 64, b1 = min(B, b0 + 64);
  if (j < FE) {
    float acc = 0.f;
    for (int b = b0; b < b1; ++b)
      acc = fmaf(dpred[b], j < F ? feat[(long long)b * F + j] : cond[(long long)b * E + j - F], acc);
    atomicAdd(dw + j, acc);
  }
  if (db != nullptr && j == 0) {
    float s = 0.f;
    for (int b = b0; b < b1; ++b) s += dpred[b];
    atomicAdd(db, s);
  }
}

// ------------------------------------------------------------------------------------ fused loss reduction
// out = sum_e weight_e / n_e * sum_j f(b_e[j] - a_e[j]);  mode 0: f = softplus (RSGAN: BCE-with-logits of (a - b) against
// ones, gan/losses.py:74-85);  mode 1: f = identity (Wasserstein critic difference, losses.py:55-68)
struct LossEntries {
  static constexpr int kMax = 24;
  const float* a[kMax];
  const float* b[kMax];
  float* da[kMax];
  float* db[kMax];
  int n[kMax];
  float weight[kMax];
  int count;
};
T2V_DEVINL float softplus_f(float x) { return x > 20.f ? x : log1pf(expf(x)); }
__global__ void __launch_bounds__(256) rel_loss_fwd_kernel(const LossEntries e, int mode, float* __restrict__ out) {
  __shared__ float red[8];
  float acc = 0.f;
  for (int k = 0; k < e.count; ++k) {
    const float sc = e.weight[k] / (float)e.n[k];
    float s = 0.f;
    for (int j = threadIdx.x; j < e.n[k]; j += blockDim.x) {
      const float d = e.b[k][j] - e.a[k][j];
      s += mode == 0 ? softplus_f(d) : d;
    }
    acc = fmaf(sc, s, acc);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    out[0] = t;
  }
}
// da[j] (+)= -g * weight/n * f'(b - a),  db[j] (+)= +...  (several entries may share a tensor: the host zeroes the
// gradient buffers once and every entry accumulates; entries run sequentially inside one block, so no atomics)
__global__ void __launch_bounds__(256) rel_loss_bwd_kernel(const LossEntries e, int mode,
                                                           const float* __restrict__ gout) {
  const float g = gout[0];
  for (int k = 0; k < e.count; ++k) {
    const float sc = g * e.weight[k] / (float)e.n[k];
    for (int j = threadIdx.x; j < e.n[k]; j += blockDim.x) {
      const float d = e.b[k][j] - e.a[k][j];
      const float fp = mode == 0 ? 1.f / (1.f + expf(-d)) : 1.f;
      if (e.da[k]) e.da[k][j] -= sc * fp;
      if (e.db[k]) e.db[k][j] += sc * fp;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------ gradient-penalty arithmetic
// out[b, s] = alpha[b] * real[b, s] + (1 - alpha[b]) * fake[b, s]        (gan/losses.py:140-145)
__global__ void lerp_rows_kernel(const float* __restrict__ real, const float* __restrict__ fake,
                                 const float* __restrict__ alpha, float* __restrict__ out, long long S,
                                 long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const float a = alpha[i / S];
  out[i] = a * real[i] + (1.f - a) * fake[i];
}

}  // namespace t2v

using namespace t2v;
#define STREAM reinterpret_cast<cudaStream_t>(stream)

extern "C" {

int t2v_split_bf16(const float* x, void* out, int64_t rows, int32_t C, int32_t layout, int32_t terms, void* stream) {
  if (!x || !out || C % 4 || layout < 0 || layout > 2 || (terms != 3 && terms != 6)) return T2V_ERR_ARG;
  const long long total4 = rows * C / 4;
  if (total4 == 0) return T2V_OK;
  split_kernel<<<nblocks(total4, 256), 256, 0, STREAM>>>(x, reinterpret_cast<__nv_bfloat16*>(out), rows, C, layout,
                                                         terms, total4);
  count_launch();
  return check_last("split_bf16");
}
int t2v_split_bf16x3(const float* x, void* out, int64_t rows, int32_t C, int32_t layout, void* stream) {
  return t2v_split_bf16(x, out, rows, C, layout, 3, stream);
}

int t2v_maxpool122_fwd(const float* x, float* y, void* idx, int64_t maps, int32_t H, int32_t W, int32_t c,
                       void* stream) {
  if (H % 2 || W % 2) return T2V_ERR_ARG;
  const long long total = maps * (H / 2) * (W / 2) * c;
  if (total == 0) return T2V_OK;
  maxpool122_fwd_kernel<<<nblocks(total, 256), 256, 0, STREAM>>>(x, y, reinterpret_cast<unsigned char*>(idx), H, W, c,
                                                                 total);
  count_launch();
  return check_last("maxpool122_fwd");
}
int t2v_pool122_gather(const float* x, const void* idx, float* y, int64_t maps, int32_t H, int32_t W, int32_t c,
                       void* stream) {
  if (H % 2 || W % 2) return T2V_ERR_ARG;
  const long long total = maps * (H / 2) * (W / 2) * c;
  if (total == 0) return T2V_OK;
  pool122_gather_kernel<<<nblocks(total, 256), 256, 0, STREAM>>>(x, reinterpret_cast<const unsigned char*>(idx), y, H,
                                                                 W, c, total);
  count_launch();
  return check_last("pool122_gather");
}
int t2v_pool122_scatter(const float* dy, const void* idx, float* dx, int64_t maps, int32_t H, int32_t W, int32_t c,
                        void* stream) {
  if (H % 2 || W % 2) return T2V_ERR_ARG;
  const long long total = maps * (H / 2) * (W / 2) * c;
  if (total == 0) return T2V_OK;
  pool122_scatter_kernel<<<nblocks(total, 256), 256, 0, STREAM>>>(dy, reinterpret_cast<const unsigned char*>(idx), dx,
                                                                  H, W, c, total);
  count_launch();
  return check_last("pool122_scatter");
}

int t2v_bmm_f32(const float* A, const float* B, float* C, int32_t batch, int32_t M, int32_t N, int32_t K,
                int32_t trans_a, int32_t trans_b, void* stream) {
  if (!A || !B || !C || batch < 0 || M < 0 || N < 0 || K < 0 || batch > 65535) return T2V_ERR_ARG;
  if (batch == 0 || M == 0 || N == 0) return T2V_OK;
  BmmParams p{A, B, C, M, N, K, trans_a ? 1 : 0, trans_b ? 1 : 0, (long long)M * K, (long long)K * N, (long long)M * N};
  dim3 grid((N + 63) / 64, (M + 63) / 64, batch);
  bmm_kernel<<<grid, 256, 0, STREAM>>>(p);
  count_launch();
  return check_last("bmm_f32");
}

int t2v_softmax_fwd(const float* S, float* out, int64_t rows, int32_t cols, void* stream) {
  if (rows == 0) return T2V_OK;
  softmax_fwd_kernel<<<nblocks(rows * 32, 256), 256, 0, STREAM>>>(S, out, rows, cols);
  count_launch();
  return check_last("softmax_fwd");
}
int t2v_softmax_bwd(const float* beta, const float* dbeta, float* dS, int64_t rows, int32_t cols, void* stream) {
  if (rows == 0) return T2V_OK;
  softmax_bwd_kernel<<<nblocks(rows * 32, 256), 256, 0, STREAM>>>(beta, dbeta, dS, rows, cols);
  count_launch();
  return check_last("softmax_bwd");
}
int t2v_softmax_bwd_bwd(const float* beta, const float* dbeta, const float* u, float* g_beta, float* g_dbeta,
                        int64_t rows, int32_t cols, void* stream) {
  if (rows == 0) return T2V_OK;
  softmax_bwd_bwd_kernel<<<nblocks(rows * 32, 256), 256, 0, STREAM>>>(beta, dbeta, u, g_beta, g_dbeta, rows, cols);
  count_launch();
  return check_last("softmax_bwd_bwd");
}

int t2v_head_fwd(const float* feat, const float* cond, const float* w, const float* bias, float* out, int32_t B,
                 int32_t F, int32_t E, void* stream) {
  if (!feat || !w || !out) return T2V_ERR_ARG;
  if (B == 0) return T2V_OK;
  head_fwd_kernel<<<nblocks((long long)B * 32, 256), 256, 0, STREAM>>>(feat, cond, w, bias, out, B, F, cond ? E : 0);
  count_launch();
  return check_last("head_fwd");
}
int t2v_head_bwd_data(const float* dpred, const float* w, float* dfeat, float* dcond, int32_t B, int32_t F, int32_t E,
                      void* stream) {
  if (!dpred || !w || !dfeat) return T2V_ERR_ARG;
  const long long total = (long long)B * (F + (dcond ? E : 0));
  if (total == 0) return T2V_OK;
  head_bwd_data_kernel<<<nblocks(total, 256), 256, 0, STREAM>>>(dpred, w, dfeat, dcond, B, F, E);
  count_launch();
  return check_last("head_bwd_data");
}
int t2v_head_bwd_weight(const float* dpred, const float* feat, const float* cond, float* dw, float* db, int32_t B,
                        int32_t F, int32_t E, int32_t accumulate, void* stream) {
  if (!dpred || !feat || !dw) return T2V_ERR_ARG;
  const int FE = F + (cond ? E : 0);
  if (!accumulate) {
    cudaMemsetAsync(dw, 0, sizeof(float) * FE, STREAM);
    if (db) cudaMemsetAsync(db, 0, sizeof(float), STREAM);
  }
  if (B == 0) return T2V_OK;
  dim3 grid((FE + 127) / 128, (B + 63) / 64, 1);
  head_bwd_weight_kernel<<<grid, 128, 0, STREAM>>>(dpred, feat, cond, dw, db, B, F, E);
  count_launch();
  return check_last("head_bwd_weight");
}

static int fill_entries(LossEntries& e, int32_t count, const float* const* a, const float* const* b, float* const* da,
                        float* const* db, const int32_t* n, const float* weight) {
  if (count < 0 || count > LossEntries::kMax) return T2V_ERR_ARG;
  e.count = count;
  for (int i = 0; i < count; ++i) {
    if (!a[i] || !b[i] || n[i] <= 0) return T2V_ERR_ARG;
    e.a[i] = a[i]; e.b[i] = b[i];
    e.da[i] = da ? da[i] : nullptr;
    e.db[i] = db ? db[i] : nullptr;
    e.n[i] = n[i]; e.weight[i] = weight[i];
  }
  return T2V_OK;
}
int t2v_rel_loss_fwd(int32_t count, const float* const* host_a, const float* const* host_b, const int32_t* host_n,
                     const float* host_weight, int32_t mode, float* out, void* stream) {
  LossEntries e;
  int rc = fill_entries(e, count, host_a, host_b, nullptr, nullptr, host_n, host_weight);
  if (rc) return rc;
  rel_loss_fwd_kernel<<<1, 256, 0, STREAM>>>(e, mode, out);
  count_launch();
  return check_last("rel_loss_fwd");
}
int t2v_rel_loss_bwd(int32_t count, const float* const* host_a, const float* const* host_b, float* const* host_da,
                     float* const* host_db, const int32_t* host_n, const float* host_weight, int32_t mode,
                     const float* gout, void* stream) {
  LossEntries e;
  int rc = fill_entries(e, count, host_a, host_b, host_da, host_db, host_n, host_weight);
  if (rc) return rc;
  rel_loss_bwd_kernel<<<1, 256, 0, STREAM>>>(e, mode, gout);
  count_launch();
  return check_last("rel_loss_bwd");
}

int t2v_lerp_rows(const float* real, const float* fake, const float* alpha, float* out, int64_t B, int64_t S,
                  void* stream) {
  const long long total = B * S;
  if (total == 0) return T2V_OK;
  lerp_rows_kernel<<<nblocks(total, 256), 256, 0, STREAM>>>(real, fake, alpha, out, S, total);
  count_launch();
  return check_last("lerp_rows");
}

}  // extern "C"
