// HBM-bound kernels of the TGANv2 training step: activation / pooling / normalisation / layout /
// index / reduction work on channels-last bf16 tensors.  One pass over the data each, 128-bit
// accesses where the channel count allows (C % 8 == 0), fp32 math.
//
// Reference call sites (txt2vid/...): ReLU models/layers.py:172,176,230,232,250; AvgPool
// layers.py:202-217 + resnet3d.py:16,18; nn.Upsample layers.py:168,180; BatchNorm2d(train)
// layers.py:171,175,249; tanh layers.py:252; torch.sum(x,[2,3,4]) resnet3d.py:48; Subsample
// layers.py:106-111; F.interpolate gan/trainer.py:149.
#include "t2v_common.cuh"

namespace t2v {

static inline unsigned blocks_for(long long n, int per_block) {
  long long b = (n + per_block - 1) / per_block;
  if (b < 1) b = 1;
  return (unsigned)b;
}

struct V8 { float f[8]; };
T2V_DEVINL V8 ld8(const __nv_bfloat16* p) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  V8 r;
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  r.f[0] = a.x; r.f[1] = a.y; r.f[2] = b.x; r.f[3] = b.y; r.f[4] = c.x; r.f[5] = c.y; r.f[6] = d.x; r.f[7] = d.y;
  return r;
}
T2V_DEVINL void st8(__nv_bfloat16* p, const V8& v) {
  *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16x2(v.f[0], v.f[1]), pack_bf16x2(v.f[2], v.f[3]),
                                            pack_bf16x2(v.f[4], v.f[5]), pack_bf16x2(v.f[6], v.f[7]));
}

// ------------------------------------------------------------------------------------ ReLU
__global__ void relu_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n8) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  V8 v = ld8(x + i * 8);
#pragma unroll
  for (int j = 0; j < 8; ++j) v.f[j] = fmaxf(v.f[j], 0.f);
  st8(y + i * 8, v);
}
// dx = dy where ref > 0 (ref = relu output or pre-activation: same mask)
__global__ void relu_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ ref,
                                __nv_bfloat16* __restrict__ dx, long long n8) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  V8 g = ld8(dy + i * 8);
  const V8 r = ld8(ref + i * 8);
#pragma unroll
  for (int j = 0; j < 8; ++j) g.f[j] = r.f[j] > 0.f ? g.f[j] : 0.f;
  st8(dx + i * 8, g);
}

// y = x > 0 ? x : slope * x  (nn.LeakyReLU(0.2), models/tcwyt/*.py) and its gradient (ref = x or y: same sign)
__global__ void leaky_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, float slope,
                                 long long n8) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  V8 v = ld8(x + i * 8);
#pragma unroll
  for (int j = 0; j < 8; ++j) v.f[j] = v.f[j] > 0.f ? v.f[j] : slope * v.f[j];
  st8(y + i * 8, v);
}
__global__ void leaky_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ ref,
                                 __nv_bfloat16* __restrict__ dx, float slope, long long n8) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  V8 g = ld8(dy + i * 8);
  const V8 r = ld8(ref + i * 8);
#pragma unroll
  for (int j = 0; j < 8; ++j) g.f[j] = r.f[j] > 0.f ? g.f[j] : slope * g.f[j];
  st8(dx + i * 8, g);
}
// tanh on a CL tensor (models/tgan/temporal_gen.py:33) and dx = dy * (1 - y^2)
__global__ void tanh_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n8) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  V8 v = ld8(x + i * 8);
#pragma unroll
  for (int j = 0; j < 8; ++j) v.f[j] = tanhf(v.f[j]);
  st8(y + i * 8, v);
}
__global__ void tanh_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ y,
                                __nv_bfloat16* __restrict__ dx, long long n8) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  V8 g = ld8(dy + i * 8);
  const V8 r = ld8(y + i * 8);
#pragma unroll
  for (int j = 0; j < 8; ++j) g.f[j] *= 1.f - r.f[j] * r.f[j];
  st8(dx + i * 8, g);
}

// ------------------------------------------------------------------------------------ avg-pool
struct PoolParams {
  int N, D, H, W, C, Do, Ho, Wo;
  int kd, kh, kw, sd, sh, sw, pd, ph, pw;
  float inv;
};
// y[n,do,ho,wo,c] = mean over window (zero padding counted) (+ residual)
// (IDX = unsigned when every element index fits 32 bits: the coordinate decode is a chain of divisions by run-time
// values, ~5x cheaper in 32-bit arithmetic -- the 64-bit form made these kernels instruction bound)
template <typename IDX>
__global__ void avgpool_fwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ res,
                                   __nv_bfloat16* __restrict__ y, const PoolParams p, long long total8) {
  const IDX i = (IDX)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (IDX)total8) return;
  const int c8 = p.C / 8;
  IDX t = i;
  const int c = (int)(t % c8) * 8; t /= c8;
  const int wo = (int)(t % p.Wo); t /= p.Wo;
  const int ho = (int)(t % p.Ho); t /= p.Ho;
  const int dz = (int)(t % p.Do); t /= p.Do;
  const int n = (int)t;
  V8 acc;
#pragma unroll
  for (int j = 0; j < 8; ++j) acc.f[j] = 0.f;
  for (int a = 0; a < p.kd; ++a) {
    const int d = dz * p.sd - p.pd + a;
    if (d < 0 || d >= p.D) continue;
    for (int b = 0; b < p.kh; ++b) {
      const int h = ho * p.sh - p.ph + b;
      if (h < 0 || h >= p.H) continue;
      for (int e = 0; e < p.kw; ++e) {
        const int w = wo * p.sw - p.pw + e;
        if (w < 0 || w >= p.W) continue;
        const V8 v = ld8(x + ((((IDX)n * p.D + d) * p.H + h) * p.W + w) * p.C + c);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc.f[j] += v.f[j];
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) acc.f[j] *= p.inv;
  if (res != nullptr) {
    const V8 r = ld8(res + i * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc.f[j] += r.f[j];
  }
  st8(y + i * 8, acc);
}
// dx[n,d,h,w,c] = dy[window containing (d,h,w)] * inv   (requires kernel <= stride: at most one window)
template <typename IDX>
__global__ void avgpool_bwd_kernel(const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dx,
                                   const PoolParams p, long long total8) {
  const IDX i = (IDX)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (IDX)total8) return;
  const int c8 = p.C / 8;
  IDX t = i;
  const int c = (int)(t % c8) * 8; t /= c8;
  const int w = (int)(t % p.W); t /= p.W;
  const int h = (int)(t % p.H); t /= p.H;
  const int d = (int)(t % p.D); t /= p.D;
  const int n = (int)t;
  V8 g;
#pragma unroll
  for (int j = 0; j < 8; ++j) g.f[j] = 0.f;
  const int qd = (d + p.pd) / p.sd, rd = (d + p.pd) % p.sd;
  const int qh = (h + p.ph) / p.sh, rh = (h + p.ph) % p.sh;
  const int qw = (w + p.pw) / p.sw, rw = (w + p.pw) % p.sw;
  if (rd < p.kd && rh < p.kh && rw < p.kw && qd < p.Do && qh < p.Ho && qw < p.Wo) {
    g = ld8(dy + ((((IDX)n * p.Do + qd) * p.Ho + qh) * p.Wo + qw) * p.C + c);
#pragma unroll
    for (int j = 0; j < 8; ++j) g.f[j] *= p.inv;
  }
  st8(dx + i * 8, g);
}

// ------------------------------------------------------------------------------------ nearest x2 (H, W)
__global__ void upsample2x_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int N,
                                      int H, int W, int C, long long total8) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  const int c8 = C / 8;
  long long t = i;
  const int c = (int)(t % c8) * 8; t /= c8;
  const int wo = (int)(t % (2 * W)); t /= 2 * W;
  const int ho = (int)(t % (2 * H)); t /= 2 * H;
  const int n = (int)t;
  *reinterpret_cast<uint4*>(y + i * 8) =
      *reinterpret_cast<const uint4*>(x + (((long long)n * H + ho / 2) * W + wo / 2) * C + c);
}
__global__ void upsample2x_bwd_kernel(const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dx, int N,
                                      int H, int W, int C, long long total8) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  const int c8 = C / 8;
  long long t = i;
  const int c = (int)(t % c8) * 8; t /= c8;
  const int w = (int)(t % W); t /= W;
  const int h = (int)(t % H); t /= H;
  const int n = (int)t;
  V8 acc;
#pragma unroll
  for (int j = 0; j < 8; ++j) acc.f[j] = 0.f;
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const V8 v = ld8(dy + (((long long)n * 2 * H + 2 * h + a) * 2 * W + 2 * w + b) * C + c);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc.f[j] += v.f[j];
    }
  st8(dx + i * 8, acc);
}

// ------------------------------------------------------------------------------------ layout
// x fp32 (N, C, S) -> y bf16 (N, S, Cp), channels >= C zero
__global__ void nchw_to_cl_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int C, int Cp,
                                  long long S, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % Cp);
  const long long ns = i / Cp;
  const long long s = ns % S, n = ns / S;
  y[i] = f2bf(c < C ? x[(n * C + c) * S + s] : 0.f);
}
// RGB clip fp32 (N, 3, S) -> bf16 (N, S, 16) and / or bf16 (N, S, 4) (zero padded), one thread per voxel:
// the discriminator's input conversion (16-channel rows for the TMA-fed skip path, 4-channel rows for the gather of
// the direct stem kernel); 12 B read, 8 + 32 B written per voxel in 8 / 16-byte stores
__global__ void rgb_to_cl_kernel(const float* __restrict__ x, uint4* __restrict__ y16, uint2* __restrict__ y4,
                                 long long S, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long n = i / S, s = i - n * S;
  const float* px = x + n * 3 * S + s;
  const uint2 v = make_uint2(pack_bf16x2(px[0], px[S]), pack_bf16x2(px[2 * S], 0.f));
  if (y4 != nullptr) y4[i] = v;
  if (y16 != nullptr) {
    y16[2 * i] = make_uint4(v.x, v.y, 0u, 0u);
    y16[2 * i + 1] = make_uint4(0u, 0u, 0u, 0u);
  }
}
// x bf16 (N, S, Cp) -> y fp32 (N, C, S)
__global__ void cl_to_nchw_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y, int C, int Cp,
                                  long long S, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long s = i % S;
  const long long nc = i / S;
  const int c = (int)(nc % C);
  const long long n = nc / C;
  y[i] = bf2f(x[(n * S + s) * Cp + c]);
}

// ------------------------------------------------------------------------------------ RGB stem im2col
// The first discriminator conv (resnet3d.py:12, 3 -> 64 channels, 3^3) has K = 81: too thin for a
// tap-by-tap implicit GEMM (27 taps x 16 zero-padded channels wasted 5x the tensor work and its weight
// gradient ran at 29 TFLOP/s).  The input is tiny (3 channels), so im2col it once:
//   col[pos][tap*C + c] = x[c][pos + tap - 1]   (zero outside the clip, channels >= 27*C are zero)
// and the conv becomes a 1x1x1 GEMM with Cin = Kp on the tcgen05 engine (fprop, dgrad and wgrad).
// One thread per voxel: the 27*C neighbourhood values are gathered with compile-time tap offsets (the clip is
// tiny and L1-resident) and written as one contiguous Kp*2-byte row in 16-byte pieces.
template <int C>
__global__ void im2col3_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ col, int D, int H, int W,
                               int Kp, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  long long pos = i;
  const int w = (int)(pos % W); pos /= W;
  const int h = (int)(pos % H); pos /= H;
  const int d = (int)(pos % D);
  const long long n = pos / D;
  const long long S = (long long)D * H * W;
  const float* xn = x + n * C * S + ((long long)d * H + h) * W + w;
  constexpr int KR = 27 * C;
  constexpr int KV = (KR + 7) / 8 * 8;
  float v[KV];
#pragma unroll
  for (int tap = 0; tap < 27; ++tap) {
    const int od = tap / 9 - 1, oh = (tap / 3) % 3 - 1, ow = tap % 3 - 1;
    const bool ok = (unsigned)(d + od) < (unsigned)D && (unsigned)(h + oh) < (unsigned)H &&
                    (unsigned)(w + ow) < (unsigned)W;
    const long long off = ((long long)od * H + oh) * W + ow;
#pragma unroll
    for (int c = 0; c < C; ++c) v[tap * C + c] = ok ? __ldg(xn + c * S + off) : 0.f;
  }
#pragma unroll
  for (int k = KR; k < KV; ++k) v[k] = 0.f;
  uint4* dst = reinterpret_cast<uint4*>(col + i * Kp);
#pragma unroll
  for (int ch = 0; ch < KV / 8; ++ch)
    dst[ch] = make_uint4(pack_bf16x2(v[8 * ch], v[8 * ch + 1]), pack_bf16x2(v[8 * ch + 2], v[8 * ch + 3]),
                         pack_bf16x2(v[8 * ch + 4], v[8 * ch + 5]), pack_bf16x2(v[8 * ch + 6], v[8 * ch + 7]));
  for (int ch = KV / 8; ch < Kp / 8; ++ch) dst[ch] = make_uint4(0u, 0u, 0u, 0u);
}
// adjoint: dx[c][q] = sum_tap dcol[q - (tap - 1)][tap*C + c]; one thread per voxel q
template <int C>
__global__ void col2im3_kernel(const __nv_bfloat16* __restrict__ dcol, float* __restrict__ dx, int D, int H, int W,
                               int Kp, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  long long pos = i;
  const int w = (int)(pos % W); pos /= W;
  const int h = (int)(pos % H); pos /= H;
  const int d = (int)(pos % D);
  const long long n = pos / D;
  const long long S = (long long)D * H * W;
  float acc[C];
#pragma unroll
  for (int c = 0; c < C; ++c) acc[c] = 0.f;
#pragma unroll
  for (int tap = 0; tap < 27; ++tap) {
    const int od = tap / 9 - 1, oh = (tap / 3) % 3 - 1, ow = tap % 3 - 1;
    const bool ok = (unsigned)(d - od) < (unsigned)D && (unsigned)(h - oh) < (unsigned)H &&
                    (unsigned)(w - ow) < (unsigned)W;
    if (ok) {
      const __nv_bfloat16* row = dcol + (i - (((long long)od * H + oh) * W + ow)) * Kp + tap * C;
#pragma unroll
      for (int c = 0; c < C; ++c) acc[c] += bf2f(row[c]);
    }
  }
  const long long s = ((long long)d * H + h) * W + w;
#pragma unroll
  for (int c = 0; c < C; ++c) dx[(n * C + c) * S + s] = acc[c];
}

// ------------------------------------------------------------------------------------ reductions
// out[c] (+)= sum over rows of x[row, c]; grid.x = column groups of 64, grid.y = row slices
__global__ void sum_rows_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, long long P, int C,
                                long long rows_per_block) {
  __shared__ float red[4][64];
  const int c = blockIdx.x * 64 + (threadIdx.x & 63);
  const int lane_row = threadIdx.x >> 6;  // 0..3
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = min(P, r0 + rows_per_block);
  float acc = 0.f;
  if (c < C)
    for (long long r = r0 + lane_row; r < r1; r += 4) acc += bf2f(x[r * C + c]);
  red[lane_row][threadIdx.x & 63] = acc;
  __syncthreads();
  if (threadIdx.x < 64 && c < C)
    atomicAdd(out + c, red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x] + red[3][threadIdx.x]);
}
// out[n, c] = sum_s x[n, s, c]   (grid: (ceil(C/64), N))
__global__ void sum_spatial_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, long long S, int C) {
  __shared__ float red[4][64];
  const int c = blockIdx.x * 64 + (threadIdx.x & 63);
  const int lr = threadIdx.x >> 6;
  const long long n = blockIdx.y;
  float acc = 0.f;
  if (c < C)
    for (long long s = lr; s < S; s += 4) acc += bf2f(x[(n * S + s) * C + c]);
  red[lr][threadIdx.x & 63] = acc;
  __syncthreads();
  if (threadIdx.x < 64 && c < C)
    out[n * C + c] = red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x] + red[3][threadIdx.x];
}
// y[n, s, c] = g[n, c]
__global__ void broadcast_spatial_kernel(const float* __restrict__ g, __nv_bfloat16* __restrict__ y, long long S,
                                         int C, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  const long long n = i / C / S;
  y[i] = f2bf(g[n * C + c]);
}

// ------------------------------------------------------------------------------------ BatchNorm (train)
// activation fused behind BatchNorm: 0 = none, 1 = ReLU, 2 = LeakyReLU(0.2)
T2V_DEVINL float bn_act_slope(int act) { return act == 1 ? 0.f : (act == 2 ? 0.2f : 1.f); }
// per-channel sum and sum of squares; grid (ceil(C/64), row slices), atomics into stats[0:C], stats[C:2C]
__global__ void bn_stats_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ stats, long long P, int C,
                                long long rows_per_block) {
  __shared__ float r1[4][64], r2[4][64];
  const int c = blockIdx.x * 64 + (threadIdx.x & 63);
  const int lr = threadIdx.x >> 6;
  const long long a = (long long)blockIdx.y * rows_per_block, b = min(P, a + rows_per_block);
  float s1 = 0.f, s2 = 0.f;
  if (c < C)
    for (long long r = a + lr; r < b; r += 4) {
      const float v = bf2f(x[r * C + c]);
      s1 += v;
      s2 = fmaf(v, v, s2);
    }
  r1[lr][threadIdx.x & 63] = s1;
  r2[lr][threadIdx.x & 63] = s2;
  __syncthreads();
  if (threadIdx.x < 64 && c < C) {
    const int t = threadIdx.x;
    atomicAdd(stats + c, r1[0][t] + r1[1][t] + r1[2][t] + r1[3][t]);
    atomicAdd(stats + C + c, r2[0][t] + r2[1][t] + r2[2][t] + r2[3][t]);
  }
}
// mean / invstd / fused scale+shift, running-stat update (momentum, unbiased variance)
__global__ void bn_finalize_kernel(const float* __restrict__ stats, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, float* __restrict__ mean_invstd,
                                   float* __restrict__ scale_shift, int C, float count, float eps, float momentum) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float mean = stats[c] / count;
  float var = stats[C + c] / count - mean * mean;
  var = fmaxf(var, 0.f);
  const float invstd = rsqrtf(var + eps);
  mean_invstd[c] = mean;
  mean_invstd[C + c] = invstd;
  const float sc = gamma[c] * invstd;
  scale_shift[c] = sc;
  scale_shift[C + c] = beta[c] - mean * sc;
  if (running_mean != nullptr) {
    const float unbiased = count > 1.f ? var * count / (count - 1.f) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
  }
}
// y = [relu](x*scale+shift) [nearest x2 in H,W]; x (N,H,W,C) -> y (N,uH,uW,C)
__global__ void bn_apply_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ scale_shift,
                                __nv_bfloat16* __restrict__ y, int H, int W, int C, int relu, int up, long long total8) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  const int c8 = C / 8;
  long long t = i;
  const int c = (int)(t % c8) * 8; t /= c8;
  const int Wo = W * up, Ho = H * up;
  const int wo = (int)(t % Wo); t /= Wo;
  const int ho = (int)(t % Ho); t /= Ho;
  const long long n = t;
  V8 v = ld8(x + ((n * H + ho / up) * W + wo / up) * C + c);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float r = fmaf(v.f[j], scale_shift[c + j], scale_shift[C + c + j]);
    v.f[j] = r > 0.f ? r : r * bn_act_slope(relu);
  }
  st8(y + i * 8, v);
}
// g = sum_{up x up} dy masked by relu;  red[0:C] += sum g, red[C:2C] += sum g * xhat
__global__ void bn_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                                     const float* __restrict__ scale_shift, const float* __restrict__ mean_invstd,
                                     float* __restrict__ red, long long P, int H, int W, int C, int relu, int up,
                                     long long rows_per_block) {
  __shared__ float r1[4][64], r2[4][64];
  const int c = blockIdx.x * 64 + (threadIdx.x & 63);
  const int lr = threadIdx.x >> 6;
  const long long a = (long long)blockIdx.y * rows_per_block, b = min(P, a + rows_per_block);
  float s1 = 0.f, s2 = 0.f;
  if (c < C) {
    const float sc = scale_shift[c], sh = scale_shift[C + c], mean = mean_invstd[c], invstd = mean_invstd[C + c];
    for (long long r = a + lr; r < b; r += 4) {
      const float xv = bf2f(x[r * C + c]);
      const float am = fmaf(xv, sc, sh) > 0.f ? 1.f : bn_act_slope(relu);
      if (am == 0.f) continue;
      float g = 0.f;
      if (up == 1) {
        g = bf2f(dy[r * C + c]);
      } else {
        const int w = (int)(r % W);
        const int h = (int)((r / W) % H);
        const long long n = r / W / H;
        for (int u = 0; u < 2; ++u)
          for (int v = 0; v < 2; ++v)
            g += bf2f(dy[((n * 2 * H + 2 * h + u) * 2 * W + 2 * w + v) * C + c]);
      }
      g *= am;
      s1 += g;
      s2 = fmaf(g, (xv - mean) * invstd, s2);
    }
  }
  r1[lr][threadIdx.x & 63] = s1;
  r2[lr][threadIdx.x & 63] = s2;
  __syncthreads();
  if (threadIdx.x < 64 && c < C) {
    const int t = threadIdx.x;
    atomicAdd(red + c, r1[0][t] + r1[1][t] + r1[2][t] + r1[3][t]);           // dbeta
    atomicAdd(red + C + c, r2[0][t] + r2[1][t] + r2[2][t] + r2[3][t]);       // dgamma
  }
}
// dx = gamma*invstd * (g - dbeta/P - xhat*dgamma/P)
__global__ void bn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                                    const float* __restrict__ scale_shift, const float* __restrict__ mean_invstd,
                                    const float* __restrict__ red, __nv_bfloat16* __restrict__ dx, long long P, int H,
                                    int W, int C, int relu, int up, long long total8) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  const int c8 = C / 8;
  const int c = (int)(i % c8) * 8;
  const long long r = i / c8;
  const V8 xv = ld8(x + r * C + c);
  V8 g;
#pragma unroll
  for (int j = 0; j < 8; ++j) g.f[j] = 0.f;
  if (up == 1) {
    g = ld8(dy + r * C + c);
  } else {
    const int w = (int)(r % W);
    const int h = (int)((r / W) % H);
    const long long n = r / W / H;
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        const V8 t = ld8(dy + ((n * 2 * H + 2 * h + u) * 2 * W + 2 * w + v) * C + c);
#pragma unroll
        for (int j = 0; j < 8; ++j) g.f[j] += t.f[j];
      }
  }
  const float invP = 1.f / (float)P;
  V8 o;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float sc = scale_shift[c + j], sh = scale_shift[C + c + j];
    const float mean = mean_invstd[c + j], invstd = mean_invstd[C + c + j];
    const float gg = fmaf(xv.f[j], sc, sh) > 0.f ? g.f[j] : g.f[j] * bn_act_slope(relu);
    const float xhat = (xv.f[j] - mean) * invstd;
    o.f[j] = sc * (gg - red[c + j] * invP - xhat * red[C + c + j] * invP);  // sc = gamma*invstd
  }
  st8(dx + i * 8, o);
}

// ------------------------------------------------------------------------------------ render (tanh + layout)
// pre (B*T, H, W, Cp) bf16 -> y (B, C, T, H, W) fp32 = tanh(pre[..., :C])
__global__ void render_fwd_kernel(const __nv_bfloat16* __restrict__ pre, float* __restrict__ y, int B, int T, int H,
                                  int W, int C, int Cp, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  long long t = i;
  const int w = (int)(t % W); t /= W;
  const int h = (int)(t % H); t /= H;
  const int f = (int)(t % T); t /= T;
  const int c = (int)(t % C); t /= C;
  const long long b = t;
  y[i] = tanhf(bf2f(pre[(((b * T + f) * H + h) * W + w) * Cp + c]));
}
// dpre (B*T,H,W,Cp) bf16 = dy * (1 - y^2) for c < C, 0 for padded channels
__global__ void render_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                  __nv_bfloat16* __restrict__ dpre, int B, int T, int H, int W, int C, int Cp,
                                  long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  long long t = i;
  const int c = (int)(t % Cp); t /= Cp;
  const int w = (int)(t % W); t /= W;
  const int h = (int)(t % H); t /= H;
  const int f = (int)(t % T); t /= T;
  const long long b = t;
  float g = 0.f;
  if (c < C) {
    const long long j = (((b * C + c) * T + f) * H + h) * W + w;
    const float yv = y[j];
    g = dy[j] * (1.f - yv * yv);
  }
  dpre[i] = f2bf(g);
}

// ------------------------------------------------------------------------------------ index kernels (bit-exact)
// Frame gather on merged-frame CL maps: out frame (b', t') <- in frame (sn*b', bt + st*t'); 16-byte units
__global__ void gather_frames_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int T, int To, int sn,
                                     int st, int bt_host, const int* __restrict__ bt_dev, long long frame_u4,
                                     long long total, int scatter) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int bt = bt_dev ? *bt_dev : bt_host;
  const long long e = i % frame_u4;
  const long long fo = i / frame_u4;
  const long long bo = fo / To, to = fo % To;
  const long long fi = (bo * sn) * T + bt + to * st;
  if (scatter) y[fi * frame_u4 + e] = x[i];       // x = grad of gathered, y = zero-filled grad of source
  else y[i] = x[fi * frame_u4 + e];
}
// x fp32 (B,C,T,H,W) -> y (Bo,C,To,Ho,Wo): y[b,c,t,h,w] = x[b*sn, c, bt + t*st, floor(h*H/Ho), floor(w*W/Wo)]
__global__ void pyramid_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int C, int T, int H, int W,
                               int Bo, int To, int Ho, int Wo, int sn, int st, int bt_host,
                               const int* __restrict__ bt_dev, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int bt = bt_dev ? *bt_dev : bt_host;
  long long t = i;
  const int w = (int)(t % Wo); t /= Wo;
  const int h = (int)(t % Ho); t /= Ho;
  const int f = (int)(t % To); t /= To;
  const int c = (int)(t % C); t /= C;
  const long long b = t;
  const int hs = (int)(((long long)h * H) / Ho), ws = (int)(((long long)w * W) / Wo);
  y[i] = x[((((b * sn) * C + c) * T + bt + (long long)f * st) * H + hs) * W + ws];
}

// ------------------------------------------------------------------------------------ LSTM cell (ConvLSTM + Bi-LSTM)
// gates fp32 (P, 4*Hd) laid out [i | f | g | o]; c_prev/c fp32 (P,Hd); h bf16 (P,Hd) and optional fp32 copy
__global__ void lstm_cell_fwd_kernel(const float* __restrict__ gates, const float* __restrict__ c_prev,
                                     float* __restrict__ c, __nv_bfloat16* __restrict__ h, float* __restrict__ h32,
                                     int Hd, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long pidx = i / Hd;
  const int j = (int)(i % Hd);
  const float* g = gates + pidx * 4 * Hd;
  const float gi = 1.f / (1.f + __expf(-g[j]));
  const float gf = 1.f / (1.f + __expf(-g[Hd + j]));
  const float gg = tanhf(g[2 * Hd + j]);
  const float go = 1.f / (1.f + __expf(-g[3 * Hd + j]));
  const float cp = c_prev ? c_prev[i] : 0.f;
  const float cn = gf * cp + gi * gg;
  c[i] = cn;
  const float hv = go * tanhf(cn);
  h[i] = f2bf(hv);
  if (h32) h32[i] = hv;
}
// dgates bf16 (P,4Hd), dc_prev fp32;  dh fp32 (may be null -> 0), dc_next fp32 (may be null -> 0)
__global__ void lstm_cell_bwd_kernel(const float* __restrict__ gates, const float* __restrict__ c_prev,
                                     const float* __restrict__ c, const float* __restrict__ dh,
                                     const float* __restrict__ dc_next, __nv_bfloat16* __restrict__ dgates,
                                     float* __restrict__ dc_prev, int Hd, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long pidx = i / Hd;
  const int j = (int)(i % Hd);
  const float* g = gates + pidx * 4 * Hd;
  const float gi = 1.f / (1.f + __expf(-g[j]));
  const float gf = 1.f / (1.f + __expf(-g[Hd + j]));
  const float gg = tanhf(g[2 * Hd + j]);
  const float go = 1.f / (1.f + __expf(-g[3 * Hd + j]));
  const float cp = c_prev ? c_prev[i] : 0.f;
  const float tc = tanhf(c[i]);
  const float dhv = dh ? dh[i] : 0.f;
  const float dcv = (dc_next ? dc_next[i] : 0.f) + dhv * go * (1.f - tc * tc);
  __nv_bfloat16* dg = dgates + pidx * 4 * Hd;
  dg[j] = f2bf(dcv * gg * gi * (1.f - gi));
  dg[Hd + j] = f2bf(dcv * cp * gf * (1.f - gf));
  dg[2 * Hd + j] = f2bf(dcv * gi * (1.f - gg * gg));
  dg[3 * Hd + j] = f2bf(dhv * tc * go * (1.f - go));
  dc_prev[i] = dcv * gf;
}

// ------------------------------------------------------------------------------------ Adam (multi-tensor)
struct AdamChunk {
  static constexpr int kMax = 48;
  float* p[kMax];
  const float* g[kMax];
  float* m[kMax];
  float* v[kMax];
  long long n[kMax];
  int count;
};
// dyn (optional, device): {lr / (1 - b1^t), 1 / sqrt(1 - b2^t)} for CUDA-graph replays, where the step
// count cannot be a baked-in kernel argument
__global__ void adam_kernel(const AdamChunk ch, float lr, float b1, float b2, float eps, float bc1, float bc2,
                            float grad_scale, const float* __restrict__ dyn) {
  const int t = blockIdx.y;
  if (t >= ch.count) return;
  float* p = ch.p[t];
  const float* g = ch.g[t];
  float* m = ch.m[t];
  float* v = ch.v[t];
  const long long n = ch.n[t];
  const float step = dyn ? dyn[0] : lr / bc1;
  const float inv_sqrt_bc2 = dyn ? dyn[1] : rsqrtf(bc2);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * grad_scale;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= step * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
  }
}

// ------------------------------------------------------------------------------------ multi-tensor copy
struct CopyChunk {
  static constexpr int kMax = 96;
  const float* src[kMax];
  float* dst[kMax];
  long long n[kMax];
  int count;
};
__global__ void multi_copy_kernel(const CopyChunk ch) {
  const int t = blockIdx.y;
  if (t >= ch.count) return;
  const float* s = ch.src[t];
  float* d = ch.dst[t];
  const long long n = ch.n[t];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    d[i] = s[i];
}

// Bulk host -> device copy by a few resident CTAs reading PINNED host memory over UVA and writing with streaming
// (evict-first) stores: the input batch crosses PCIe while a training step runs, without the copy engine (whose
// queue the step's own small transfers share) and without displacing the step's L2-resident operands.
__global__ void __launch_bounds__(256) stream_copy_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst,
                                                         long long n16) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n16; i += 4 * stride) {          // four loads in flight per thread
    const uint4 a = __ldcs(src + i), b = __ldcs(src + i + stride), c = __ldcs(src + i + 2 * stride),
                d = __ldcs(src + i + 3 * stride);
    __stcs(dst + i, a); __stcs(dst + i + stride, b); __stcs(dst + i + 2 * stride, c); __stcs(dst + i + 3 * stride, d);
  }
  for (; i < n16; i += stride) __stcs(dst + i, __ldcs(src + i));
}

}  // namespace t2v

using namespace t2v;
#define STREAM reinterpret_cast<cudaStream_t>(stream)
#define BF(p) reinterpret_cast<__nv_bfloat16*>(p)
#define CBF(p) reinterpret_cast<const __nv_bfloat16*>(p)

// ------------------------------------------------------------------------------------ vectorised column reductions
// Per-channel sums over the rows of a [P][C] bf16 matrix with 16-byte loads: thread = (row slot, 8 channels),
// C/8 a power of two <= 256.  One kernel body, three uses (bias gradients, BatchNorm statistics, BatchNorm
// backward reductions); partial sums are combined through shared memory and one atomicAdd per channel and block.
struct ColRed {
  long long P;
  int C, H, W, act, up;                      // H, W, act, up: BatchNorm backward only
  const __nv_bfloat16* x;
  const __nv_bfloat16* dy;
  const float* scale_shift;
  const float* mean_invstd;
  float* out;                                 // [C] (mode 0) or [2C] (modes 1, 2)
  long long rows_per_block;
};
template <int MODE>   // 0: sum x   1: sum x, sum x^2   2: BatchNorm backward {sum g, sum g*xhat}
__global__ void __launch_bounds__(256) colred_v8_kernel(const ColRed p) {
  __shared__ float sm1[256][9], sm2[256][9];
  const int Cb = p.C < 256 ? p.C : 256;       // channels per block; blockIdx.y selects the 256-channel slab
  const int ct = Cb >> 3;
  const int tc = threadIdx.x % ct, tr = threadIdx.x / ct, rpi = 256 / ct;
  const long long a = (long long)blockIdx.x * p.rows_per_block, b = min(p.P, a + p.rows_per_block);
  const int c = blockIdx.y * 256 + tc * 8;
  float s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
  float sc[8], sh[8], mean[8], invstd[8];
  if (MODE == 2) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sc[j] = p.scale_shift[c + j]; sh[j] = p.scale_shift[p.C + c + j];
      mean[j] = p.mean_invstd[c + j]; invstd[j] = p.mean_invstd[p.C + c + j];
    }
  }
  const float slope = bn_act_slope(p.act);
#pragma unroll 2
  for (long long r = a + tr; r < b; r += rpi) {
    const V8 v = ld8(p.x + r * p.C + c);
    if (MODE == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) s1[j] += v.f[j];
    } else if (MODE == 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { s1[j] += v.f[j]; s2[j] = fmaf(v.f[j], v.f[j], s2[j]); }
    } else {
      V8 g;
      if (p.up == 1) {
        g = ld8(p.dy + r * p.C + c);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) g.f[j] = 0.f;
        const int w = (int)(r % p.W);
        const int h = (int)((r / p.W) % p.H);
        const long long n = r / p.W / p.H;
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const V8 t = ld8(p.dy + ((n * 2 * p.H + 2 * h + u) * 2 * p.W + 2 * w + q) * p.C + c);
#pragma unroll
            for (int j = 0; j < 8; ++j) g.f[j] += t.f[j];
          }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float gg = fmaf(v.f[j], sc[j], sh[j]) > 0.f ? g.f[j] : g.f[j] * slope;
        s1[j] += gg;
        s2[j] = fmaf(gg, (v.f[j] - mean[j]) * invstd[j], s2[j]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) { sm1[threadIdx.x][j] = s1[j]; if (MODE != 0) sm2[threadIdx.x][j] = s2[j]; }
  __syncthreads();
  // thread t < C sums channel t over the rpi row slots
  if ((int)threadIdx.x < Cb) {
    const int cc = threadIdx.x >> 3, j = threadIdx.x & 7;
    float t1 = 0.f, t2 = 0.f;
    for (int q = 0; q < rpi; ++q) { t1 += sm1[q * ct + cc][j]; if (MODE != 0) t2 += sm2[q * ct + cc][j]; }
    const int co = blockIdx.y * 256 + threadIdx.x;
    atomicAdd(p.out + co, t1);
    if (MODE != 0) atomicAdd(p.out + p.C + co, t2);
  }
}
static bool colred_ok(int C) {
  if (C % 8) return false;
  if (C > 256) return C % 256 == 0;                     // 256-channel slabs along grid.y
  const int ct = C / 8;
  return ct >= 1 && (ct & (ct - 1)) == 0;               // C/8 a power of two: one block covers all channels
}
template <int MODE>
static void colred_launch(ColRed p, cudaStream_t s) {
  const int slabs = p.C > 256 ? p.C / 256 : 1;
  const int rpi = 256 / ((p.C > 256 ? 256 : p.C) / 8);
  long long blocks = 8 * 148 / slabs;
  const long long maxb = (p.P + (long long)rpi * 4 - 1) / ((long long)rpi * 4);
  if (blocks > maxb) blocks = maxb;
  if (blocks < 1) blocks = 1;
  p.rows_per_block = (p.P + blocks - 1) / blocks;
  p.rows_per_block = (p.rows_per_block + rpi - 1) / rpi * rpi;
  blocks = (p.P + p.rows_per_block - 1) / p.rows_per_block;
  colred_v8_kernel<MODE><<<dim3((unsigned)blocks, (unsigned)slabs, 1), 256, 0, s>>>(p);
}

extern "C" {

int t2v_relu_fwd(const void* x, void* y, int64_t n, void* stream) {
  if (n % 8) return T2V_ERR_ARG;
  if (n == 0) return T2V_OK;
  relu_fwd_kernel<<<blocks_for(n / 8, 256), 256, 0, STREAM>>>(CBF(x), BF(y), n / 8);
  count_launch();
  return check_last("relu_fwd");
}
int t2v_relu_bwd(const void* dy, const void* ref, void* dx, int64_t n, void* stream) {
  if (n % 8) return T2V_ERR_ARG;
  if (n == 0) return T2V_OK;
  relu_bwd_kernel<<<blocks_for(n / 8, 256), 256, 0, STREAM>>>(CBF(dy), CBF(ref), BF(dx), n / 8);
  count_launch();
  return check_last("relu_bwd");
}

int t2v_leaky_relu_fwd(const void* x, void* y, int64_t n, float slope, void* stream) {
  if (n % 8) return T2V_ERR_ARG;
  if (n == 0) return T2V_OK;
  leaky_fwd_kernel<<<blocks_for(n / 8, 256), 256, 0, STREAM>>>(CBF(x), BF(y), slope, n / 8);
  count_launch();
  return check_last("leaky_relu_fwd");
}
int t2v_leaky_relu_bwd(const void* dy, const void* ref, void* dx, int64_t n, float slope, void* stream) {
  if (n % 8) return T2V_ERR_ARG;
  if (n == 0) return T2V_OK;
  leaky_bwd_kernel<<<blocks_for(n / 8, 256), 256, 0, STREAM>>>(CBF(dy), CBF(ref), BF(dx), slope, n / 8);
  count_launch();
  return check_last("leaky_relu_bwd");
}
int t2v_tanh_fwd(const void* x, void* y, int64_t n, void* stream) {
  if (n % 8) return T2V_ERR_ARG;
  if (n == 0) return T2V_OK;
  tanh_fwd_kernel<<<blocks_for(n / 8, 256), 256, 0, STREAM>>>(CBF(x), BF(y), n / 8);
  count_launch();
  return check_last("tanh_fwd");
}
int t2v_tanh_bwd(const void* dy, const void* y, void* dx, int64_t n, void* stream) {
  if (n % 8) return T2V_ERR_ARG;
  if (n == 0) return T2V_OK;
  tanh_bwd_kernel<<<blocks_for(n / 8, 256), 256, 0, STREAM>>>(CBF(dy), CBF(y), BF(dx), n / 8);
  count_launch();
  return check_last("tanh_bwd");
}

static int pool_params(PoolParams& p, const int32_t* shape, const int32_t* k, const int32_t* s, const int32_t* pad) {
  p.N = shape[0]; p.D = shape[1]; p.H = shape[2]; p.W = shape[3]; p.C = shape[4];
  p.kd = k[0]; p.kh = k[1]; p.kw = k[2]; p.sd = s[0]; p.sh = s[1]; p.sw = s[2];
  p.pd = pad[0]; p.ph = pad[1]; p.pw = pad[2];
  p.Do = (p.D + 2 * p.pd - p.kd) / p.sd + 1;
  p.Ho = (p.H + 2 * p.ph - p.kh) / p.sh + 1;
  p.Wo = (p.W + 2 * p.pw - p.kw) / p.sw + 1;
  p.inv = 1.f / (float)(p.kd * p.kh * p.kw);
  if (p.C % 8) return T2V_ERR_ARG;
  if (p.kd > p.sd || p.kh > p.sh || p.kw > p.sw) return T2V_ERR_ARG;
  return T2V_OK;
}
int t2v_avgpool_fwd(const void* x, const void* residual, void* y, const int32_t* in_shape, const int32_t* kernel,
                    const int32_t* stride, const int32_t* pad, void* stream) {
  PoolParams p;
  int rc = pool_params(p, in_shape, kernel, stride, pad);
  if (rc) return rc;
  const long long total8 = (long long)p.N * p.Do * p.Ho * p.Wo * p.C / 8;
  if (total8 == 0) return T2V_OK;
  const long long in_elems = (long long)p.N * p.D * p.H * p.W * p.C;
  if (in_elems < 0x7fffffffLL && total8 * 8 < 0x7fffffffLL)
    avgpool_fwd_kernel<unsigned><<<blocks_for(total8, 256), 256, 0, STREAM>>>(CBF(x), CBF(residual), BF(y), p, total8);
  else
    avgpool_fwd_kernel<long long><<<blocks_for(total8, 256), 256, 0, STREAM>>>(CBF(x), CBF(residual), BF(y), p, total8);
  count_launch();
  return check_last("avgpool_fwd");
}
int t2v_avgpool_bwd(const void* dy, void* dx, const int32_t* in_shape, const int32_t* kernel, const int32_t* stride,
                    const int32_t* pad, void* stream) {
  PoolParams p;
  int rc = pool_params(p, in_shape, kernel, stride, pad);
  if (rc) return rc;
  const long long total8 = (long long)p.N * p.D * p.H * p.W * p.C / 8;
  if (total8 == 0) return T2V_OK;
  if (total8 * 8 < 0x7fffffffLL)
    avgpool_bwd_kernel<unsigned><<<blocks_for(total8, 256), 256, 0, STREAM>>>(CBF(dy), BF(dx), p, total8);
  else
    avgpool_bwd_kernel<long long><<<blocks_for(total8, 256), 256, 0, STREAM>>>(CBF(dy), BF(dx), p, total8);
  count_launch();
  return check_last("avgpool_bwd");
}
int t2v_upsample2x_fwd(const void* x, void* y, int32_t N, int32_t H, int32_t W, int32_t C, void* stream) {
  if (C % 8) return T2V_ERR_ARG;
  const long long total8 = (long long)N * 4 * H * W * C / 8;
  if (total8 == 0) return T2V_OK;
  upsample2x_fwd_kernel<<<blocks_for(total8, 256), 256, 0, STREAM>>>(CBF(x), BF(y), N, H, W, C, total8);
  count_launch();
  return check_last("upsample2x_fwd");
}
int t2v_upsample2x_bwd(const void* dy, void* dx, int32_t N, int32_t H, int32_t W, int32_t C, void* stream) {
  if (C % 8) return T2V_ERR_ARG;
  const long long total8 = (long long)N * H * W * C / 8;
  if (total8 == 0) return T2V_OK;
  upsample2x_bwd_kernel<<<blocks_for(total8, 256), 256, 0, STREAM>>>(CBF(dy), BF(dx), N, H, W, C, total8);
  count_launch();
  return check_last("upsample2x_bwd");
}
int t2v_nchw_to_cl(const float* x, void* y, int64_t N, int32_t C, int64_t S, int32_t Cp, void* stream) {
  const long long total = N * S * Cp;
  if (total == 0) return T2V_OK;
  nchw_to_cl_kernel<<<blocks_for(total, 256), 256, 0, STREAM>>>(x, BF(y), C, Cp, S, total);
  count_launch();
  return check_last("nchw_to_cl");
}
int t2v_rgb_to_cl(const float* x, void* y16, void* y4, int64_t N, int64_t S, void* stream) {
  const long long total = N * S;
  if (total == 0 || (!y16 && !y4)) return T2V_OK;
  rgb_to_cl_kernel<<<blocks_for(total, 256), 256, 0, STREAM>>>(x, reinterpret_cast<uint4*>(y16),
                                                               reinterpret_cast<uint2*>(y4), S, total);
  count_launch();
  return check_last("rgb_to_cl");
}
int t2v_cl_to_nchw(const void* x, float* y, int64_t N, int32_t C, int64_t S, int32_t Cp, void* stream) {
  const long long total = N * S * C;
  if (total == 0) return T2V_OK;
  cl_to_nchw_kernel<<<blocks_for(total, 256), 256, 0, STREAM>>>(CBF(x), y, C, Cp, S, total);
  count_launch();
  return check_last("cl_to_nchw");
}
int t2v_im2col3(const float* x, void* col, int64_t N, int32_t C, int32_t D, int32_t H, int32_t W, int32_t Kp,
                void* stream) {
  if (C < 1 || C > 4 || Kp % 8 || Kp < 27 * C) return T2V_ERR_ARG;
  const long long total = N * D * H * W;
  if (total == 0) return T2V_OK;
  const unsigned nb = blocks_for(total, 128);
  switch (C) {
    case 1: im2col3_kernel<1><<<nb, 128, 0, STREAM>>>(x, BF(col), D, H, W, Kp, total); break;
    case 2: im2col3_kernel<2><<<nb, 128, 0, STREAM>>>(x, BF(col), D, H, W, Kp, total); break;
    case 3: im2col3_kernel<3><<<nb, 128, 0, STREAM>>>(x, BF(col), D, H, W, Kp, total); break;
    default: im2col3_kernel<4><<<nb, 128, 0, STREAM>>>(x, BF(col), D, H, W, Kp, total); break;
  }
  count_launch();
  return check_last("im2col3");
}
int t2v_col2im3(const void* dcol, float* dx, int64_t N, int32_t C, int32_t D, int32_t H, int32_t W, int32_t Kp,
                void* stream) {
  if (C < 1 || C > 4 || Kp < 27 * C) return T2V_ERR_ARG;
  const long long total = N * D * H * W;
  if (total == 0) return T2V_OK;
  const unsigned nb = blocks_for(total, 128);
  switch (C) {
    case 1: col2im3_kernel<1><<<nb, 128, 0, STREAM>>>(CBF(dcol), dx, D, H, W, Kp, total); break;
    case 2: col2im3_kernel<2><<<nb, 128, 0, STREAM>>>(CBF(dcol), dx, D, H, W, Kp, total); break;
    case 3: col2im3_kernel<3><<<nb, 128, 0, STREAM>>>(CBF(dcol), dx, D, H, W, Kp, total); break;
    default: col2im3_kernel<4><<<nb, 128, 0, STREAM>>>(CBF(dcol), dx, D, H, W, Kp, total); break;
  }
  count_launch();
  return check_last("col2im3");
}
static void row_split(long long P, int C, dim3* grid, long long* rows_per_block) {
  const int cg = (C + 63) / 64;
  long long slices = (4 * 148 + cg - 1) / cg;
  if (slices > (P + 63) / 64) slices = (P + 63) / 64;
  if (slices < 1) slices = 1;
  *rows_per_block = (P + slices - 1) / slices;
  *grid = dim3(cg, (unsigned)slices, 1);
}
static int sum_rows_impl(const void* x, float* out, int64_t P, int32_t C, int accumulate, void* stream);
int t2v_sum_rows(const void* x, float* out, int64_t P, int32_t C, void* stream) {
  return sum_rows_impl(x, out, P, C, 0, stream);
}
int t2v_sum_rows_acc(const void* x, float* out, int64_t P, int32_t C, void* stream) {
  return sum_rows_impl(x, out, P, C, 1, stream);
}
static int sum_rows_impl(const void* x, float* out, int64_t P, int32_t C, int accumulate, void* stream) {
  if (!accumulate) cudaMemsetAsync(out, 0, sizeof(float) * C, STREAM);
  if (P == 0) return T2V_OK;
  if (colred_ok(C)) {
    ColRed cr{};
    cr.P = P; cr.C = C; cr.x = CBF(x); cr.out = out;
    colred_launch<0>(cr, STREAM);
    count_launch();
    return check_last("sum_rows");
  }
  dim3 grid;
  long long rpb;
  row_split(P, C, &grid, &rpb);
  sum_rows_kernel<<<grid, 256, 0, STREAM>>>(CBF(x), out, P, C, rpb);
  count_launch();
  return check_last("sum_rows");
}
int t2v_sum_spatial(const void* x, float* out, int64_t N, int64_t S, int32_t C, void* stream) {
  if (N == 0) return T2V_OK;
  dim3 grid((C + 63) / 64, (unsigned)N, 1);
  sum_spatial_kernel<<<grid, 256, 0, STREAM>>>(CBF(x), out, S, C);
  count_launch();
  return check_last("sum_spatial");
}
int t2v_broadcast_spatial(const float* g, void* y, int64_t N, int64_t S, int32_t C, void* stream) {
  const long long total = N * S * C;
  if (total == 0) return T2V_OK;
  broadcast_spatial_kernel<<<blocks_for(total, 256), 256, 0, STREAM>>>(g, BF(y), S, C, total);
  count_launch();
  return check_last("broadcast_spatial");
}
int t2v_bn_stats(const void* x, float* stats, int64_t P, int32_t C, void* stream) {
  cudaMemsetAsync(stats, 0, sizeof(float) * 2 * C, STREAM);
  if (colred_ok(C) && P > 0) {
    ColRed cr{};
    cr.P = P; cr.C = C; cr.x = CBF(x); cr.out = stats;
    colred_launch<1>(cr, STREAM);
    count_launch();
    return check_last("bn_stats");
  }
  dim3 grid;
  long long rpb;
  row_split(P, C, &grid, &rpb);
  bn_stats_kernel<<<grid, 256, 0, STREAM>>>(CBF(x), stats, P, C, rpb);
  count_launch();
  return check_last("bn_stats");
}
int t2v_bn_finalize(const float* stats, const float* gamma, const float* beta, float* running_mean,
                    float* running_var, float* mean_invstd, float* scale_shift, int32_t C, int64_t count,
                    float eps, float momentum, void* stream) {
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, STREAM>>>(stats, gamma, beta, running_mean, running_var, mean_invstd,
                                                          scale_shift, C, (float)count, eps, momentum);
  count_launch();
  return check_last("bn_finalize");
}
int t2v_bn_apply(const void* x, const float* scale_shift, void* y, int64_t N, int32_t H, int32_t W, int32_t C,
                 int32_t relu, int32_t up, void* stream) {
  if (C % 8 || (up != 1 && up != 2)) return T2V_ERR_ARG;
  const long long total8 = N * H * W * up * up * C / 8;
  if (total8 == 0) return T2V_OK;
  bn_apply_kernel<<<blocks_for(total8, 256), 256, 0, STREAM>>>(CBF(x), scale_shift, BF(y), H, W, C, relu, up, total8);
  count_launch();
  return check_last("bn_apply");
}
int t2v_bn_bwd(const void* dy, const void* x, const float* scale_shift, const float* mean_invstd, float* red,
               void* dx, int64_t N, int32_t H, int32_t W, int32_t C, int32_t relu, int32_t up, void* stream) {
  if (C % 8 || (up != 1 && up != 2)) return T2V_ERR_ARG;
  const long long P = N * H * W;
  cudaMemsetAsync(red, 0, sizeof(float) * 2 * C, STREAM);
  if (P == 0) return T2V_OK;
  if (colred_ok(C)) {
    ColRed cr{};
    cr.P = P; cr.C = C; cr.H = H; cr.W = W; cr.act = relu; cr.up = up;
    cr.x = CBF(x); cr.dy = CBF(dy); cr.scale_shift = scale_shift; cr.mean_invstd = mean_invstd; cr.out = red;
    colred_launch<2>(cr, STREAM);
  } else {
    dim3 grid;
    long long rpb;
    row_split(P, C, &grid, &rpb);
    bn_bwd_reduce_kernel<<<grid, 256, 0, STREAM>>>(CBF(dy), CBF(x), scale_shift, mean_invstd, red, P, H, W, C, relu,
                                                   up, rpb);
  }
  const long long total8 = P * C / 8;
  bn_bwd_apply_kernel<<<blocks_for(total8, 256), 256, 0, STREAM>>>(CBF(dy), CBF(x), scale_shift, mean_invstd, red,
                                                                   BF(dx), P, H, W, C, relu, up, total8);
  count_launch(2);
  return check_last("bn_bwd");
}
int t2v_render_fwd(const void* pre, float* y, int32_t B, int32_t T, int32_t H, int32_t W, int32_t C, int32_t Cp,
                   void* stream) {
  const long long total = (long long)B * C * T * H * W;
  if (total == 0) return T2V_OK;
  render_fwd_kernel<<<blocks_for(total, 256), 256, 0, STREAM>>>(CBF(pre), y, B, T, H, W, C, Cp, total);
  count_launch();
  return check_last("render_fwd");
}
int t2v_render_bwd(const float* dy, const float* y, void* dpre, int32_t B, int32_t T, int32_t H, int32_t W,
                   int32_t C, int32_t Cp, void* stream) {
  const long long total = (long long)B * T * H * W * Cp;
  if (total == 0) return T2V_OK;
  render_bwd_kernel<<<blocks_for(total, 256), 256, 0, STREAM>>>(dy, y, BF(dpre), B, T, H, W, C, Cp, total);
  count_launch();
  return check_last("render_bwd");
}
int t2v_gather_frames(const void* x, void* y, int32_t B, int32_t T, int64_t frame_bytes, int32_t sn, int32_t st,
                      int32_t bt, const int32_t* bt_dev, int32_t scatter, void* stream) {
  if (frame_bytes % 16 || bt < 0) return T2V_ERR_ARG;
  const int Bo = (B + sn - 1) / sn;
  const int To = T > bt ? (T - bt + st - 1) / st : 0;
  const long long fu4 = frame_bytes / 16;
  if (scatter) cudaMemsetAsync(y, 0, (size_t)B * T * frame_bytes, STREAM);
  const long long total = (long long)Bo * To * fu4;
  if (total == 0) return T2V_OK;
  gather_frames_kernel<<<blocks_for(total, 256), 256, 0, STREAM>>>(reinterpret_cast<const uint4*>(x),
                                                                   reinterpret_cast<uint4*>(y), T, To, sn, st, bt, bt_dev,
                                                                   fu4, total, scatter);
  count_launch();
  return check_last("gather_frames");
}
int t2v_pyramid_level(const float* x, float* y, const int32_t* in_shape, int32_t Ho, int32_t Wo, int32_t sn,
                      int32_t st, int32_t bt, const int32_t* bt_dev, void* stream) {
  const int B = in_shape[0], C = in_shape[1], T = in_shape[2], H = in_shape[3], W = in_shape[4];
  if (sn < 1 || st < 1 || bt < 0) return T2V_ERR_ARG;
  const int Bo = (B + sn - 1) / sn;
  const int To = T > bt ? (T - bt + st - 1) / st : 0;
  const long long total = (long long)Bo * C * To * Ho * Wo;
  if (total == 0) return T2V_OK;
  pyramid_kernel<<<blocks_for(total, 256), 256, 0, STREAM>>>(x, y, B, C, T, H, W, Bo, To, Ho, Wo, sn, st, bt, bt_dev,
                                                             total);
  count_launch();
  return check_last("pyramid_level");
}
int t2v_lstm_cell_fwd(const float* gates, const float* c_prev, float* c, void* h, float* h32, int64_t P, int32_t Hd,
                      void* stream) {
  const long long total = P * Hd;
  if (total == 0) return T2V_OK;
  lstm_cell_fwd_kernel<<<blocks_for(total, 256), 256, 0, STREAM>>>(gates, c_prev, c, BF(h), h32, Hd, total);
  count_launch();
  return check_last("lstm_cell_fwd");
}
int t2v_lstm_cell_bwd(const float* gates, const float* c_prev, const float* c, const float* dh, const float* dc_next,
                      void* dgates, float* dc_prev, int64_t P, int32_t Hd, void* stream) {
  const long long total = P * Hd;
  if (total == 0) return T2V_OK;
  lstm_cell_bwd_kernel<<<blocks_for(total, 256), 256, 0, STREAM>>>(gates, c_prev, c, dh, dc_next, BF(dgates), dc_prev,
                                                                   Hd, total);
  count_launch();
  return check_last("lstm_cell_bwd");
}
int t2v_adam_step(int32_t count, float* const* host_params, const float* const* host_grads, float* const* host_m,
                  float* const* host_v, const int64_t* host_sizes, float lr, float beta1, float beta2, float eps,
                  int32_t step, float grad_scale, const float* dyn_dev, void* stream) {
  const float bc1 = 1.f - powf(beta1, (float)step), bc2 = 1.f - powf(beta2, (float)step);
  for (int base = 0; base < count; base += AdamChunk::kMax) {
    AdamChunk ch;
    ch.count = count - base < AdamChunk::kMax ? count - base : AdamChunk::kMax;
    long long maxn = 0;
    for (int i = 0; i < ch.count; ++i) {
      ch.p[i] = host_params[base + i];
      ch.g[i] = host_grads[base + i];
      ch.m[i] = host_m[base + i];
      ch.v[i] = host_v[base + i];
      ch.n[i] = host_sizes[base + i];
      if (ch.n[i] > maxn) maxn = ch.n[i];
    }
    long long bx = (maxn + 256 * 4 - 1) / (256 * 4);
    if (bx > 2048) bx = 2048;
    if (bx < 1) bx = 1;
    dim3 grid((unsigned)bx, ch.count, 1);
    adam_kernel<<<grid, 256, 0, STREAM>>>(ch, lr, beta1, beta2, eps, bc1, bc2, grad_scale, dyn_dev);
    count_launch();
  }
  return check_last("adam_step");
}

int t2v_stream_copy(const void* src, void* dst, int64_t nbytes, int32_t ctas, void* stream) {
  if (!src || !dst || nbytes < 0 || nbytes % 16 || ctas < 1) return T2V_ERR_ARG;
  if (nbytes == 0) return T2V_OK;
  stream_copy_kernel<<<(unsigned)ctas, 256, 0, STREAM>>>(reinterpret_cast<const uint4*>(src),
                                                         reinterpret_cast<uint4*>(dst), nbytes / 16);
  count_launch();
  return check_last("stream_copy");
}

int t2v_multi_copy(int32_t count, const float* const* host_src, float* const* host_dst, const int64_t* host_sizes,
                   void* stream) {
  for (int base = 0; base < count; base += CopyChunk::kMax) {
    CopyChunk ch;
    ch.count = count - base < CopyChunk::kMax ? count - base : CopyChunk::kMax;
    long long maxn = 0;
    for (int i = 0; i < ch.count; ++i) {
      ch.src[i] = host_src[base + i];
      ch.dst[i] = host_dst[base + i];
      ch.n[i] = host_sizes[base + i];
      if (ch.n[i] > maxn) maxn = ch.n[i];
    }
    long long bx = (maxn + 256 * 4 - 1) / (256 * 4);
    if (bx > 1024) bx = 1024;
    if (bx < 1) bx = 1;
    multi_copy_kernel<<<dim3((unsigned)bx, ch.count, 1), 256, 0, STREAM>>>(ch);
    count_launch();
  }
  return check_last("multi_copy");
}

}  // extern "C"
