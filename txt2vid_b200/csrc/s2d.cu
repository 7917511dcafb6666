// Stride-2 convolutions and transposed convolutions of the TGAN / TCWYT families on the tcgen05 engine.
//
// A kernel-4 / stride-2 / padding-1 convolution (txt2vid/models/tcwyt/video_discrim.py:12-25,
// tcwyt/frame_discrim.py:9-21; the critics of configs 1 and 2) reads, for output o, the input samples 2o-1 .. 2o+2 per
// strided axis.  Group the input into blocks of two samples aligned at ODD positions, block b = (2b-1, 2b), b = 0..O
// ("shifted space-to-depth": O+1 blocks, the first and the last half padding): output o reads exactly the blocks o and
// o+1, i.e. the layer is a DENSE kernel-2 stride-1 convolution over the block tensor with 2^s * Cin channels -- no
// zero taps, no im2col.  The engine runs it as a windowed implicit GEMM (t2v_conv_fprop_win: output extents O, input
// extents O+1, live taps {1,2} of a 3-tap axis).  The transposed convolutions of the generators
// (tgan/gen.py:20-23, tgan/temporal_gen.py:112-115, tcwyt/gen.py:18-26) are the data gradient of that convolution:
// the same GEMM with the transposed pack (taps {0,1}, output extents O+1) followed by the inverse block permutation.
//
//   t2v_s2d_shift         x (N,D,H,W,C)           -> xs (N,D',H',W',Cp)   block tensor, channel = phase*creal + c
//   t2v_d2s_shift         xs                      -> x                    inverse (pad channels of x zeroed)
//   t2v_s2d_embed_weight  w [Co][k taps][Ci]      -> [Co][3^s taps][Cp]   (fprop / wgrad operand order) or the
//                                                    transposed, tap-flipped [Cp][3^s taps][Co] (dgrad operand)
//   t2v_s2d_extract_wgrad dw_emb [Co][3^s][Cp]    -> dw [Co][k taps][Ci]
// Per axis: mode 2 = kernel 4 stride 2 padding 1 (even extent), mode 1 = kernel 1 stride 1 padding 0.
// Kernel index k4 = 2j + phase for block offset j in {0,1} (engine tap t = j + 1).
#include "t2v_common.cuh"

namespace t2v {

static inline unsigned blocks_for(long long n, int threads) { return (unsigned)((n + threads - 1) / threads); }

struct S2dParams {
  long long N;
  int D, H, W, C;          // the full-resolution tensor
  int Db, Hb, Wb, Cp;      // the block tensor
  int creal;               // real channels of x (<= C); block channel = phase * creal + c
  int md, mh, mw;          // 2: strided axis, 1: unit axis
  long long total;         // output units
};

// U = 16-byte vector (creal, C, Cp counted in vectors) or a scalar of the element size
template <typename U>
__global__ void s2d_shift_kernel(const U* __restrict__ x, U* __restrict__ xs, const S2dParams p) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.total) return;
  const int ch = (int)(i % p.Cp);
  long long t = i / p.Cp;
  const int bw = (int)(t % p.Wb); t /= p.Wb;
  const int bh = (int)(t % p.Hb); t /= p.Hb;
  const int bd = (int)(t % p.Db); t /= p.Db;
  const long long n = t;
  U v{};
  const int phase = ch / p.creal, c = ch - phase * p.creal;
  const int nph = (p.md == 2 ? 2 : 1) * (p.mh == 2 ? 2 : 1) * (p.mw == 2 ? 2 : 1);
  if (phase < nph) {
    int ph = phase;
    int d = bd, h = bh, w = bw;
    if (p.mw == 2) { w = 2 * bw - 1 + (ph & 1); ph >>= 1; }
    if (p.mh == 2) { h = 2 * bh - 1 + (ph & 1); ph >>= 1; }
    if (p.md == 2) { d = 2 * bd - 1 + (ph & 1); }
    if (d >= 0 && d < p.D && h >= 0 && h < p.H && w >= 0 && w < p.W)
      v = x[((((long long)n * p.D + d) * p.H + h) * p.W + w) * p.C + c];
  }
  xs[i] = v;
}

template <typename U>
__global__ void d2s_shift_kernel(const U* __restrict__ xs, U* __restrict__ x, const S2dParams p) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.total) return;
  const int c = (int)(i % p.C);
  long long t = i / p.C;
  const int w = (int)(t % p.W); t /= p.W;
  const int h = (int)(t % p.H); t /= p.H;
  const int d = (int)(t % p.D); t /= p.D;
  const long long n = t;
  U v{};
  if (c < p.creal) {
    int phase = 0, bd = d, bh = h, bw = w;
    if (p.md == 2) { bd = (d + 1) >> 1; phase = (d + 1) & 1; }
    if (p.mh == 2) { bh = (h + 1) >> 1; phase = phase * 2 + ((h + 1) & 1); }
    if (p.mw == 2) { bw = (w + 1) >> 1; phase = phase * 2 + ((w + 1) & 1); }
    v = xs[((((long long)n * p.Db + bd) * p.Hb + bh) * p.Wb + bw) * p.Cp + phase * p.creal + c];
  }
  x[i] = v;
}

struct S2dWeightParams {
  int Co, Ci, creal, Cp;      // Ci: channel stride of the 4-tap tensor (padded), creal <= Ci real channels
  int md, mh, mw;
  int transposed;
  long long total;
};

// decode an engine tap (3 per strided axis, 1 per unit axis) + phase into the kernel-4 tap index; false: dead entry
__device__ __forceinline__ bool s2d_k4_index(const S2dWeightParams& p, int te, int phase, int* k4) {
  const int ew = p.mw == 2 ? 3 : 1, eh = p.mh == 2 ? 3 : 1;
  const int kw = p.mw == 2 ? 4 : 1, kh = p.mh == 2 ? 4 : 1;
  const int tw = te % ew, th = (te / ew) % eh, td = te / (ew * eh);
  int ph = phase, a_w = 0, a_h = 0, a_d = 0;
  if (p.mw == 2) { if (tw == 0) return false; a_w = 2 * (tw - 1) + (ph & 1); ph >>= 1; }
  if (p.mh == 2) { if (th == 0) return false; a_h = 2 * (th - 1) + (ph & 1); ph >>= 1; }
  if (p.md == 2) { if (td == 0) return false; a_d = 2 * (td - 1) + (ph & 1); ph >>= 1; }
  if (ph != 0) return false;
  *k4 = (a_d * kh + a_h) * kw + a_w;
  return true;
}

// one thread per element of the embedded pack
__global__ void s2d_embed_weight_kernel(const __nv_bfloat16* __restrict__ w, __nv_bfloat16* __restrict__ we,
                                        const S2dWeightParams p) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.total) return;
  const int ntap = (p.md == 2 ? 3 : 1) * (p.mh == 2 ? 3 : 1) * (p.mw == 2 ? 3 : 1);
  const int k4taps = (p.md == 2 ? 4 : 1) * (p.mh == 2 ? 4 : 1) * (p.mw == 2 ? 4 : 1);
  int co, te, ch;
  if (!p.transposed) {                        // [Co][taps][Cp]
    ch = (int)(i % p.Cp); te = (int)((i / p.Cp) % ntap); co = (int)(i / ((long long)p.Cp * ntap));
  } else {                                    // [Cp][taps flipped][Co]
    co = (int)(i % p.Co); te = ntap - 1 - (int)((i / p.Co) % ntap); ch = (int)(i / ((long long)p.Co * ntap));
  }
  const int phase = ch / p.creal, c = ch - phase * p.creal;
  int k4;
  __nv_bfloat16 v = __float2bfloat16(0.f);
  if (s2d_k4_index(p, te, phase, &k4)) v = w[((long long)co * k4taps + k4) * p.Ci + c];
  we[i] = v;
}

// one thread per element of dw [Co][k4 taps][Ci]
__global__ void s2d_extract_wgrad_kernel(const float* __restrict__ dwe, float* __restrict__ dw, const S2dWeightParams p) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.total) return;
  const int ntap = (p.md == 2 ? 3 : 1) * (p.mh == 2 ? 3 : 1) * (p.mw == 2 ? 3 : 1);
  const int kw = p.mw == 2 ? 4 : 1, kh = p.mh == 2 ? 4 : 1, kd = p.md == 2 ? 4 : 1;
  const int c = (int)(i % p.Ci);
  const int k4 = (int)((i / p.Ci) % (kd * kh * kw));
  const int co = (int)(i / ((long long)p.Ci * kd * kh * kw));
  float v = 0.f;
  if (c < p.creal) {
    const int a_w = k4 % kw, a_h = (k4 / kw) % kh, a_d = k4 / (kw * kh);
    int phase = 0, td = 0, th = 0, tw = 0;
    if (p.md == 2) { td = a_d / 2 + 1; phase = a_d & 1; }
    if (p.mh == 2) { th = a_h / 2 + 1; phase = phase * 2 + (a_h & 1); }
    if (p.mw == 2) { tw = a_w / 2 + 1; phase = phase * 2 + (a_w & 1); }
    const int ew = p.mw == 2 ? 3 : 1, eh = p.mh == 2 ? 3 : 1;
    const int te = (td * eh + th) * ew + tw;
    v = dwe[((long long)co * ntap + te) * p.Cp + phase * p.creal + c];
  }
  dw[i] = v;
}

// wT[ci][taps - 1 - t][co] = w[co][t][ci]: the data-gradient / transposed-convolution operand of a stride-1 layer
__global__ void transpose_flip_kernel(const __nv_bfloat16* __restrict__ w, __nv_bfloat16* __restrict__ wT, int Co,
                                      int taps, int Ci, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int co = (int)(i % Co);
  const int t = taps - 1 - (int)((i / Co) % taps);
  const int ci = (int)(i / ((long long)Co * taps));
  wT[i] = w[((long long)co * taps + t) * Ci + ci];
}

// corner window x[:, :kd, :kh, :kw, :] <-> contiguous rows [N][kd*kh*kw*C] (units of 16 bytes); scatter = 1 writes the
// rows back into x (the rest of x was zeroed by the caller): a convolution whose single output position reads that
// window is a Linear layer over the gathered rows (tcwyt/frame_discrim.py:55, motion_discrim.py:19, video_discrim.py:46)
__global__ void window_rows_kernel(uint4* __restrict__ x, uint4* __restrict__ rows, int D, int H, int W, int C16, int kd,
                                   int kh, int kw, int scatter, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C16);
  long long t = i / C16;
  const int a_w = (int)(t % kw); t /= kw;
  const int a_h = (int)(t % kh); t /= kh;
  const int a_d = (int)(t % kd); t /= kd;
  const long long n = t;
  const long long xi = ((((long long)n * D + a_d) * H + a_h) * W + a_w) * C16 + c;
  if (scatter) x[xi] = rows[i];
  else rows[i] = x[xi];
}

__global__ void s2d_tile_bias_kernel(const float* __restrict__ bias, float* __restrict__ out, int creal, int phases,
                                     int Cp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Cp) return;
  out[i] = i < phases * creal ? bias[i % creal] : 0.f;
}

static bool s2d_fill(S2dParams* p, int64_t N, int D, int H, int W, int C, int creal, int md, int mh, int mw, int Cp) {
  auto okm = [](int m, int e) { return m == 1 || (m == 2 && e % 2 == 0); };
  if (N <= 0 || D <= 0 || H <= 0 || W <= 0 || C <= 0 || creal <= 0 || creal > C) return false;
  if (!okm(md, D) || !okm(mh, H) || !okm(mw, W)) return false;
  const int nph = (md == 2 ? 2 : 1) * (mh == 2 ? 2 : 1) * (mw == 2 ? 2 : 1);
  if (Cp < nph * creal) return false;
  p->N = N; p->D = D; p->H = H; p->W = W; p->C = C; p->creal = creal; p->md = md; p->mh = mh; p->mw = mw;
  p->Db = md == 2 ? D / 2 + 1 : D; p->Hb = mh == 2 ? H / 2 + 1 : H; p->Wb = mw == 2 ? W / 2 + 1 : W;
  p->Cp = Cp;
  return true;
}

// run `kern<U>` with the widest unit the channel counts allow
template <typename F16, typename F8, typename F4, typename F2>
static int s2d_dispatch(S2dParams p, int elem_bytes, bool to_blocks, cudaStream_t s, F16 k16, F8 k8, F4 k4, F2 k2) {
  int unit = 16;
  while (unit > elem_bytes && ((p.C * elem_bytes) % unit || (p.Cp * elem_bytes) % unit || (p.creal * elem_bytes) % unit))
    unit >>= 1;
  const int per = unit / elem_bytes;
  p.C /= per; p.Cp /= per; p.creal /= per;
  p.total = to_blocks ? p.N * p.Db * p.Hb * p.Wb * p.Cp : p.N * p.D * p.H * p.W * p.C;
  if (p.total == 0) return T2V_OK;
  const unsigned blocks = blocks_for(p.total, 256);
  if (unit == 16) k16(blocks, p, s);
  else if (unit == 8) k8(blocks, p, s);
  else if (unit == 4) k4(blocks, p, s);
  else k2(blocks, p, s);
  count_launch();
  return T2V_OK;
}

}  // namespace t2v

using namespace t2v;

extern "C" {

int t2v_s2d_shift(const void* x, void* xs, int64_t N, int32_t D, int32_t H, int32_t W, int32_t C, int32_t creal,
                  int32_t md, int32_t mh, int32_t mw, int32_t Cp, int32_t elem_bytes, void* stream) {
  S2dParams p{};
  if (!x || !xs || (elem_bytes != 2 && elem_bytes != 4) || !s2d_fill(&p, N, D, H, W, C, creal, md, mh, mw, Cp))
    return T2V_ERR_ARG;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  auto run = [&](auto tag) {
    using U = decltype(tag);
    return [=](unsigned blocks, const S2dParams& q, cudaStream_t st) {
      s2d_shift_kernel<U><<<blocks, 256, 0, st>>>(reinterpret_cast<const U*>(x), reinterpret_cast<U*>(xs), q);
    };
  };
  const int rc = s2d_dispatch(p, elem_bytes, true, s, run(uint4{}), run(uint2{}), run(uint32_t{}), run(uint16_t{}));
  return rc ? rc : check_last("s2d_shift");
}

int t2v_d2s_shift(const void* xs, void* x, int64_t N, int32_t D, int32_t H, int32_t W, int32_t C, int32_t creal,
                  int32_t md, int32_t mh, int32_t mw, int32_t Cp, int32_t elem_bytes, void* stream) {
  S2dParams p{};
  if (!x || !xs || (elem_bytes != 2 && elem_bytes != 4) || !s2d_fill(&p, N, D, H, W, C, creal, md, mh, mw, Cp))
    return T2V_ERR_ARG;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  auto run = [&](auto tag) {
    using U = decltype(tag);
    return [=](unsigned blocks, const S2dParams& q, cudaStream_t st) {
      d2s_shift_kernel<U><<<blocks, 256, 0, st>>>(reinterpret_cast<const U*>(xs), reinterpret_cast<U*>(x), q);
    };
  };
  const int rc = s2d_dispatch(p, elem_bytes, false, s, run(uint4{}), run(uint2{}), run(uint32_t{}), run(uint16_t{}));
  return rc ? rc : check_last("d2s_shift");
}

int t2v_s2d_embed_weight(const void* w, void* we, int32_t Co, int32_t Ci, int32_t creal, int32_t Cp, int32_t md,
                         int32_t mh, int32_t mw, int32_t transposed, void* stream) {
  if (!w || !we || Co <= 0 || Ci <= 0 || creal <= 0 || creal > Ci) return T2V_ERR_ARG;
  const int nph = (md == 2 ? 2 : 1) * (mh == 2 ? 2 : 1) * (mw == 2 ? 2 : 1);
  const int ntap = (md == 2 ? 3 : 1) * (mh == 2 ? 3 : 1) * (mw == 2 ? 3 : 1);
  if (Cp < nph * creal) return T2V_ERR_ARG;
  S2dWeightParams p{Co, Ci, creal, Cp, md, mh, mw, transposed, (long long)Co * ntap * Cp};
  s2d_embed_weight_kernel<<<blocks_for(p.total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(w), reinterpret_cast<__nv_bfloat16*>(we), p);
  count_launch();
  return check_last("s2d_embed_weight");
}

int t2v_transpose_flip_bf16(const void* w, void* wT, int32_t Co, int32_t taps, int32_t Ci, void* stream) {
  if (!w || !wT || Co <= 0 || taps <= 0 || Ci <= 0) return T2V_ERR_ARG;
  const long long total = (long long)Co * taps * Ci;
  transpose_flip_kernel<<<blocks_for(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(w), reinterpret_cast<__nv_bfloat16*>(wT), Co, taps, Ci, total);
  count_launch();
  return check_last("transpose_flip_bf16");
}

int t2v_window_rows(void* x, void* rows, int64_t N, int32_t D, int32_t H, int32_t W, int32_t C, int32_t kd, int32_t kh,
                    int32_t kw, int32_t elem_bytes, int32_t scatter, void* stream) {
  if (!x || !rows || N <= 0 || kd <= 0 || kh <= 0 || kw <= 0 || kd > D || kh > H || kw > W ||
      (elem_bytes != 2 && elem_bytes != 4) || (C * elem_bytes) % 16)
    return T2V_ERR_ARG;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int C16 = C * elem_bytes / 16;
  if (scatter) cudaMemsetAsync(x, 0, (size_t)N * D * H * W * C * elem_bytes, s);
  const long long total = (long long)N * kd * kh * kw * C16;
  window_rows_kernel<<<blocks_for(total, 256), 256, 0, s>>>(reinterpret_cast<uint4*>(x), reinterpret_cast<uint4*>(rows),
                                                            D, H, W, C16, kd, kh, kw, scatter, total);
  count_launch();
  return check_last("window_rows");
}

int t2v_s2d_tile_bias(const float* bias, float* out, int32_t creal, int32_t phases, int32_t Cp, void* stream) {
  if (!bias || !out || creal <= 0 || phases <= 0 || Cp < creal * phases) return T2V_ERR_ARG;
  s2d_tile_bias_kernel<<<blocks_for(Cp, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(bias, out, creal, phases,
                                                                                                Cp);
  count_launch();
  return check_last("s2d_tile_bias");
}

int t2v_s2d_extract_wgrad(const float* dwe, float* dw, int32_t Co, int32_t Ci, int32_t creal, int32_t Cp, int32_t md,
                          int32_t mh, int32_t mw, void* stream) {
  if (!dwe || !dw || Co <= 0 || Ci <= 0 || creal <= 0 || creal > Ci) return T2V_ERR_ARG;
  const int nph = (md == 2 ? 2 : 1) * (mh == 2 ? 2 : 1) * (mw == 2 ? 2 : 1);
  const int k4taps = (md == 2 ? 4 : 1) * (mh == 2 ? 4 : 1) * (mw == 2 ? 4 : 1);
  if (Cp < nph * creal) return T2V_ERR_ARG;
  S2dWeightParams p{Co, Ci, creal, Cp, md, mh, mw, 0, (long long)Co * k4taps * Ci};
  s2d_extract_wgrad_kernel<<<blocks_for(p.total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(dwe, dw, p);
  count_launch();
  return check_last("s2d_extract_wgrad");
}

}  // extern "C"
