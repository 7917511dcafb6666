// On-device input pipeline: the synthetic moving-MNIST clips and their captions, and the
// ToTensor + Normalize transform of stored uint8 frames.
//
//   t2v_moving_digits   frames of txt2vid/data/synthetic/generate.py:18-47 (generate_frames): a black RGB frame with one
//                       28x28 grey patch pasted at a per-frame position (PIL paste of an 'L' image into 'RGB': the grey
//                       value goes to all three channels).  The positions are the host's (the reference computes them
//                       in float64 from Python / numpy draws; data.MovingDigits restates that and uploads B*T pairs);
//                       the kernel is a pure index kernel -> bit-exact frames, written either as stored (uint8) or
//                       already through transforms.ToTensor() + Normalize(0.5, 0.5) (data/__init__.py:362-364).
//   t2v_grammar_tokens  token rows of the captions "digit {c} is {a} and {b}." (generate.py:102-182) as
//                       Vocab.tokenize + collate_fn produce them (data/__init__.py:260-355): START digit c is a and b END.
//   t2v_u8_normalize    x / 255, then (x - 0.5) / 0.5 with IEEE division: the values torch's CPU transform produces.
#include "t2v_common.cuh"

namespace t2v {

static inline unsigned sy_blocks(long long n, int threads) { return (unsigned)((n + threads - 1) / threads); }

__device__ __forceinline__ float to_tensor_normalize(unsigned v) {
  return (__fdiv_rn((float)v, 255.f) - 0.5f) * 2.f;      // (x - 0.5) / 0.5 is exactly (x - 0.5) * 2
}

// one thread per (b, t, y, x); layout 0: (B, T, 3, H, W) loader order, 1: (B, 3, T, H, W) training order
template <typename OUT>
__global__ void moving_digits_kernel(const uint8_t* __restrict__ bank, const int* __restrict__ digit,
                                     const int* __restrict__ pos, OUT* __restrict__ out, int B, int T, int H, int W,
                                     int oh, int ow, int layout, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int x = (int)(i % W);
  const int y = (int)((i / W) % H);
  const int t = (int)((i / ((long long)W * H)) % T);
  const int b = (int)(i / ((long long)W * H * T));
  const int px = pos[((long long)b * T + t) * 2], py = pos[((long long)b * T + t) * 2 + 1];
  unsigned v = 0;
  const int dx = x - px, dy = y - py;
  if (dx >= 0 && dx < ow && dy >= 0 && dy < oh) v = bank[((long long)digit[b] * oh + dy) * ow + dx];
  OUT o;
  if (sizeof(OUT) == 1) o = (OUT)v;
  else o = (OUT)to_tensor_normalize(v);
  const long long plane = (long long)H * W;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const long long idx = layout == 0 ? (((long long)b * T + t) * 3 + c) * plane + (long long)y * W + x
                                      : (((long long)b * 3 + c) * T + t) * plane + (long long)y * W + x;
    out[idx] = o;
  }
}

// table: [0] START, [1] "digit", [2..11] "0".."9", [12] "is", [13] "and", [14] END, [15 + 2 m], [16 + 2 m]: the two
// direction words of move m (0 bottom/top, 1 top/bottom, 2 right/left, 3 left/right = horizontal * 2 + l2r)
__global__ void grammar_tokens_kernel(const int* __restrict__ cls, const int* __restrict__ move,
                                      const long long* __restrict__ table, long long* __restrict__ tokens, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  long long* row = tokens + (long long)b * 8;
  const int m = move[b];
  row[0] = table[0]; row[1] = table[1]; row[2] = table[2 + cls[b]]; row[3] = table[12];
  row[4] = table[15 + 2 * m]; row[5] = table[13]; row[6] = table[16 + 2 * m]; row[7] = table[14];
}

__global__ void u8_normalize_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, long long n4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const uchar4 v = reinterpret_cast<const uchar4*>(src)[i];
  reinterpret_cast<float4*>(dst)[i] = make_float4(to_tensor_normalize(v.x), to_tensor_normalize(v.y),
                                                  to_tensor_normalize(v.z), to_tensor_normalize(v.w));
}
__global__ void u8_normalize_tail_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, long long from,
                                         long long n) {
  const long long i = from + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = to_tensor_normalize(src[i]);
}

}  // namespace t2v

using namespace t2v;

extern "C" {

int t2v_moving_digits(const void* bank, const int32_t* digit, const int32_t* pos, void* out, int32_t B, int32_t T,
                      int32_t H, int32_t W, int32_t oh, int32_t ow, int32_t out_f32, int32_t layout, void* stream) {
  if (!bank || !digit || !pos || !out || B <= 0 || T <= 0 || H <= 0 || W <= 0 || oh <= 0 || ow <= 0 ||
      (layout != 0 && layout != 1))
    return T2V_ERR_ARG;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const long long total = (long long)B * T * H * W;
  if (out_f32)
    moving_digits_kernel<float><<<sy_blocks(total, 256), 256, 0, s>>>(reinterpret_cast<const uint8_t*>(bank), digit, pos,
                                                                      reinterpret_cast<float*>(out), B, T, H, W, oh, ow,
                                                                      layout, total);
  else
    moving_digits_kernel<uint8_t><<<sy_blocks(total, 256), 256, 0, s>>>(reinterpret_cast<const uint8_t*>(bank), digit,
                                                                        pos, reinterpret_cast<uint8_t*>(out), B, T, H, W,
                                                                        oh, ow, layout, total);
  count_launch();
  return check_last("moving_digits");
}

int t2v_grammar_tokens(const int32_t* cls, const int32_t* move, const int64_t* table, int64_t* tokens, int32_t B,
                       void* stream) {
  if (!cls || !move || !table || !tokens || B <= 0) return T2V_ERR_ARG;
  grammar_tokens_kernel<<<sy_blocks(B, 128), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      cls, move, reinterpret_cast<const long long*>(table), reinterpret_cast<long long*>(tokens), B);
  count_launch();
  return check_last("grammar_tokens");
}

int t2v_u8_normalize(const void* src, float* dst, int64_t n, void* stream) {
  if (!src || !dst || n < 0) return T2V_ERR_ARG;
  if (n == 0) return T2V_OK;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const bool aligned = (reinterpret_cast<uintptr_t>(src) % 4 == 0) && (reinterpret_cast<uintptr_t>(dst) % 16 == 0);
  const long long n4 = aligned ? n / 4 : 0;
  if (n4 > 0)
    u8_normalize_kernel<<<sy_blocks(n4, 256), 256, 0, s>>>(reinterpret_cast<const uint8_t*>(src), dst, n4);
  if (n4 * 4 < n)
    u8_normalize_tail_kernel<<<sy_blocks(n - n4 * 4, 256), 256, 0, s>>>(reinterpret_cast<const uint8_t*>(src), dst,
                                                                        n4 * 4, n);
  count_launch();
  return check_last("u8_normalize");
}

}  // extern "C"
