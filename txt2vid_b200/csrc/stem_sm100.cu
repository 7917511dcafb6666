// Direct tensor-core kernels for the RGB stem convolution of the discriminator (sm_100a):
//     Conv3d(3 -> 64, kernel 3, padding 1) + bias + ReLU      (txt2vid/models/resnet3d.py:12-13)
//
// K = 27 taps x 3 channels = 81 is too thin for the implicit GEMM of igemm_sm100.cu (a tap would be a 6-byte row),
// and an im2col tensor in HBM costs 192 B per voxel to write and again to read -- more than the layer's input
// (32 B) and output (128 B) together.  Here the im2col tile exists only in shared memory: producer warps gather
// the 27 neighbours of each voxel from the bf16 channels-last clip (16 channels per voxel, RGB in the first four:
// one 8-byte load per tap) and store them as rows of the canonical 128B-swizzled operand layout, k = tap * 4 + c
// (K padded 108 -> 128); the tensor core reads what the threads wrote (fence.proxy.async in between).
//
//   fprop: D[128 voxels x 64 cout] = A[voxel, k] * W[cout, k]        A: the tile above (K-major), W: TMA, resident
//   wgrad: D[cout x 128 k]        += dy[voxel, cout]^T * A[voxel, k] A tile as MN-major B operand (same bytes),
//          dy boxes via TMA (MN-major A; the upper 64 rows of M = 128 read a zeroed region), positions split over
//          the CTAs, fp32 red.add of the 64 x 81 useful entries into dw[cout][tap][c].
// The data gradient (gradient-penalty and generator steps) stays on the im2col formulation (t2v_col2im3).
#include <cstdio>
#include <cstdlib>

#include "t2v_common.cuh"
#include "tmap.cuh"

namespace t2v {

struct StemParams {
  int N, D, H, W;
  long long P;                   // voxels
  const __nv_bfloat16* xc;       // (N, D, H, W, 1 << cshift) bf16, channels 0..2 = RGB, the rest zero
  int cshift;                    // 2 or 4: log2 of the channels per voxel
  const float* bias;
  __nv_bfloat16* y;              // (N, D, H, W, 64)
  int relu;
  uint32_t idesc;
  float* dw;                     // [64][27][3] fp32
  long long blocks_total;        // wgrad: 64-voxel blocks
  int per_cta;
};

static constexpr uint32_t kBlkA = 128u * 128u;   // one K block of the fprop A tile: 128 rows x 128 B
static constexpr uint32_t kBlkW = 64u * 128u;    // one K block of the weights / one 64-voxel block of the wgrad tiles

// Row r of a tile whose K blocks (64 k = 128 B per row) are `blk` bytes apart: tap t (4 bf16 = 8 bytes, k = 4 t) lives
// in block t / 16, 16-byte chunk ((t % 16) / 2) ^ (r % 8) (the 128B swizzle), half t % 2.
struct StemRow {
  uint32_t chunk[8];       // byte offset of logical chunk j of this row inside block 0
  int base;                // element offset of the voxel in xc (16 bf16 per voxel)
  uint32_t vmask;          // bit (a_d) | bit (3 + a_h) | bit (6 + a_w): that neighbour coordinate is inside the clip
};

__device__ __forceinline__ StemRow stem_row(const StemParams& p, int r, long long pos) {
  StemRow s;
  const uint32_t rowb = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u;
#pragma unroll
  for (int j = 0; j < 8; ++j) s.chunk[j] = rowb + ((uint32_t)(j ^ (r & 7)) << 4);
  const bool valid = pos < p.P;
  int q = valid ? (int)pos : 0;
  const int w = q % p.W; q /= p.W;
  const int h = q % p.H; q /= p.H;
  const int d = q % p.D;
  s.base = (valid ? (int)pos : 0) << p.cshift;
  uint32_t m = 0;
  if (valid) {
    m |= (d > 0 ? 1u : 0u) | 2u | (d + 1 < p.D ? 4u : 0u);
    m |= (h > 0 ? 8u : 0u) | 16u | (h + 1 < p.H ? 32u : 0u);
    m |= (w > 0 ? 64u : 0u) | 128u | (w + 1 < p.W ? 256u : 0u);
  }
  s.vmask = m;
  return s;
}

// the taps [T0, T1) of the row's voxel: loads (one predicate + one 8-byte load per tap; tap coordinates are compile-time)
// and stores into the tile (chunk index and K block are compile-time) are separate so that callers can overlap them
template <int T0, int T1>
__device__ __forceinline__ void stem_load(const StemParams& p, const StemRow& s, uint2* v) {
  const int sw = 1 << p.cshift, sh = p.W << p.cshift, sd = (p.H * p.W) << p.cshift;
#pragma unroll
  for (int tap = T0; tap < T1; ++tap) {
    const int a_w = tap % 3, a_h = (tap / 3) % 3, a_d = tap / 9;
    const uint32_t need = (1u << a_d) | (8u << a_h) | (64u << a_w);
    v[tap - T0] = make_uint2(0u, 0u);
    if ((s.vmask & need) == need)
      v[tap - T0] = *reinterpret_cast<const uint2*>(p.xc + s.base + (a_d - 1) * sd + (a_h - 1) * sh + (a_w - 1) * sw);
  }
}
template <int T0, int T1>
__device__ __forceinline__ void stem_store(uint8_t* tile, uint32_t blk, const StemRow& s, const uint2* v) {
#pragma unroll
  for (int tap = T0; tap < T1; ++tap) {
    const int b = tap >> 4, tt = tap & 15;
    *reinterpret_cast<uint2*>(tile + (uint32_t)b * blk + s.chunk[tt >> 1] + (uint32_t)(tt & 1) * 8u) = v[tap - T0];
  }
}
template <int T0, int T1>
__device__ __forceinline__ void stem_gather(const StemParams& p, uint8_t* tile, uint32_t blk, const StemRow& s) {
  uint2 v[T1 - T0];
  stem_load<T0, T1>(p, s, v);
  stem_store<T0, T1>(tile, blk, s, v);
}
__device__ __forceinline__ void stem_pad(uint8_t* tile, uint32_t blk, const StemRow& s) {
#pragma unroll
  for (int tap = 27; tap < 32; ++tap) {
    const int tt = tap & 15;
    *reinterpret_cast<uint2*>(tile + blk + s.chunk[tt >> 1] + (uint32_t)(tt & 1) * 8u) = make_uint2(0u, 0u);
  }
}

// ------------------------------------------------------------------------------------ fprop
// Persistent CTAs (two per SM): warps 0-3 gather the A tile of tile i+1 while warp 8 multiplies tile i and warps 4-7
// drain the accumulator of tile i-1 (two A buffers, two TMEM accumulators); the weights are loaded once per CTA.
static constexpr int kStemFpThreads = 288;

__global__ void __launch_bounds__(kStemFpThreads, 2)
stem_fprop_kernel(const __grid_constant__ CUtensorMap tmW, const StemParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;                       // 2 buffers x (2 K blocks x 16 KB)
  uint8_t* sW = smem + 4 * kBlkA;           // 2 x 8 KB
  uint64_t* w_full = reinterpret_cast<uint64_t*>(sW + 2 * kBlkW);
  uint64_t* a_full = w_full + 1;
  uint64_t* a_empty = a_full + 2;
  uint64_t* acc_full = a_empty + 2;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long tiles = (p.P + 127) / 128;
  const int iters = (int)((tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
  if (warp == 8) {
    if (lane == 0) {
      tma_prefetch_desc(&tmW);
      mbar_init(w_full, 1);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&a_full[i], 128);
        mbar_init(&a_empty[i], 1);
        mbar_init(&acc_full[i], 1);
        mbar_init(&acc_empty[i], 4);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 128);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    if (lane == 0) {
      mbar_expect_tx(w_full, 2 * kBlkW);
      tma_load_2d(sW, &tmW, w_full, 0, 0);
      tma_load_2d(sW + kBlkW, &tmW, w_full, 64, 0);
    }
    const uint32_t d_hi = desc_hi(1024, 2);
    const uint32_t leader = elect_one();
    mbar_wait(w_full, 0);
    const uint32_t w16 = smem_u32(sW) >> 4;
    for (int it = 0; it < iters; ++it) {
      const int buf = it & 1;
      const uint32_t par = (uint32_t)(it >> 1) & 1u;
      mbar_wait(&a_full[buf], par);
      mbar_wait(&acc_empty[buf], par ^ 1u);
      tc_fence_after();
      const uint32_t a16 = smem_u32(sA + (size_t)buf * 2 * kBlkA) >> 4;
#pragma unroll
      for (int b = 0; b < 2; ++b) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (leader)
            umma_bf16_ss2(tmem_base + (uint32_t)(buf * 64), desc_lo(a16 + (uint32_t)b * (kBlkA >> 4) + 2u * k, 0), d_hi,
                          desc_lo(w16 + (uint32_t)b * (kBlkW >> 4) + 2u * k, 0), d_hi, p.idesc, (b | k) != 0 ? 1u : 0u);
        }
      }
      if (leader) {
        umma_commit(&a_empty[buf]);
        umma_commit(&acc_full[buf]);
      }
      __syncwarp();
    }
  } else if (warp < 4) {
    const int r = threadIdx.x;
    for (int it = 0; it < iters; ++it) {
      const int buf = it & 1;
      const long long pos = ((long long)blockIdx.x + (long long)it * gridDim.x) * 128 + r;
      const StemRow row = stem_row(p, r, pos);
      uint2 v0[14], v1[13];
      stem_load<0, 14>(p, row, v0);              // loads are in flight while the buffer is still being multiplied
      stem_load<14, 27>(p, row, v1);
      mbar_wait(&a_empty[buf], ((uint32_t)(it >> 1) & 1u) ^ 1u);
      uint8_t* tile = sA + (size_t)buf * 2 * kBlkA;
      stem_store<0, 14>(tile, kBlkA, row, v0);
      stem_store<14, 27>(tile, kBlkA, row, v1);
      stem_pad(tile, kBlkA, row);
      fence_proxy_async();
      mbar_arrive(&a_full[buf]);
    }
  } else {
    // epilogue (warps 4-7): TMEM lane = row
    const int q = warp & 3;
    const int r = q * 32 + lane;
    for (int it = 0; it < iters; ++it) {
      const int buf = it & 1;
      const long long pos = ((long long)blockIdx.x + (long long)it * gridDim.x) * 128 + r;
      mbar_wait(&acc_full[buf], (uint32_t)(it >> 1) & 1u);
      tc_fence_after();
      const bool ok = pos < p.P;
#pragma unroll
      for (int c = 0; c < 64; c += 16) {
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 64 + c), v);
        tmem_ld_wait();
        if (ok) {
          float f[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
          if (p.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 b = *reinterpret_cast<const float4*>(p.bias + c + j);
              f[j] += b.x; f[j + 1] += b.y; f[j + 2] += b.z; f[j + 3] += b.w;
            }
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
          }
          st_global_v8(p.y + pos * 64 + c,
                       make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7])),
                       make_uint4(pack_bf16x2(f[8], f[9]), pack_bf16x2(f[10], f[11]), pack_bf16x2(f[12], f[13]), pack_bf16x2(f[14], f[15])));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

// ------------------------------------------------------------------------------------ wgrad
static constexpr int kStemWgThreads = 192;   // warps 0-3: gather + epilogue, warp 4: MMA, warp 5: dy TMA
static constexpr int kStemWgStages = 3;
static constexpr uint32_t kWgStage = 4 * kBlkW;   // dy box 8 KB | zero 8 KB | col tile 2 x 8 KB

__global__ void __launch_bounds__(kStemWgThreads, 2)
stem_wgrad_kernel(const __grid_constant__ CUtensorMap tmDy, const StemParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kStemWgStages * kWgStage);
  uint64_t* empty = full + kStemWgStages;
  uint64_t* accum = empty + kStemWgStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long b0 = (long long)blockIdx.x * p.per_cta;
  const long long b1 = b0 + p.per_cta < p.blocks_total ? b0 + p.per_cta : p.blocks_total;
  const int nb = b1 > b0 ? (int)(b1 - b0) : 0;

  if (warp == 4) {
    if (lane == 0) {
      tma_prefetch_desc(&tmDy);
      for (int s = 0; s < kStemWgStages; ++s) {
        mbar_init(&full[s], 129);      // 128 gather threads + the TMA thread's expect_tx arrive
        mbar_init(&empty[s], 1);
      }
      mbar_init(accum, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 128);
    tmem_relinquish();
  }
  // rows 64..127 of the M = 128 operand: a zeroed atom column, one per stage
  for (int i = threadIdx.x; i < kStemWgStages * (int)(kBlkW / 16); i += blockDim.x) {
    const int s = i / (int)(kBlkW / 16), o = i % (int)(kBlkW / 16);
    reinterpret_cast<uint4*>(smem + (size_t)s * kWgStage + kBlkW)[o] = make_uint4(0u, 0u, 0u, 0u);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (nb > 0) {
    if (warp == 5) {
      if (lane == 0) {
        int st = 0; uint32_t ph = 0;
        for (int i = 0; i < nb; ++i) {
          mbar_wait(&empty[st], ph ^ 1u);
          mbar_expect_tx(&full[st], kBlkW);
          tma_load_2d(smem + (size_t)st * kWgStage, &tmDy, &full[st], 0, (int)((b0 + i) * 64));
          if (++st == kStemWgStages) { st = 0; ph ^= 1u; }
        }
      }
    } else if (warp == 4) {
      int st = 0; uint32_t ph = 0;
      const uint32_t d_hi = desc_hi(1024, 2);
      const uint32_t leader = elect_one();
      for (int i = 0; i < nb; ++i) {
        mbar_wait(&full[st], ph);
        tc_fence_after();
        const uint32_t a16 = smem_u32(smem + (size_t)st * kWgStage) >> 4;
        const uint32_t b16 = a16 + ((2 * kBlkW) >> 4);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (leader)
            umma_bf16_ss2(tmem_base, desc_lo(a16 + 128u * k, kBlkW), d_hi, desc_lo(b16 + 128u * k, kBlkW), d_hi,
                          p.idesc, (i | k) != 0 ? 1u : 0u);
        }
        if (leader) {
          umma_commit(&empty[st]);
          if (i == nb - 1) umma_commit(accum);
        }
        __syncwarp();
        if (++st == kStemWgStages) { st = 0; ph ^= 1u; }
      }
    } else {
      // gather: two threads per voxel row (taps 0..13 | 14..26 and the K padding)
      // (the loads of block i+1 are issued before the stores of block i: their latency overlaps the stores, the
      // barrier hand-off and the wait for a free stage)
      const int r = threadIdx.x & 63, half = threadIdx.x >> 6;
      int st = 0; uint32_t ph = 0;
      uint2 cur[14], nxt[14];
      StemRow row = stem_row(p, r, b0 * 64 + r);
      if (half == 0) stem_load<0, 14>(p, row, cur); else stem_load<14, 27>(p, row, cur);
      for (int i = 0; i < nb; ++i) {
        StemRow row_n = row;
        if (i + 1 < nb) {
          row_n = stem_row(p, r, (b0 + i + 1) * 64 + r);
          if (half == 0) stem_load<0, 14>(p, row_n, nxt); else stem_load<14, 27>(p, row_n, nxt);
        }
        mbar_wait(&empty[st], ph ^ 1u);
        uint8_t* tile = smem + (size_t)st * kWgStage + 2 * kBlkW;
        if (half == 0) {
          stem_store<0, 14>(tile, kBlkW, row, cur);
        } else {
          stem_store<14, 27>(tile, kBlkW, row, cur);
          stem_pad(tile, kBlkW, row);
        }
        fence_proxy_async();
        mbar_arrive(&full[st]);
        if (++st == kStemWgStages) { st = 0; ph ^= 1u; }
        row = row_n;
#pragma unroll
        for (int j = 0; j < 14; ++j) cur[j] = nxt[j];
      }
      // epilogue: TMEM lane = cout (rows 64..127 are the zero half), column = k = tap * 4 + c
      mbar_wait(accum, 0);
      tc_fence_after();
      const int co = warp * 32 + lane;
      for (int c = 0; c < 112; c += 16) {
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
        tmem_ld_wait();
        if (co < 64) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int k = c + j, tap = k >> 2, ch = k & 3;
            if (tap < 27 && ch < 3) atomicAdd(p.dw + (co * 27 + tap) * 3 + ch, __uint_as_float(v[j]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

}  // namespace t2v

using namespace t2v;

extern "C" {

int t2v_stem_fprop(const void* xc, int32_t cpv, const void* wp, const float* bias, void* y, int64_t N, int32_t D,
                   int32_t H, int32_t W, int32_t relu, void* stream) {
  if (!xc || !wp || !y || N <= 0 || D <= 0 || H <= 0 || W <= 0 || (cpv != 4 && cpv != 16)) return T2V_ERR_ARG;
  StemParams p{};
  p.cshift = cpv == 4 ? 2 : 4;
  p.N = (int)N; p.D = D; p.H = H; p.W = W;
  p.P = (long long)N * D * H * W;
  if (p.P > 0x7ffffffLL) return T2V_ERR_ARG;      // 32-bit element offsets (16 per voxel)
  p.xc = reinterpret_cast<const __nv_bfloat16*>(xc);
  p.bias = bias;
  p.y = reinterpret_cast<__nv_bfloat16*>(y);
  p.relu = relu ? 1 : 0;
  p.idesc = make_idesc_bf16(128, 64, 0, 0);
  CUtensorMap tmW;
  int rc = make_w_map(&tmW, wp, 64, 128, 64, 64);
  if (rc) return rc;
  const size_t smem = 4 * kBlkA + 2 * kBlkW + 1024 + 128;
  cudaFuncSetAttribute(stem_fprop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  long long tiles = (p.P + 127) / 128;
  if (tiles > 2 * 148) tiles = 2 * 148;          // persistent: two CTAs per SM
  cudaStream_t cs = reinterpret_cast<cudaStream_t>(stream);
  ProfRec rec;
  t2v_conv_geom pg = {(int32_t)N, D, H, W, 3, 64, 3, 3, 3};
  if (g_prof_on) prof_begin(cs, &rec, 4, 2.0 * (double)p.P * 81.0 * 64.0, &pg, (int)tiles);
  stem_fprop_kernel<<<(unsigned)tiles, kStemFpThreads, smem, cs>>>(tmW, p);
  if (g_prof_on) prof_end(cs, &rec);
  count_launch();
  return check_last("stem_fprop");
}

int t2v_stem_wgrad(const void* dy, const void* xc, int32_t cpv, float* dw, int64_t N, int32_t D, int32_t H, int32_t W,
                   int32_t accumulate, void* stream) {
  if (!dy || !xc || !dw || N <= 0 || D <= 0 || H <= 0 || W <= 0 || (cpv != 4 && cpv != 16)) return T2V_ERR_ARG;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  StemParams p{};
  p.cshift = cpv == 4 ? 2 : 4;
  p.N = (int)N; p.D = D; p.H = H; p.W = W;
  p.P = (long long)N * D * H * W;
  if (p.P > 0x7ffffffLL) return T2V_ERR_ARG;
  p.xc = reinterpret_cast<const __nv_bfloat16*>(xc);
  p.dw = dw;
  p.idesc = make_idesc_bf16(128, 128, 1, 1);
  p.blocks_total = (p.P + 63) / 64;
  int ctas = 2 * 148;
  if (ctas > p.blocks_total) ctas = (int)p.blocks_total;
  p.per_cta = (int)((p.blocks_total + ctas - 1) / ctas);
  ctas = (int)((p.blocks_total + p.per_cta - 1) / p.per_cta);
  CUtensorMap tmDy;
  int rc = make_w_map(&tmDy, dy, (int)p.P, 64, 64, 64);
  if (rc) return rc;
  if (!accumulate) cudaMemsetAsync(dw, 0, 64 * 81 * sizeof(float), s);
  const size_t smem = kStemWgStages * kWgStage + 1024 + (2 * kStemWgStages + 1) * 8 + 16;
  cudaFuncSetAttribute(stem_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  ProfRec rec;
  t2v_conv_geom pg = {(int32_t)N, D, H, W, 3, 64, 3, 3, 3};
  if (g_prof_on) prof_begin(s, &rec, 5, 2.0 * (double)p.P * 81.0 * 64.0, &pg, ctas);
  stem_wgrad_kernel<<<ctas, kStemWgThreads, smem, s>>>(tmDy, p);
  if (g_prof_on) prof_end(s, &rec);
  count_launch();
  return check_last("stem_wgrad");
}

}  // extern "C"
