// Micro-probes that pin down tcgen05 facts the conv-engine design depends on (run on a B200 via gpurun):
//   1. MMA issue/throughput from RESIDENT shared memory for M=128, N in {32,64,128,256}, K-major and
//      MN-major operands (is a 128x64 tile shared-memory-read bound?).
//   2. K-major SW128 A operand whose start address is shifted by s rows (s*128 B, not 1024-aligned):
//      does the swizzle follow the absolute address (base_offset = 0) or does it need base_offset?
//   3. MN-major SW128 A operand: M = 128 built from two 64-channel groups LBO bytes apart (two filter
//      taps = two row shifts of the same tile) + row-shifted starts along K.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I include scripts/umma_probe.cu -o gpurun_out/umma_probe
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../txt2vid_b200/csrc/t2v_common.cuh"

using namespace t2v;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint64_t desc_with_base_offset(uint64_t d, uint32_t bo) {
  return d | ((uint64_t)(bo & 7u) << 49);
}

// ------------------------------------------------------------------------------------------ probe 1
// grid = #SMs, one CTA per SM; thread 0 issues `iters` groups of 4 MMAs (K = 64) on resident smem.
__global__ void __launch_bounds__(128, 1)
probe_rate(int N, int mn_major, int iters, long long* cycles_out, int a_stride_bytes) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  // fill 96 KB with small bf16 values
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async();
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t sa = smem_u32(smem), sb = sa + 64 * 1024;
    const uint32_t idesc = make_idesc_bf16(128, N, mn_major, mn_major);
    uint64_t adesc, bdesc;
    if (!mn_major) {
      adesc = make_smem_desc(sa, 0, 1024, 2);
      bdesc = make_smem_desc(sb, 0, 1024, 2);
    } else {
      adesc = make_smem_desc(sa, 8192, 1024, 2);
      bdesc = make_smem_desc(sb, 8192, 1024, 2);
    }
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      // rotate over 4 A tiles (a_stride_bytes apart) so that consecutive MMAs read different smem
      const uint64_t ad = adesc + (uint64_t)(((it & 3) * a_stride_bytes) >> 4);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t step = mn_major ? (uint64_t)(128 * k) : (uint64_t)(2 * k);
        umma_bf16_ss(tmem + (uint32_t)((it & 1) * 256), ad + step, bdesc + step, idesc, 1u);
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    cycles_out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// ------------------------------------------------------------------------------------------ probe 2/3
// smem "tile": R rows x 64 bf16, stored like TMA SWIZZLE_128B writes it (16-byte chunk j of row r lands at
// chunk position j ^ (r & 7)).  Values: tile[r][c] from global.  B tile: 64 rows x 64 bf16, same storage.
// mode 0: K-major A = rows [s, s+128) of the tile (K = 64 channels), B K-major (N = 64 rows, K = 64).
//         D[i][n] = sum_c tile[s+i][c] * B[n][c]
// mode 1: MN-major A: M = 128 = channels of rows shifted by s1 (m < 64) / s2 (m >= 64), K = 64 rows;
//         B MN-major: dy tile rows = K, 64 channels = N.  D[m][n] = sum_k tile[sm + k][m%64] * B[k][n]
__global__ void __launch_bounds__(128, 1)
probe_shift(const __nv_bfloat16* tile_g, int R, const __nv_bfloat16* b_g, int mode, int s1, int s2, int use_bo,
            float* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  uint8_t* sA = smem;
  uint8_t* sB = smem + 64 * 1024;
  for (int i = threadIdx.x; i < R * 64; i += blockDim.x) {
    const int r = i / 64, c = i % 64;
    const int off = r * 128 + (((c >> 3) ^ (r & 7)) << 4) + (c & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(sA + off) = tile_g[i];
  }
  for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {
    const int r = i / 64, c = i % 64;
    const int off = r * 128 + (((c >> 3) ^ (r & 7)) << 4) + (c & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(sB + off) = b_g[i];
  }
  fence_proxy_async();
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&tmem_slot, 64); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t sa = smem_u32(sA), sb = smem_u32(sB);
    if (mode == 0) {
      const uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
      const uint32_t a0 = sa + (uint32_t)s1 * 128u;
      uint64_t adesc = make_smem_desc(a0, 0, 1024, 2);
      if (use_bo) adesc = desc_with_base_offset(adesc, (a0 >> 7) & 7u);
      const uint64_t bdesc = make_smem_desc(sb, 0, 1024, 2);
      for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, k != 0);
    } else {
      const uint32_t idesc = make_idesc_bf16(128, 64, 1, 1);
      const uint32_t a0 = sa + (uint32_t)s1 * 128u;
      const uint32_t lbo = (uint32_t)(s2 - s1) * 128u;     // second 64-channel group = the tile s2 rows down
      uint64_t adesc = make_smem_desc(a0, lbo, 1024, 2);
      if (use_bo) adesc = desc_with_base_offset(adesc, (a0 >> 7) & 7u);
      const uint64_t bdesc = make_smem_desc(sb, 8192, 1024, 2);
      for (int k = 0; k < 4; ++k)
        umma_bf16_ss(tmem, adesc + (uint64_t)(128 * k), bdesc + (uint64_t)(128 * k), idesc, k != 0);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c = 0; c < 64; c += 16) {
    uint32_t v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * 64 + c + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 64); }
}

static float bf(float v) { return __bfloat162float(__float2bfloat16(v)); }

int main() {
  int dev = 0;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  printf("device %s, %d SMs, clock %d kHz\n", prop.name, prop.multiProcessorCount, prop.clockRate);
  const int nsm = prop.multiProcessorCount;

  // ---------------- probe 2/3 first (correctness)
  const int R = 256;
  std::vector<__nv_bfloat16> tile(R * 64), bt(64 * 64);
  std::vector<float> tf(R * 64), bfv(64 * 64);
  srand(7);
  for (int i = 0; i < R * 64; ++i) { float v = (float)((rand() % 17) - 8) / 8.f; tile[i] = __float2bfloat16(v); tf[i] = bf(v); }
  for (int i = 0; i < 64 * 64; ++i) { float v = (float)((rand() % 13) - 6) / 4.f; bt[i] = __float2bfloat16(v); bfv[i] = bf(v); }
  __nv_bfloat16 *d_tile, *d_b;
  float* d_out;
  CK(cudaMalloc(&d_tile, tile.size() * 2));
  CK(cudaMalloc(&d_b, bt.size() * 2));
  CK(cudaMalloc(&d_out, 128 * 64 * 4));
  CK(cudaMemcpy(d_tile, tile.data(), tile.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_b, bt.data(), bt.size() * 2, cudaMemcpyHostToDevice));
  const size_t smem = 64 * 1024 + 8 * 1024 + 2048;
  CK(cudaFuncSetAttribute(probe_shift, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  std::vector<float> out(128 * 64);
  for (int use_bo = 0; use_bo < 2; ++use_bo) {
    for (int s = 0; s <= 9; ++s) {
      CK(cudaMemset(d_out, 0, 128 * 64 * 4));
      probe_shift<<<1, 128, smem>>>(d_tile, R, d_b, 0, s, 0, use_bo, d_out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode0 s=%d bo=%d: CUDA error %s\n", s, use_bo, cudaGetErrorString(e)); return 1; }
      CK(cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost));
      double maxerr = 0;
      for (int i = 0; i < 128; ++i)
        for (int n = 0; n < 64; ++n) {
          double ref = 0;
          for (int c = 0; c < 64; ++c) ref += (double)tf[(s + i) * 64 + c] * bfv[n * 64 + c];
          maxerr = fmax(maxerr, fabs(ref - out[i * 64 + n]));
        }
      printf("probe2 K-major row shift s=%d base_offset=%s: max abs err %.4g %s\n", s, use_bo ? "(addr>>7)&7" : "0", maxerr,
             maxerr < 1e-3 ? "OK" : "MISMATCH");
    }
  }
  const int pairs[][2] = {{0, 8}, {8, 24}, {16, 96}, {1, 9}, {3, 12}, {5, 5}, {2, 9}, {0, 1}};
  for (int use_bo = 0; use_bo < 1; ++use_bo) {
    for (auto& pr : pairs) {
      const int s1 = pr[0], s2 = pr[1];
      CK(cudaMemset(d_out, 0, 128 * 64 * 4));
      probe_shift<<<1, 128, smem>>>(d_tile, R, d_b, 1, s1, s2, use_bo, d_out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode1 s1=%d s2=%d: CUDA error %s\n", s1, s2, cudaGetErrorString(e)); return 1; }
      CK(cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost));
      double maxerr = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 64; ++n) {
          const int sm = m < 64 ? s1 : s2;
          double ref = 0;
          for (int k = 0; k < 64; ++k) ref += (double)tf[(sm + k) * 64 + (m & 63)] * bfv[k * 64 + n];
          maxerr = fmax(maxerr, fabs(ref - out[m * 64 + n]));
        }
      printf("probe3 MN-major stacked taps s1=%d s2=%d base_offset=%s: max abs err %.4g %s\n", s1, s2,
             use_bo ? "(addr>>7)&7" : "0", maxerr, maxerr < 1e-3 ? "OK" : "MISMATCH");
    }
  }

  // ---------------- probe 1 (rates)
  long long* d_cyc;
  CK(cudaMalloc(&d_cyc, nsm * sizeof(long long)));
  const size_t smem1 = 96 * 1024 + 2048;
  CK(cudaFuncSetAttribute(probe_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
  std::vector<long long> cyc(nsm);
  const int iters = 4096;
  for (int mn = 0; mn < 2; ++mn)
    for (int stride : {0, 16384})
      for (int N : {32, 64, 128, 256}) {
        if (mn && N > 64 && false) continue;
        for (int grid : {1, nsm}) {
          probe_rate<<<grid, 128, smem1>>>(N, mn, iters, d_cyc, stride);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("probe1 N=%d mn=%d: CUDA error %s\n", N, mn, cudaGetErrorString(e)); return 1; }
          CK(cudaMemcpy(cyc.data(), d_cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost));
          long long mx = 0;
          for (int i = 0; i < grid; ++i) mx = cyc[i] > mx ? cyc[i] : mx;
          const double per = (double)mx / (iters * 4.0);
          printf("probe1 %s M=128 N=%3d a_rotate=%5d grid=%3d: %.1f cycles per MMA (K=16)  floor %.0f  -> %.0f%% of tensor peak\n",
                 mn ? "MN-major" : "K-major ", N, stride, grid, per, 128.0 * N / 256.0, 100.0 * (128.0 * N / 256.0) / per);
        }
      }
  printf("done\n");
  return 0;
}
