// Host-side TMA tensor-map builders shared by the conv-engine translation units.
#pragma once
#include "t2v_common.cuh"

namespace t2v {

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                        const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_tmapEncodeTiled get_encode() {
  static PFN_tmapEncodeTiled fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_tmapEncodeTiled>(ptr);
  }
  return fn;
}

inline CUtensorMapSwizzle swizzle_for(int inner_bytes) {
  return inner_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                            : (inner_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

// activation CL [N][D][H][W][C] bf16, box (bc, bw, bh, bd, bn)
inline int make_act_map(CUtensorMap* m, const void* base, int N, int D, int H, int W, int C, int bc,
                        int bw, int bh, int bd, int bn) {
  PFN_tmapEncodeTiled enc = get_encode();
  if (!enc) return T2V_ERR_DRIVER;
  cuuint64_t gdim[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
  cuuint64_t gstr[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2,
                        (cuuint64_t)D * H * W * C * 2};
  cuuint32_t box[5] = {(cuuint32_t)bc, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bd, (cuuint32_t)bn};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), gdim, gstr, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(bc * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? T2V_OK : T2V_ERR_ARG;
}

// weight [rows][K] bf16 (K contiguous), box (bk, brows)
inline int make_w_map(CUtensorMap* m, const void* base, int rows, int K, int bk, int brows) {
  PFN_tmapEncodeTiled enc = get_encode();
  if (!enc) return T2V_ERR_DRIVER;
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {(cuuint32_t)bk, (cuuint32_t)brows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(bk * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? T2V_OK : T2V_ERR_ARG;
}

// Optional per-launch CUDA-event timing (defined in igemm_sm100.cu; bench.py reads it through t2v_profile_read).
struct ProfRec { cudaEvent_t a, b; double flops; int kind; t2v_conv_geom g; int ctas; };
extern bool g_prof_on;
void prof_begin(cudaStream_t s, ProfRec* r, int kind, double flops, const t2v_conv_geom* g, int ctas);
void prof_end(cudaStream_t s, ProfRec* r);

}  // namespace t2v
