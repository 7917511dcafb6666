/*
 * t2v.h -- C ABI of libt2v_b200.so: the sm_100a kernels behind txt2vid's GAN training step.
 *
 * The reference (miguelmartin75/txt2vid) has no FFI of its own: its "plugin" boundary is Python
 * reflection (txt2vid/util/reflection.py:12-50) and every FLOP is an ATen call inside
 * txt2vid/models/ and txt2vid/gan/.  Each entry point below replaces the ATen call sites named in its
 * comment (file:line relative to the reference tree).  The Python host side (txt2vid_b200/) binds
 * these with ctypes; see INTEGRATION.md for the binding a maintainer of the reference would add.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless named host_*;
 *   - every call enqueues work on `stream` (a cudaStream_t passed as void*), never synchronises,
 *     never allocates device memory and never throws; returns T2V_OK (0) or a negative T2V_ERR_*;
 *   - "CL" = channels-last activation layout [N][D][H][W][C] (2-D feature maps use D = 1);
 *     activations are bf16, accumulation is fp32, master weights / gradients of weights are fp32;
 *   - conv weights are [Cout][kd][kh][kw][Cin] ("channels-last" memory of the reference's
 *     (Cout,Cin,kd,kh,kw) parameter), so a weight gradient written by the kernels IS the parameter
 *     gradient's memory.
 */
#ifndef T2V_H_
#define T2V_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define T2V_OK 0
#define T2V_ERR_ARG (-1)      /* unsupported shape / null pointer / misalignment */
#define T2V_ERR_LAUNCH (-2)   /* CUDA reported an error at launch */
#define T2V_ERR_DRIVER (-3)   /* driver entry point (tensor-map encode) unavailable */

/* algo selector for the convolution engine */
#define T2V_ALGO_AUTO 0   /* tcgen05 implicit GEMM when the shape allows it, else SIMT */
#define T2V_ALGO_TC 1     /* force the tcgen05/TMEM/TMA kernel (error if shape unsupported) */
#define T2V_ALGO_SIMT 2   /* force the CUDA-core kernel (cross-check + odd shapes) */

/* epilogue flags of t2v_conv_fprop */
#define T2V_EPI_RELU 1u       /* y = max(y, 0) after bias/residual */
#define T2V_EPI_OUT_F32 2u    /* y is fp32 CL instead of bf16 CL */

/* Stride-1, "same"-padded convolution geometry (every conv on the TGANv2 path:
 * models/layers.py:174,177,183,231,233,237,251; models/resnet3d.py:12-17; 1x1(x1) convs of the
 * attention blocks layers.py:16-19,45-48; nn.Linear is the k=1, D=H=W=1 case). */
typedef struct {
  int32_t N, D, H, W;    /* batch (or merged batch*frames) and spatial extents            */
  int32_t Cin, Cout;     /* channels; both multiples of 16 (pad with zeros otherwise)      */
  int32_t kd, kh, kw;    /* kernel extents, each 1 or 3; padding is k/2                     */
} t2v_conv_geom;

/* library / bookkeeping ------------------------------------------------------------------- */
int t2v_version(void);
/* number of kernels launched by this library since load (bench.py: "gpu_launches") */
unsigned long long t2v_launch_count(void);

/* convolution engine (tcgen05 implicit GEMM; replaces F.conv2d/conv3d/linear = cuDNN/cuBLAS)  */
/* y[n,d,h,w,co] = sum_{taps,ci} x[n,d+a-pd,h+b-ph,w+c-pw,ci] * w[co,a,b,c,ci] + bias[co] (+ residual)
 * x: bf16 CL, w: bf16 [Cout][taps][Cin], bias: fp32[Cout] or NULL, residual: bf16 CL [..,Cout] or NULL */
int t2v_conv_fprop(const t2v_conv_geom* g, const void* x, const void* w, const float* bias,
                   const void* residual, void* y, uint32_t epi_flags, int algo, void* stream);
/* dx = conv_transpose(dy, w): same engine, fed with the weight pack made by t2v_pack_dgrad_weight:
 * wT bf16 [Cin][taps (flipped)][Cout]; g is the FORWARD geometry. */
int t2v_conv_dgrad(const t2v_conv_geom* g, const void* dy, const void* wT, const void* residual,
                   void* dx, uint32_t epi_flags, int algo, void* stream);
/* dw[co,tap,ci] (+)= sum_pos dy[pos,co] * x[pos+tap,ci]   (fp32, layout [Cout][taps][Cin]).
 * accumulate = 0 overwrites dw, 1 adds into it.                                              */
int t2v_conv_wgrad(const t2v_conv_geom* g, const void* dy, const void* x, float* dw,
                   int accumulate, int algo, void* stream);
/* fp32 master weight [Cout][taps][Cin] -> bf16 same layout (fprop operand)                    */
int t2v_cast_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream);
int t2v_cast_bf16_to_f32(const void* src, float* dst, int64_t n, void* stream);
/* fp32 master weight [Cout][taps][Cin] -> bf16 [Cin][taps reversed][Cout] (dgrad operand)     */
int t2v_pack_dgrad_weight(const float* w, void* wT, int32_t Cout, int32_t taps, int32_t Cin,
                          void* stream);

#ifdef __cplusplus
}
#endif
#endif /* T2V_H_ */
