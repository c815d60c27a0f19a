// sl_pop_bg_simt: the background (class 0) logit of the POP head in exact fp32 on CUDA cores.
//   reference: networks/pspnet_pop.py:112,118 (out_bg = q - sum_k out_fg_k) pushed through
//   classifier / classifier_n (:154-157, :178-182).  With W1' = W1 (I - S^T S) folded by
//   sl_pop_prepare:   logit_0 = w3 . relu(W2 relu(W1' q)).
// This is the precise / any-shape path (and the on-device check for the tensor-core path):
// a dense per-pixel C x C two-layer MLP, 4*C^2 FLOP per pixel, FP32-FMA bound.
//
// One CTA = 32 pixels.  X[C][32] and H1[C][32] live in shared memory (fp32); 256 threads each own
// an 8 (out-channel) x 8 (pixel) register tile per 512-channel chunk and stream the transposed
// weights W^T[i][o] (L2-resident, 2 KB per input channel) with 128-bit loads.
#include "common.cuh"

namespace sl {

constexpr int BG_P = 32;  // pixels per CTA

// acc[8 o][8 p] += sum_i Wt[i][o0..o0+8) * X[i][p0..p0+8)
__device__ __forceinline__ void mlp_tile(const float* __restrict__ Wt, const float* __restrict__ X, int C, int o0,
                                         int p0, float (&acc)[8][8]) {
#pragma unroll 2
  for (int i = 0; i < C; ++i) {
    const float4 wa = __ldg(reinterpret_cast<const float4*>(Wt + static_cast<size_t>(i) * C + o0));
    const float4 wb = __ldg(reinterpret_cast<const float4*>(Wt + static_cast<size_t>(i) * C + o0 + 4));
    const float4 xa = *reinterpret_cast<const float4*>(X + i * BG_P + p0);
    const float4 xb = *reinterpret_cast<const float4*>(X + i * BG_P + p0 + 4);
    const float w[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
    const float x[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
      for (int p = 0; p < 8; ++p) acc[a][p] = fmaf(w[a], x[p], acc[a][p]);
  }
}

__global__ void __launch_bounds__(256, 1) pop_bg_simt_kernel(const uint16_t* __restrict__ feat, int C, int N,
                                                             const float* __restrict__ W1p_t,
                                                             const float* __restrict__ W2_t,
                                                             const float* __restrict__ w3, float* __restrict__ logits,
                                                             int Ktot, int ch) {
  extern __shared__ __align__(16) float smem[];
  float* X = smem;                 // [C][32]
  float* H = smem + C * BG_P;      // [C][32]
  __shared__ float red[64][BG_P + 1];
  const int b = blockIdx.y;
  const int n0 = blockIdx.x * BG_P;
  const int tp = threadIdx.x & 3;   // pixel group: pixels tp*8 .. tp*8+7
  const int to = threadIdx.x >> 2;  // out-channel group within a 512-chunk: channels to*8 .. to*8+7

  // stage X: C rows of 32 bf16 pixels (64 B) -> fp32
  const uint16_t* src = feat + (static_cast<size_t>(b) * C) * N + n0;
  for (int idx = threadIdx.x; idx < C * 4; idx += 256) {
    const int c = idx >> 2, q = idx & 3;
    float4 lo = make_float4(0, 0, 0, 0), hi = lo;
    if (n0 + q * 8 < N) {
      const uint4 v = ld_stream_u4(src + static_cast<size_t>(c) * N + q * 8);
      lo = make_float4(bf16lo(v.x), bf16hi(v.x), bf16lo(v.y), bf16hi(v.y));
      hi = make_float4(bf16lo(v.z), bf16hi(v.z), bf16lo(v.w), bf16hi(v.w));
    }
    *reinterpret_cast<float4*>(X + c * BG_P + q * 8) = lo;
    *reinterpret_cast<float4*>(X + c * BG_P + q * 8 + 4) = hi;
  }
  __syncthreads();

  // layer 1: H = relu(W1' X)
  for (int ob = 0; ob < C; ob += 512) {
    const int o0 = ob + to * 8;
    if (o0 < C) {
      float acc[8][8];
#pragma unroll
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int p = 0; p < 8; ++p) acc[a][p] = 0.f;
      mlp_tile(W1p_t, X, C, o0, tp * 8, acc);
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        float4* dst = reinterpret_cast<float4*>(H + (o0 + a) * BG_P + tp * 8);
        dst[0] = make_float4(fmaxf(acc[a][0], 0.f), fmaxf(acc[a][1], 0.f), fmaxf(acc[a][2], 0.f), fmaxf(acc[a][3], 0.f));
        dst[1] = make_float4(fmaxf(acc[a][4], 0.f), fmaxf(acc[a][5], 0.f), fmaxf(acc[a][6], 0.f), fmaxf(acc[a][7], 0.f));
      }
    }
  }
  __syncthreads();

  // layer 2 + layer 3: out[p] = sum_o w3[o] * relu(W2 H)[o][p]
  float outp[8];
#pragma unroll
  for (int p = 0; p < 8; ++p) outp[p] = 0.f;
  for (int ob = 0; ob < C; ob += 512) {
    const int o0 = ob + to * 8;
    if (o0 < C) {
      float acc[8][8];
#pragma unroll
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int p = 0; p < 8; ++p) acc[a][p] = 0.f;
      mlp_tile(W2_t, H, C, o0, tp * 8, acc);
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        const float wv = __ldg(w3 + o0 + a);
#pragma unroll
        for (int p = 0; p < 8; ++p) outp[p] = fmaf(wv, fmaxf(acc[a][p], 0.f), outp[p]);
      }
    }
  }
#pragma unroll
  for (int p = 0; p < 8; ++p) red[to][tp * 8 + p] = outp[p];
  __syncthreads();
  if (threadIdx.x < BG_P && n0 + threadIdx.x < N) {
    float t = 0.f;
#pragma unroll 8
    for (int g = 0; g < 64; ++g) t += red[g][threadIdx.x];
    logits[(static_cast<size_t>(b) * Ktot + ch) * N + n0 + threadIdx.x] = t;
  }
}

}  // namespace sl

extern "C" int sl_pop_bg_simt(const uint16_t* feat, int B, int C, int N, const float* W1p_t, const float* W2_t,
                              const float* w3_bg, float* logits, int Ktot, int ch, void* stream) {
  SL_CHECK_PTR(feat); SL_CHECK_PTR(W1p_t); SL_CHECK_PTR(W2_t); SL_CHECK_PTR(w3_bg); SL_CHECK_PTR(logits);
  SL_CHECK_ARG(B >= 1 && B <= 65535 && Ktot >= 1 && Ktot <= SL_MAX_CLASSES && ch >= 0 && ch < Ktot);
  SL_CHECK_ARG(C >= 8 && C <= 768 && C % 8 == 0 && N >= 8 && N % 8 == 0);
  SL_CHECK_ALIGN(feat, 16); SL_CHECK_ALIGN(W1p_t, 16); SL_CHECK_ALIGN(W2_t, 16);
  const size_t smem = static_cast<size_t>(2) * C * sl::BG_P * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(sl::pop_bg_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  dim3 grid((N + sl::BG_P - 1) / sl::BG_P, B);
  sl::pop_bg_simt_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(feat, C, N, W1p_t, W2_t, w3_bg,
                                                                                logits, Ktot, ch);
  return SL_LAUNCH_RESULT();
}
