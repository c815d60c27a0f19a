// sl_pop_fg_lowres: the K foreground logits of the POP head at feature resolution.
//   reference: networks/pspnet_pop.py:108-109,114-115 (proj = s_hat @ q, out_fg = proj * s_hat)
//              followed by classifier / classifier_n on every rank-1 vector (:150-157, :178-182).
// Because the classifier is bias-free with ReLUs it is positively homogeneous, so
//   logit_k(pixel) = p >= 0 ? p * alpha_k : -p * beta_k,   p = s_hat_k . q(pixel)
// and the kernel is a skinny [K x C] x [C x N] contraction streamed once over the bf16 features:
// HBM-bound (C*N*2 bytes in, K*N*4 bytes out).
//
// Mapping: one CTA = 64 pixels x all C channels of one image.  256 threads = 8 warps; a lane owns
// 8 consecutive pixels (one 128-bit load per channel) and one of 32 channel groups
// (group = warp*4 + lane/8, channels group, group+32, ...).  A warp-level load therefore touches
// 4 channels x 128 contiguous bytes.  Partial sums meet through two shuffles (across lane/8) and
// one shared-memory pass (across warps), in a fixed order -> bit-reproducible.
#include "common.cuh"

namespace sl {

struct ChMap { int ch[SL_MAX_CLASSES]; };

template <int KP>  // classes padded to a multiple of 4
__global__ void __launch_bounds__(256) pop_fg_kernel(const uint16_t* __restrict__ feat, int C, int N,
                                                     const float* __restrict__ s_hat, const float* __restrict__ alpha,
                                                     const float* __restrict__ beta, int K, int k_base,
                                                     float* __restrict__ logits, int Ktot, ChMap map) {
  extern __shared__ __align__(16) float smem[];
  float* st = smem;                    // [C][KP]  transposed prototypes (zero padded)
  float* part = smem + C * KP;         // [8][KP][64]
  const int b = blockIdx.y;
  const int n0 = blockIdx.x * 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pl = lane & 7, cg = lane >> 3;

  for (int idx = threadIdx.x; idx < C * KP; idx += 256) {
    const int c = idx / KP, k = idx - c * KP;
    st[idx] = (k_base + k < K) ? s_hat[static_cast<size_t>(k_base + k) * C + c] : 0.f;
  }
  __syncthreads();

  float acc[KP][8];
#pragma unroll
  for (int k = 0; k < KP; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[k][j] = 0.f;

  const int n = n0 + pl * 8;
  const bool live = n < N;
  const uint16_t* base = feat + (static_cast<size_t>(b) * C) * N + n;
  constexpr int U = 4;
  int c = warp * 4 + cg;
  for (; c + 32 * (U - 1) < C; c += 32 * U) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      v[u] = live ? ld_stream_u4(base + static_cast<size_t>(c + 32 * u) * N) : make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const float x[8] = {bf16lo(v[u].x), bf16hi(v[u].x), bf16lo(v[u].y), bf16hi(v[u].y),
                          bf16lo(v[u].z), bf16hi(v[u].z), bf16lo(v[u].w), bf16hi(v[u].w)};
      const float4* srow = reinterpret_cast<const float4*>(st + (c + 32 * u) * KP);
#pragma unroll
      for (int k4 = 0; k4 < KP / 4; ++k4) {
        const float4 s = srow[k4];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[4 * k4 + 0][j] = fmaf(s.x, x[j], acc[4 * k4 + 0][j]);
          acc[4 * k4 + 1][j] = fmaf(s.y, x[j], acc[4 * k4 + 1][j]);
          acc[4 * k4 + 2][j] = fmaf(s.z, x[j], acc[4 * k4 + 2][j]);
          acc[4 * k4 + 3][j] = fmaf(s.w, x[j], acc[4 * k4 + 3][j]);
        }
      }
    }
  }
  for (; c < C; c += 32) {
    const uint4 v = live ? ld_stream_u4(base + static_cast<size_t>(c) * N) : make_uint4(0, 0, 0, 0);
    const float x[8] = {bf16lo(v.x), bf16hi(v.x), bf16lo(v.y), bf16hi(v.y),
                        bf16lo(v.z), bf16hi(v.z), bf16lo(v.w), bf16hi(v.w)};
    const float4* srow = reinterpret_cast<const float4*>(st + c * KP);
#pragma unroll
    for (int k4 = 0; k4 < KP / 4; ++k4) {
      const float4 s = srow[k4];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[4 * k4 + 0][j] = fmaf(s.x, x[j], acc[4 * k4 + 0][j]);
        acc[4 * k4 + 1][j] = fmaf(s.y, x[j], acc[4 * k4 + 1][j]);
        acc[4 * k4 + 2][j] = fmaf(s.z, x[j], acc[4 * k4 + 2][j]);
        acc[4 * k4 + 3][j] = fmaf(s.w, x[j], acc[4 * k4 + 3][j]);
      }
    }
  }

  // across the 4 channel groups of the warp (lane bits 3,4)
#pragma unroll
  for (int k = 0; k < KP; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float a = acc[k][j];
      a += __shfl_xor_sync(0xffffffffu, a, 8);
      a += __shfl_xor_sync(0xffffffffu, a, 16);
      acc[k][j] = a;
    }
  // lane (cg, pl) publishes classes k == cg (mod 4) for its 8 pixels
#pragma unroll
  for (int k = 0; k < KP; ++k) {
    if ((k & 3) == cg) {
      float4* dst = reinterpret_cast<float4*>(part + (warp * KP + k) * 64 + pl * 8);
      dst[0] = make_float4(acc[k][0], acc[k][1], acc[k][2], acc[k][3]);
      dst[1] = make_float4(acc[k][4], acc[k][5], acc[k][6], acc[k][7]);
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < KP * 64; idx += 256) {
    const int k = idx >> 6, px = idx & 63;
    const int kk = k_base + k;
    if (kk >= K || n0 + px >= N) continue;
    float p = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) p += part[(w * KP + k) * 64 + px];
    const float v = p >= 0.f ? p * alpha[kk] : -p * beta[kk];
    logits[(static_cast<size_t>(b) * Ktot + map.ch[kk]) * N + n0 + px] = v;
  }
}

template <int KP>
static int launch_fg(const uint16_t* feat, int B, int C, int N, const float* s_hat, const float* alpha,
                     const float* beta, int K, int k_base, float* logits, int Ktot, const ChMap& map,
                     cudaStream_t st) {
  const size_t smem = (static_cast<size_t>(C) * KP + 8 * KP * 64) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(pop_fg_kernel<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  dim3 grid((N + 63) / 64, B);
  pop_fg_kernel<KP><<<grid, 256, smem, st>>>(feat, C, N, s_hat, alpha, beta, K, k_base, logits, Ktot, map);
  return SL_LAUNCH_RESULT();
}

}  // namespace sl

extern "C" int sl_pop_fg_lowres(const uint16_t* feat, int B, int C, int N, const float* s_hat, const float* alpha,
                                const float* beta, int K, float* logits, int Ktot, const int* ch_map_host,
                                void* stream) {
  SL_CHECK_PTR(feat); SL_CHECK_PTR(s_hat); SL_CHECK_PTR(alpha); SL_CHECK_PTR(beta); SL_CHECK_PTR(logits);
  SL_CHECK_PTR(ch_map_host);
  SL_CHECK_ARG(B >= 1 && B <= 65535 && K >= 1 && K < SL_MAX_CLASSES && Ktot >= K && Ktot <= SL_MAX_CLASSES);
  SL_CHECK_ARG(C >= 8 && C <= 1024 && C % 8 == 0 && N >= 8 && N % 8 == 0);
  SL_CHECK_ALIGN(feat, 16);
  sl::ChMap map;
  for (int k = 0; k < SL_MAX_CLASSES; ++k) map.ch[k] = 0;
  for (int k = 0; k < K; ++k) {
    SL_CHECK_ARG(ch_map_host[k] >= 0 && ch_map_host[k] < Ktot);
    map.ch[k] = ch_map_host[k];
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // Up to 16 classes per pass keep the accumulators in registers; more classes take extra passes.
  for (int k_base = 0; k_base < K; k_base += 16) {
    const int kc = (K - k_base) < 16 ? (K - k_base) : 16;
    int rc;
    if (kc <= 4) rc = sl::launch_fg<4>(feat, B, C, N, s_hat, alpha, beta, K, k_base, logits, Ktot, map, st);
    else if (kc <= 8) rc = sl::launch_fg<8>(feat, B, C, N, s_hat, alpha, beta, K, k_base, logits, Ktot, map, st);
    else if (kc <= 12) rc = sl::launch_fg<12>(feat, B, C, N, s_hat, alpha, beta, K, k_base, logits, Ktot, map, st);
    else rc = sl::launch_fg<16>(feat, B, C, N, s_hat, alpha, beta, K, k_base, logits, Ktot, map, st);
    if (rc != 0) return rc;
  }
  return SL_OK;
}
