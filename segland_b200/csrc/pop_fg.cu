// sl_pop_fg_lowres: the K foreground logits of the POP head at feature resolution.
//   reference: networks/pspnet_pop.py:108-109,114-115 (proj = s_hat @ q, out_fg = proj * s_hat)
//              followed by classifier / classifier_n on every rank-1 vector (:150-157, :178-182).
// Because the classifier is bias-free with ReLUs it is positively homogeneous, so
//   logit_k(pixel) = p >= 0 ? p * alpha_k : -p * beta_k,   p = s_hat_k . q(pixel)
// and the kernel is a skinny [K x C] x [C x N] contraction streamed once over the bf16 features:
// HBM-bound (C*N*2 bytes in, K*N*4 bytes out) as long as the FMA pipe keeps up, which at K ~ 7..11
// classes per feature byte pair is close: the inner product runs on packed fma.rn.f32x2.
//
// Persistent CTAs (2 per SM) stage the transposed prototypes in shared memory once and then loop
// over work items of PX*32 consecutive pixels of one image.  Within an item warp w owns channels
// [w*C/8, (w+1)*C/8) and lane l owns PX consecutive pixels, so every warp-level load is one
// contiguous 512-byte (PX=8) or 256-byte (PX=4) row segment, PF of them in flight per thread, and
// the only cross-thread traffic is one 8-way shared-memory reduction per item, in fixed order
// (bit-reproducible).
#include "common.cuh"

namespace sl {

struct ChMap { int ch[SL_MAX_CLASSES]; };

constexpr int FG_WARPS = 8;
constexpr int FG_THREADS = FG_WARPS * 32;
constexpr int FG_PF = 8;   // loads in flight per thread

// d = a * b + d on two packed fp32 lanes (sm_100 FFMA2)
__device__ __forceinline__ void ffma2(float2& d, const float2 a, const float2 b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;"
      : "+l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<const unsigned long long&>(a)), "l"(reinterpret_cast<const unsigned long long&>(b)));
}

template <int PX> struct PxVec;
template <> struct PxVec<8> {
  using T = uint4;
  static __device__ __forceinline__ T load(const uint16_t* p) { return ld_stream_u4(p); }
  static __device__ __forceinline__ T zero() { return make_uint4(0, 0, 0, 0); }
  static __device__ __forceinline__ uint32_t word(const T& v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }
};
template <> struct PxVec<4> {
  using T = uint2;
  static __device__ __forceinline__ T load(const uint16_t* p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
  }
  static __device__ __forceinline__ T zero() { return make_uint2(0, 0); }
  static __device__ __forceinline__ uint32_t word(const T& v, int i) { return i == 0 ? v.x : v.y; }
};

template <int KP, int PX>  // KP: classes padded to a multiple of 4; PX: pixels per thread
__global__ void __launch_bounds__(FG_THREADS, KP <= 8 ? 2 : 1)
pop_fg_kernel(const uint16_t* __restrict__ feat, int B, int C, int N, const float* __restrict__ s_hat,
              const float* __restrict__ alpha, const float* __restrict__ beta, int K, int k_base,
              float* __restrict__ logits, int Ktot, ChMap map) {
  constexpr int ITEM_PX = PX * 32;
  extern __shared__ __align__(16) float smem[];
  float* st = smem;                      // [C][KP] transposed prototypes (zero padded)
  float* part = smem + C * KP;           // [FG_WARPS][KP][ITEM_PX] partial sums
  __shared__ int ch_of[SL_MAX_CLASSES];
  __shared__ float alpha_s[SL_MAX_CLASSES], beta_s[SL_MAX_CLASSES];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < SL_MAX_CLASSES) {
    int ch = 0;
#pragma unroll
    for (int k = 0; k < SL_MAX_CLASSES; ++k) ch = (k == static_cast<int>(threadIdx.x)) ? map.ch[k] : ch;
    ch_of[threadIdx.x] = ch;
    alpha_s[threadIdx.x] = static_cast<int>(threadIdx.x) < K ? alpha[threadIdx.x] : 0.f;
    beta_s[threadIdx.x] = static_cast<int>(threadIdx.x) < K ? beta[threadIdx.x] : 0.f;
  }

  for (int idx = threadIdx.x; idx < C * KP; idx += FG_THREADS) {
    const int c = idx / KP, k = idx - c * KP;
    st[idx] = (k_base + k < K) ? s_hat[static_cast<size_t>(k_base + k) * C + c] : 0.f;
  }
  __syncthreads();

  const int items_per_image = (N + ITEM_PX - 1) / ITEM_PX;
  const long long n_items = static_cast<long long>(B) * items_per_image;
  // channel range of this warp (remainder channels go to the first warps)
  const int cbase = C / FG_WARPS, crem = C % FG_WARPS;
  const int c_lo = warp * cbase + min(warp, crem);
  const int c_hi = c_lo + cbase + (warp < crem ? 1 : 0);

  for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int b = static_cast<int>(item / items_per_image);
    const int n0 = static_cast<int>(item - static_cast<long long>(b) * items_per_image) * ITEM_PX;
    const int n = n0 + lane * PX;
    const bool live = n < N;                                   // N % PX == 0: all PX pixels or none
    const uint16_t* base = feat + (static_cast<size_t>(b) * C) * N + n;

    float2 acc[KP][PX / 2];
#pragma unroll
    for (int k = 0; k < KP; ++k)
#pragma unroll
      for (int j = 0; j < PX / 2; ++j) acc[k][j] = make_float2(0.f, 0.f);

    auto consume = [&](const typename PxVec<PX>::T& v, int c) {
      float2 x[PX / 2];
#pragma unroll
      for (int j = 0; j < PX / 2; ++j) {
        const uint32_t u = PxVec<PX>::word(v, j);
        x[j] = make_float2(bf16lo(u), bf16hi(u));
      }
      const float4* srow = reinterpret_cast<const float4*>(st + c * KP);
#pragma unroll
      for (int k4 = 0; k4 < KP / 4; ++k4) {
        const float4 s = srow[k4];
        const float sv[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const float2 s2 = make_float2(sv[kk], sv[kk]);
#pragma unroll
          for (int j = 0; j < PX / 2; ++j) ffma2(acc[4 * k4 + kk][j], s2, x[j]);
        }
      }
    };

    // (a ping-pong register double buffer was tried here: it spills at the 128-register budget that
    // two CTAs per SM allow and halves the throughput)
    using V = typename PxVec<PX>::T;
    int c = c_lo;
    for (; c + FG_PF <= c_hi; c += FG_PF) {
      V v[FG_PF];
#pragma unroll
      for (int u = 0; u < FG_PF; ++u)
        v[u] = live ? PxVec<PX>::load(base + static_cast<size_t>(c + u) * N) : PxVec<PX>::zero();
#pragma unroll
      for (int u = 0; u < FG_PF; ++u) consume(v[u], c + u);
    }
    for (; c < c_hi; ++c) {
      const V v = live ? PxVec<PX>::load(base + static_cast<size_t>(c) * N) : PxVec<PX>::zero();
      consume(v, c);
    }

    // 8-way reduction across the warps, then alpha/beta and the store
    __syncthreads();                                           // previous item's readers are done with `part`
#pragma unroll
    for (int k = 0; k < KP; ++k) {
      float* dst = part + (warp * KP + k) * ITEM_PX + lane * PX;
#pragma unroll
      for (int j = 0; j < PX / 4; ++j)
        *reinterpret_cast<float4*>(dst + 4 * j) =
            make_float4(acc[k][2 * j].x, acc[k][2 * j].y, acc[k][2 * j + 1].x, acc[k][2 * j + 1].y);
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < KP * ITEM_PX; idx += FG_THREADS) {
      const int k = idx / ITEM_PX, px = idx - k * ITEM_PX;
      const int kk = k_base + k;
      if (kk >= K || n0 + px >= N) continue;
      float p = 0.f;
#pragma unroll
      for (int w = 0; w < FG_WARPS; ++w) p += part[(w * KP + k) * ITEM_PX + px];
      const float v = p >= 0.f ? p * alpha_s[kk] : -p * beta_s[kk];
      logits[(static_cast<size_t>(b) * Ktot + ch_of[kk]) * N + n0 + px] = v;
    }
  }
}

template <int KP, int PX>
static int launch_fg(const uint16_t* feat, int B, int C, int N, const float* s_hat, const float* alpha,
                     const float* beta, int K, int k_base, float* logits, int Ktot, const ChMap& map,
                     cudaStream_t st) {
  const size_t smem = (static_cast<size_t>(C) * KP + static_cast<size_t>(FG_WARPS) * KP * PX * 32) * sizeof(float);
  auto kern = pop_fg_kernel<KP, PX>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  const long long items = static_cast<long long>(B) * ((N + PX * 32 - 1) / (PX * 32));
  const long long cap = static_cast<long long>(kNumSMs) * (KP <= 8 ? 2 : 1);
  const int grid = static_cast<int>(items < cap ? items : cap);
  kern<<<grid, FG_THREADS, smem, st>>>(feat, B, C, N, s_hat, alpha, beta, K, k_base, logits, Ktot, map);
  return SL_LAUNCH_RESULT();
}

template <int KP>
static int launch_fg_px(const uint16_t* feat, int B, int C, int N, const float* s_hat, const float* alpha,
                        const float* beta, int K, int k_base, float* logits, int Ktot, const ChMap& map,
                        cudaStream_t st) {
  // 256-pixel items when that still gives every SM work, 128-pixel items for small batches
  const long long items8 = static_cast<long long>(B) * ((N + 255) / 256);
  if (items8 >= 2 * kNumSMs)
    return launch_fg<KP, 8>(feat, B, C, N, s_hat, alpha, beta, K, k_base, logits, Ktot, map, st);
  return launch_fg<KP, 4>(feat, B, C, N, s_hat, alpha, beta, K, k_base, logits, Ktot, map, st);
}

}  // namespace sl

extern "C" int sl_pop_fg_lowres(const uint16_t* feat, int B, int C, int N, const float* s_hat, const float* alpha,
                                const float* beta, int K, float* logits, int Ktot, const int* ch_map_host,
                                void* stream) {
  SL_CHECK_PTR(feat); SL_CHECK_PTR(s_hat); SL_CHECK_PTR(alpha); SL_CHECK_PTR(beta); SL_CHECK_PTR(logits);
  SL_CHECK_PTR(ch_map_host);
  SL_CHECK_ARG(B >= 1 && K >= 1 && K < SL_MAX_CLASSES && Ktot >= K && Ktot <= SL_MAX_CLASSES);
  SL_CHECK_ARG(C >= 8 && C <= 1024 && C % 8 == 0 && N >= 8 && N % 8 == 0);
  SL_CHECK_ALIGN(feat, 16);
  sl::ChMap map;
  for (int k = 0; k < SL_MAX_CLASSES; ++k) map.ch[k] = 0;
  for (int k = 0; k < K; ++k) {
    SL_CHECK_ARG(ch_map_host[k] >= 0 && ch_map_host[k] < Ktot);
    map.ch[k] = ch_map_host[k];
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // Up to 12 classes per pass keep the accumulators in registers; more classes take extra passes.
  for (int k_base = 0; k_base < K; k_base += 12) {
    const int kc = (K - k_base) < 12 ? (K - k_base) : 12;
    int rc;
    if (kc <= 4) rc = sl::launch_fg_px<4>(feat, B, C, N, s_hat, alpha, beta, K, k_base, logits, Ktot, map, st);
    else if (kc <= 8) rc = sl::launch_fg_px<8>(feat, B, C, N, s_hat, alpha, beta, K, k_base, logits, Ktot, map, st);
    else rc = sl::launch_fg_px<12>(feat, B, C, N, s_hat, alpha, beta, K, k_base, logits, Ktot, map, st);
    if (rc != 0) return rc;
  }
  return SL_OK;
}
