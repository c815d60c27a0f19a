// sl_pop_fg_lowres: the K foreground logits of the POP head at feature resolution.
//   reference: networks/pspnet_pop.py:108-109,114-115 (proj = s_hat @ q, out_fg = proj * s_hat)
//              followed by classifier / classifier_n on every rank-1 vector (:150-157, :178-182).
// Because the classifier is bias-free with ReLUs it is positively homogeneous, so
//   logit_k(pixel) = p >= 0 ? p * alpha_k : -p * beta_k,   p = s_hat_k . q(pixel)
// and the kernel is a skinny [K x C] x [C x N] contraction streamed once over the bf16 features:
// HBM-bound (C*N*2 bytes in, K*N*4 bytes out) as long as the FMA pipe keeps up; the inner product
// runs on packed fma.rn.f32x2 (FFMA2).
//
// One persistent CTA per SM, 16 autonomous warps.  A work item is 256 consecutive pixels of one image
// (all C channels) and belongs to ONE warp: lane l owns pixels 8l..8l+7, so there is no cross-warp
// reduction at all and the result is bit-reproducible.  Every warp runs its own TMA pipeline: lane 0
// issues cp.async.bulk.tensor loads of [8 channels x 256 pixels] (4 KB) boxes into a warp-private
// 3-stage shared-memory ring and the warp waits on the stage's mbarrier, so the bytes in flight
// (16 warps x 12 KB per SM) are independent of the register budget; the lanes then read the stage with
// conflict-free 128-bit shared loads.  Ragged edges (N % 256, C % 8) are zero-filled by TMA.
#include "tma.cuh"

namespace sl {

struct ChMap { int ch[SL_MAX_CLASSES]; };

constexpr int FG_WARPS = 16;
constexpr int FG_THREADS = FG_WARPS * 32;
constexpr int FG_CB = 8;                       // channels per stage
constexpr int FG_RING_BYTES = 12288;           // per-warp ring: 3 x 4 KB (8 px/lane) or 6 x 2 KB (4 px/lane)

// d = a * b + d on two packed fp32 lanes (sm_100 FFMA2)
__device__ __forceinline__ void ffma2(float2& d, const float2 a, const float2 b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;"
      : "+l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<const unsigned long long&>(a)), "l"(reinterpret_cast<const unsigned long long&>(b)));
}

// KC: classes handled by this launch (1..12).  PXL: pixels per lane -- 8 (items of 256 pixels, 128-bit
// shared loads) while the KC x 8 accumulators fit the 128-register budget of a 512-thread CTA, else 4.
template <int KC, int PXL>
__global__ void __launch_bounds__(FG_THREADS, 1)
pop_fg_kernel(const __grid_constant__ CUtensorMap map_x, int B, int C, int N, const float* __restrict__ s_hat,
              const float* __restrict__ alpha, const float* __restrict__ beta, int K, int k_base,
              float* __restrict__ logits, int Ktot, ChMap map) {
  constexpr int KP = (KC + 3) & ~3;            // prototype row padded for 128-bit shared loads
  constexpr int FG_PX = 32 * PXL;              // pixels per item
  constexpr int FG_STAGE_BYTES = FG_CB * FG_PX * 2;
  constexpr int FG_STAGES = FG_RING_BYTES / FG_STAGE_BYTES;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  // [ring: FG_WARPS x 12 KB][st: Cpad x KP fp32][barriers: FG_WARPS x 8]
  const int Cpad = (C + FG_CB - 1) / FG_CB * FG_CB;
  uint8_t* ring = smem_raw;
  float* st = reinterpret_cast<float*>(smem_raw + FG_WARPS * FG_RING_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(st + Cpad * KP);
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
  for (int idx = threadIdx.x; idx < Cpad * KP; idx += FG_THREADS) {
    const int c = idx / KP, k = idx - c * KP;
    st[idx] = (c < C && k < KC && k_base + k < K) ? s_hat[static_cast<size_t>(k_base + k) * C + c] : 0.f;
  }
  const uint32_t my_ring = tc::smem_u32(ring) + static_cast<uint32_t>(warp * FG_RING_BYTES);
  const uint32_t my_bars = tc::smem_u32(bars) + static_cast<uint32_t>(warp * 8 * 8);
  if (lane == 0) {
    for (int s = 0; s < FG_STAGES; ++s) tc::mbar_init(my_bars + 8u * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (warp == 0) tc::tma_prefetch_desc(&map_x);
  }
  __syncthreads();

  const int items_per_image = (N + FG_PX - 1) / FG_PX;
  const long long n_items = static_cast<long long>(B) * items_per_image;
  const int n_cb = Cpad / FG_CB;
  // item i -> CTA (i % grid), warp ((i / grid) % 16): consecutive items land on different SMs
  const long long first = static_cast<long long>(blockIdx.x) + static_cast<long long>(warp) * gridDim.x;
  const long long stride = static_cast<long long>(gridDim.x) * FG_WARPS;
  uint32_t phase_bits = 0;                                     // bit s = parity to wait for on stage s
  const uint8_t* my_ring_ptr = ring + warp * FG_RING_BYTES;

  for (long long item = first; item < n_items; item += stride) {
    const int b = static_cast<int>(item / items_per_image);
    const int n0 = static_cast<int>(item - static_cast<long long>(b) * items_per_image) * FG_PX;
    auto issue = [&](int cb) {                                 // lane 0 only
      const int s = cb % FG_STAGES;
      tc::mbar_expect_tx(my_bars + 8u * s, FG_STAGE_BYTES);
      tc::tma_load_3d(my_ring + static_cast<uint32_t>(s * FG_STAGE_BYTES), &map_x, my_bars + 8u * s, n0, cb * FG_CB, b,
                      tc::L2_EVICT_FIRST);
    };
    if (lane == 0)
      for (int cb = 0; cb < FG_STAGES - 1 && cb < n_cb; ++cb) issue(cb);

    float2 acc[KC][PXL / 2];
#pragma unroll
    for (int k = 0; k < KC; ++k)
#pragma unroll
      for (int j = 0; j < PXL / 2; ++j) acc[k][j] = make_float2(0.f, 0.f);

    for (int cb = 0; cb < n_cb; ++cb) {
      const int s = cb % FG_STAGES;
      // refill the stage consumed in the previous iteration (every lane finished reading it: __syncwarp below)
      if (lane == 0 && cb + FG_STAGES - 1 < n_cb) issue(cb + FG_STAGES - 1);
      tc::mbar_wait(my_bars + 8u * s, (phase_bits >> s) & 1u);
      phase_bits ^= 1u << s;
      const uint8_t* src = my_ring_ptr + s * FG_STAGE_BYTES + lane * (PXL * 2);
#pragma unroll
      for (int cc = 0; cc < FG_CB; ++cc) {
        float2 x[PXL / 2];
        if constexpr (PXL == 8) {
          const uint4 v = *reinterpret_cast<const uint4*>(src + cc * (FG_PX * 2));
          x[0] = make_float2(bf16lo(v.x), bf16hi(v.x)); x[1] = make_float2(bf16lo(v.y), bf16hi(v.y));
          x[2] = make_float2(bf16lo(v.z), bf16hi(v.z)); x[3] = make_float2(bf16lo(v.w), bf16hi(v.w));
        } else {
          const uint2 v = *reinterpret_cast<const uint2*>(src + cc * (FG_PX * 2));
          x[0] = make_float2(bf16lo(v.x), bf16hi(v.x)); x[1] = make_float2(bf16lo(v.y), bf16hi(v.y));
        }
        const float4* srow = reinterpret_cast<const float4*>(st + (cb * FG_CB + cc) * KP);
#pragma unroll
        for (int k4 = 0; k4 < KP / 4; ++k4) {
          const float4 sv4 = srow[k4];
          const float sv[4] = {sv4.x, sv4.y, sv4.z, sv4.w};
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            if (4 * k4 + kk < KC) {
              const float2 s2 = make_float2(sv[kk], sv[kk]);
#pragma unroll
              for (int j = 0; j < PXL / 2; ++j) ffma2(acc[4 * k4 + kk][j], s2, x[j]);
            }
          }
        }
      }
      __syncwarp();                                            // stage s may be overwritten from here on
    }

    const int n = n0 + lane * PXL;
    if (n < N) {                                               // N % 8 == 0: all PXL pixels or none
#pragma unroll
      for (int k = 0; k < KC; ++k) {
        const int kk = k_base + k;
        if (kk < K) {
          const float a = alpha_s[kk], bt = beta_s[kk];
          float o[PXL];
#pragma unroll
          for (int j = 0; j < PXL / 2; ++j) {
            const float p0 = acc[k][j].x, p1 = acc[k][j].y;
            o[2 * j] = p0 >= 0.f ? p0 * a : -p0 * bt;
            o[2 * j + 1] = p1 >= 0.f ? p1 * a : -p1 * bt;
          }
          float4* dst = reinterpret_cast<float4*>(logits + (static_cast<size_t>(b) * Ktot + ch_of[kk]) * N + n);
          dst[0] = make_float4(o[0], o[1], o[2], o[3]);
          if constexpr (PXL == 8) dst[1] = make_float4(o[4], o[5], o[6], o[7]);
        }
      }
    }
  }
}

template <int KC, int PXL>
static int launch_fg(const CUtensorMap& map_x, int B, int C, int N, const float* s_hat, const float* alpha,
                     const float* beta, int K, int k_base, float* logits, int Ktot, const ChMap& map,
                     cudaStream_t st) {
  constexpr int KP = (KC + 3) & ~3;
  const int Cpad = (C + FG_CB - 1) / FG_CB * FG_CB;
  const size_t smem = static_cast<size_t>(FG_WARPS) * FG_RING_BYTES + static_cast<size_t>(Cpad) * KP * 4 +
                      static_cast<size_t>(FG_WARPS) * 8 * 8;
  auto kern = pop_fg_kernel<KC, PXL>;
  constexpr int FG_PX = 32 * PXL;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  const long long items = static_cast<long long>(B) * ((N + FG_PX - 1) / FG_PX);
  const int grid = static_cast<int>(items < num_sms() ? items : num_sms());
  kern<<<grid, FG_THREADS, smem, st>>>(map_x, B, C, N, s_hat, alpha, beta, K, k_base, logits, Ktot, map);
  return SL_LAUNCH_RESULT();
}

int launch_fg_mma(const uint16_t* feat, int B, int C, int N, const float* s_hat, const float* alpha, const float* beta,
                  int K, int k_base, float* logits, int Ktot, const int* ch_map_host, cudaStream_t st);   // pop_fg_mma.cu

}  // namespace sl

extern "C" int sl_pop_fg_lowres(const uint16_t* feat, int B, int C, int N, const float* s_hat, const float* alpha,
                                const float* beta, int K, float* logits, int Ktot, const int* ch_map_host,
                                void* stream) {
  SL_CHECK_PTR(feat); SL_CHECK_PTR(s_hat); SL_CHECK_PTR(alpha); SL_CHECK_PTR(beta); SL_CHECK_PTR(logits);
  SL_CHECK_PTR(ch_map_host);
  SL_CHECK_ARG(B >= 1 && K >= 1 && K < SL_MAX_CLASSES && Ktot >= K && Ktot <= SL_MAX_CLASSES);
  SL_CHECK_ARG(C >= 8 && C <= 512 && C % 8 == 0 && N >= 8 && N % 8 == 0);   // prototypes + rings fill 222 KB at C = 512
  SL_CHECK_ALIGN(feat, 16); SL_CHECK_ALIGN(logits, 16);
  sl::ChMap map;
  for (int k = 0; k < SL_MAX_CLASSES; ++k) map.ch[k] = 0;
  for (int k = 0; k < K; ++k) {
    SL_CHECK_ARG(ch_map_host[k] >= 0 && ch_map_host[k] < Ktot);
    map.ch[k] = ch_map_host[k];
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // Default: the projections run on the legacy tensor path, 16 classes per pass (pop_fg_mma.cu): 88-102 % of the HBM
  // copy peak at K = 7 and K = 11 against 92 % / 66-70 % for the FFMA2 kernel below (profiles/r2_fg_mma_probe.txt),
  // which stays as the SL_FG_MMA=0 comparison.
  const int mma_env = sl::env().fg_mma;
  if (mma_env != 0) {
    for (int k_base = 0; k_base < K; k_base += 16) {
      const int rc = sl::launch_fg_mma(feat, B, C, N, s_hat, alpha, beta, K, k_base, logits, Ktot, ch_map_host, st);
      if (rc != 0) return rc;
    }
    return SL_OK;
  }
  // Up to 12 classes per pass keep the accumulators in registers; more classes take extra passes.
  for (int k_base = 0; k_base < K; k_base += 12) {
    const int kc = (K - k_base) < 12 ? (K - k_base) : 12;
    const int pxl = kc <= 8 ? 8 : 4;
    CUtensorMap map_x;   // features [B][C][N] bf16: box = 32*pxl pixels x 8 channels, no swizzle (rows are read linearly)
    {
      cuuint64_t dims[3] = {static_cast<cuuint64_t>(N), static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(B)};
      cuuint32_t box[3] = {static_cast<cuuint32_t>(32 * pxl), sl::FG_CB, 1};
      const int rcm = sl::tc::make_map(&map_x, feat, 3, dims, box, CU_TENSOR_MAP_SWIZZLE_NONE);
      if (rcm) return rcm;
    }
    int rc = SL_EINVAL;
#define SL_FG_CASE(KC, PXL) case KC: rc = sl::launch_fg<KC, PXL>(map_x, B, C, N, s_hat, alpha, beta, K, k_base, logits, Ktot, map, st); break
    switch (kc) {
      SL_FG_CASE(1, 8); SL_FG_CASE(2, 8); SL_FG_CASE(3, 8); SL_FG_CASE(4, 8); SL_FG_CASE(5, 8); SL_FG_CASE(6, 8);
      SL_FG_CASE(7, 8); SL_FG_CASE(8, 8); SL_FG_CASE(9, 4); SL_FG_CASE(10, 4); SL_FG_CASE(11, 4); SL_FG_CASE(12, 4);
    }
#undef SL_FG_CASE
    if (rc != 0) return rc;
  }
  return SL_OK;
}
