// sl_pop_fg_lowres for more than 8 classes (ft mode: K = 11 on OEM): the K foreground projections
//   p_k = s_hat_k . q   (networks/pspnet_pop.py:108-109,114-115),  logit_k = p_k >= 0 ? p_k*alpha_k : -p_k*beta_k
// on the legacy tensor path (mma.sync m16n8k16, SASS HMMA) instead of FFMA2.
//
// Why: the CUDA-core kernel (pop_fg.cu) spends K FMAs per feature element; at K = 11 that is 5.5 packed FMAs per
// 2-byte element and the FMA pipe, not HBM, is the limit (66-70 % of the copy peak, ncu: math_pipe_throttle).  The
// contraction is [16 classes x 16 channels] x [16 channels x 8 pixels] per instruction here, ~8x fewer issue slots,
// so the kernel is HBM-bound again.  tcgen05 is not the tool: M = 16 classes would waste 7/8 of a 128-row UMMA tile
// and the work is one pass over the features with nothing to amortise a TMEM round trip.
//
// Precision: the features are exactly bf16; the fp32 prototypes are split into three bf16 terms (hi + mid + lo, 24
// mantissa bits) and the three products accumulate in fp32, so the result matches the fp32 FMA kernel to ~1e-7 relative
// (summation order differs).  Deterministic: one warp owns an item (128 pixels, all channels), no atomics.
//
// Layout: as pop_fg.cu -- one persistent CTA per SM, 16 autonomous warps, each with a private TMA ring -- but a stage
// is [16 channels][128 pixels] as two 64-pixel SWIZZLE_128B blocks, read with ldmatrix.trans (the B fragment wants the
// two channels of a pair in one register; the NCHW tile has pixels contiguous), conflict-free thanks to the swizzle.
// The A fragments (prototype splits) are built once per CTA in shared memory in register order.
#include "tma.cuh"

namespace sl {

struct ChMap32 { int ch[SL_MAX_CLASSES]; };

constexpr int FM_WARPS = 16;
constexpr int FM_THREADS = FM_WARPS * 32;
constexpr int FM_PX = 128;                      // pixels per item
constexpr int FM_KS = 16;                       // channels per k-step
constexpr int FM_STAGE_BYTES = FM_KS * FM_PX * 2;   // 4 KB: two [16][64] blocks of 2 KB
constexpr int FM_STAGES = 2;                    // per warp when all 16 warps of a CTA have items
constexpr int FM_RING_BYTES = FM_STAGES * FM_STAGE_BYTES;
constexpr int FM_MAX_STAGES = 32;               // one warp with items (a single tile: 128 items on 148 SMs) takes the whole ring
constexpr int FM_SPLITS = 3;

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint4& a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  return static_cast<uint32_t>(f32_to_bf16_rn(lo)) | (static_cast<uint32_t>(f32_to_bf16_rn(hi)) << 16);
}

__global__ void __launch_bounds__(FM_THREADS, 1)
pop_fg_mma_kernel(const __grid_constant__ CUtensorMap map_x, int B, int C, int N, const float* __restrict__ s_hat,
                  const float* __restrict__ alpha, const float* __restrict__ beta, int K, int k_base,
                  float* __restrict__ logits, int Ktot, ChMap32 map) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // [ring: FM_WARPS x 8 KB][A fragments: ksteps x 3 splits x 32 lanes x 16 B][barriers]
  const int ksteps = (C + FM_KS - 1) / FM_KS;
  uint8_t* ring = smem_raw;
  uint4* afrag = reinterpret_cast<uint4*>(smem_raw + FM_WARPS * FM_RING_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(afrag + static_cast<size_t>(ksteps) * FM_SPLITS * 32);
  __shared__ int ch_of[16];
  __shared__ float alpha_s[16], beta_s[16];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if ((tc::smem_u32(smem_raw) & 1023u) != 0) asm volatile("trap;");      // SWIZZLE_128B wants 1024-byte aligned blocks

  if (threadIdx.x < 16) {
    const int kk = k_base + static_cast<int>(threadIdx.x);
    int ch = 0;
#pragma unroll
    for (int k = 0; k < SL_MAX_CLASSES; ++k) ch = (k == kk) ? map.ch[k] : ch;
    ch_of[threadIdx.x] = ch;
    alpha_s[threadIdx.x] = kk < K ? alpha[kk] : 0.f;
    beta_s[threadIdx.x] = kk < K ? beta[kk] : 0.f;
  }
  // A fragments in mma register order: lane l, register r holds A[row][col], A[row][col + 1] with
  //   row = l/4 + 8*(r & 1), col = 2*(l % 4) + 8*(r >> 1); rows = classes k_base.., cols = channels of the k-step
  for (int idx = threadIdx.x; idx < ksteps * 32; idx += FM_THREADS) {
    const int ks = idx >> 5, l = idx & 31;
    uint32_t w[FM_SPLITS][4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int row = (l >> 2) + 8 * (r & 1), col = 2 * (l & 3) + 8 * (r >> 1);
      const int kk = k_base + row;
      float v[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = ks * FM_KS + col + e;
        v[e] = (kk < K && c < C) ? s_hat[static_cast<size_t>(kk) * C + c] : 0.f;
      }
      float rem0 = v[0], rem1 = v[1];
#pragma unroll
      for (int s = 0; s < FM_SPLITS; ++s) {
        const uint32_t p = pack_bf16(rem0, rem1);
        w[s][r] = p;
        rem0 -= bf16lo(p);                                       // exact: the residual of a rounding is representable
        rem1 -= bf16hi(p);
      }
    }
#pragma unroll
    for (int s = 0; s < FM_SPLITS; ++s)
      afrag[(static_cast<size_t>(ks) * FM_SPLITS + s) * 32 + l] = make_uint4(w[s][0], w[s][1], w[s][2], w[s][3]);
  }
  // Small batches: only the first `active` warps of a CTA have items (item i belongs to warp (i / gridDim.x) of CTA
  // i % gridDim.x), and one warp streaming its 131 KB through two 4 KB stages is latency-bound (18 us for one tile).
  // The 128 KB of ring space are therefore divided among the warps that have work: `stages` = 2 x 16 / active (a power of
  // two, up to 32), same loads, same order, same arithmetic -- only more of them in flight.
  const int items_per_image = (N + FM_PX - 1) / FM_PX;
  const long long n_items = static_cast<long long>(B) * items_per_image;
  int active = static_cast<int>((n_items + gridDim.x - 1) / gridDim.x);
  active = active >= 16 ? 16 : active > 8 ? 16 : active > 4 ? 8 : active > 2 ? 4 : active > 1 ? 2 : 1;
  const int stages = FM_STAGES * (FM_WARPS / active);
  const uint32_t my_ring = tc::smem_u32(ring) + static_cast<uint32_t>(warp * stages * FM_STAGE_BYTES);
  const uint32_t my_bars = tc::smem_u32(bars) + static_cast<uint32_t>(warp * 8 * stages);
  if (lane == 0 && warp < active) {
    for (int s = 0; s < stages; ++s) tc::mbar_init(my_bars + 8u * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (warp == 0) tc::tma_prefetch_desc(&map_x);
  }
  __syncthreads();

  const long long first = static_cast<long long>(blockIdx.x) + static_cast<long long>(warp) * gridDim.x;
  const long long stride = static_cast<long long>(gridDim.x) * FM_WARPS;
  uint32_t phase_bits = 0;
  // ldmatrix row address of this lane inside a 2 KB block: matrix = lane / 8 -> (k half = matrix & 1, chunk + (matrix >> 1))
  const int lm_row = (lane & 7) + 8 * ((lane >> 3) & 1);         // channel row 0..15 inside the k-step
  const int lm_chunk_add = lane >> 4;                            // second pair of matrices: the next 8-pixel chunk

  for (long long item = first; item < n_items; item += stride) {
    const int b = static_cast<int>(item / items_per_image);
    const int n0 = static_cast<int>(item - static_cast<long long>(b) * items_per_image) * FM_PX;
    auto issue = [&](int ks) {                                   // lane 0 only
      const int s = ks & (stages - 1);
      const uint32_t dst = my_ring + static_cast<uint32_t>(s * FM_STAGE_BYTES);
      tc::mbar_expect_tx(my_bars + 8u * s, FM_STAGE_BYTES);
      tc::tma_load_3d(dst, &map_x, my_bars + 8u * s, n0, ks * FM_KS, b, tc::L2_EVICT_FIRST);
      tc::tma_load_3d(dst + FM_STAGE_BYTES / 2, &map_x, my_bars + 8u * s, n0 + 64, ks * FM_KS, b, tc::L2_EVICT_FIRST);
    };
    if (lane == 0)
      for (int ks = 0; ks < stages - 1 && ks < ksteps; ++ks) issue(ks);

    float acc[FM_PX / 8][4];
#pragma unroll
    for (int t = 0; t < FM_PX / 8; ++t)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[t][j] = 0.f;

    for (int ks = 0; ks < ksteps; ++ks) {
      const int s = ks & (stages - 1);
      if (lane == 0 && ks + stages - 1 < ksteps) issue(ks + stages - 1);
      tc::mbar_wait(my_bars + 8u * s, (phase_bits >> s) & 1u);
      phase_bits ^= 1u << s;
      const uint4 a0 = afrag[(static_cast<size_t>(ks) * FM_SPLITS + 0) * 32 + lane];
      const uint4 a1 = afrag[(static_cast<size_t>(ks) * FM_SPLITS + 1) * 32 + lane];
      const uint4 a2 = afrag[(static_cast<size_t>(ks) * FM_SPLITS + 2) * 32 + lane];
      const uint32_t stage = my_ring + static_cast<uint32_t>(s * FM_STAGE_BYTES);
#pragma unroll
      for (int tp = 0; tp < FM_PX / 16; ++tp) {                  // pairs of 8-pixel n-tiles
        const int blk = tp >> 2;                                 // 64-pixel block
        const int chunk = ((tp & 3) << 1) + lm_chunk_add;        // 16-byte chunk (8 pixels) inside the 128-byte row
        const uint32_t addr = stage + static_cast<uint32_t>(blk * (FM_STAGE_BYTES / 2) + lm_row * 128 +
                                                             ((chunk ^ (lm_row & 7)) << 4));
        uint32_t b0, b1, b2, b3;
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                     : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3) : "r"(addr));
        // smallest terms first: (lo, mid, hi) so the large product is added last
        mma_bf16_16816(acc[2 * tp], a2, b0, b1);
        mma_bf16_16816(acc[2 * tp], a1, b0, b1);
        mma_bf16_16816(acc[2 * tp], a0, b0, b1);
        mma_bf16_16816(acc[2 * tp + 1], a2, b2, b3);
        mma_bf16_16816(acc[2 * tp + 1], a1, b2, b3);
        mma_bf16_16816(acc[2 * tp + 1], a0, b2, b3);
      }
      __syncwarp();                                              // stage s may be overwritten from here on
    }

    // C fragment: d0,d1 = class lane/4, pixels 8t + 2*(lane%4) + {0,1}; d2,d3 = class lane/4 + 8
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int row = (lane >> 2) + 8 * half;
      if (k_base + row < K) {
        const float a = alpha_s[row], bt = beta_s[row];
        float* dst = logits + (static_cast<size_t>(b) * Ktot + ch_of[row]) * N + n0 + 2 * (lane & 3);
#pragma unroll
        for (int t = 0; t < FM_PX / 8; ++t) {
          if (n0 + 8 * t + 2 * (lane & 3) < N) {                 // N % 8 == 0: both pixels of the pair or neither
            const float p0 = acc[t][2 * half], p1 = acc[t][2 * half + 1];
            *reinterpret_cast<float2*>(dst + 8 * t) = make_float2(p0 >= 0.f ? p0 * a : -p0 * bt, p1 >= 0.f ? p1 * a : -p1 * bt);
          }
        }
      }
    }
  }
}

// One pass of up to 16 classes starting at k_base.  Returns SL_EINVAL when the shape is outside the kernel's range.
int launch_fg_mma(const uint16_t* feat, int B, int C, int N, const float* s_hat, const float* alpha, const float* beta,
                  int K, int k_base, float* logits, int Ktot, const int* ch_map_host, cudaStream_t st) {
  if (C < 8 || C > 512 || C % 8 != 0 || N < 8 || N % 8 != 0) return SL_EINVAL;
  ChMap32 map;
  for (int k = 0; k < SL_MAX_CLASSES; ++k) map.ch[k] = k < K ? ch_map_host[k] : 0;
  CUtensorMap map_x;   // features [B][C][N] bf16: box = 64 pixels x 16 channels, SWIZZLE_128B
  {
    cuuint64_t dims[3] = {static_cast<cuuint64_t>(N), static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(B)};
    cuuint32_t box[3] = {64, FM_KS, 1};
    const int rcm = tc::make_map(&map_x, feat, 3, dims, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rcm) return rcm;
  }
  const int ksteps = (C + FM_KS - 1) / FM_KS;
  const size_t smem = static_cast<size_t>(FM_WARPS) * FM_RING_BYTES + static_cast<size_t>(ksteps) * FM_SPLITS * 32 * 16 +
                      static_cast<size_t>(FM_WARPS) * 8 * FM_STAGES;
  cudaError_t e = cudaFuncSetAttribute(pop_fg_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  const long long items = static_cast<long long>(B) * ((N + FM_PX - 1) / FM_PX);
  const int grid = static_cast<int>(items < num_sms() ? items : num_sms());
  pop_fg_mma_kernel<<<grid, FM_THREADS, smem, st>>>(map_x, B, C, N, s_hat, alpha, beta, K, k_base, logits, Ktot, map);
  return SL_LAUNCH_RESULT();
}

}  // namespace sl
