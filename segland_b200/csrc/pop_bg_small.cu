// Background MLP for narrow heads (C <= 128: Swin-T/S/B's 96 / 128 channels; networks/swin_pop.py:182,
// backbones/swintransformer.py:490-506), precise mode.  Same maths and numerics as pop_bg_tc.cu
//     logit_0 = w3 . relu(W2 relu(W1' q)),   split-bf16: 2 + 3 MMA passes, fp32 accumulation in TMEM
// but a different data flow.  The general kernel streams a weight tile from L2 for every k-block of every pass
// of every 128-pixel tile; at C = 96 that is 280 KB of L2->shared traffic per 25 KB of features, and the kernel
// runs at the L2 bandwidth (14.9 us per 256x256 tile, 8x off both the HBM and the tensor roofline).  Here
//   * all four weight planes (W1' hi/lo, W2 hi/lo; <= 128 KB in UMMA layout) are loaded ONCE per CTA and stay
//     in shared memory,
//   * a feature tile is loaded once (one ring slot holds all its k-blocks) and serves both layer-1 passes,
//   * the hidden tile never leaves the SM: the layer-1 epilogue packs relu(.) as bf16 hi/lo pairs and writes them
//     with tcgen05.st into TENSOR MEMORY, and layer 2 takes its A operand from there (tcgen05.mma with a TMEM A
//     address), so its MMAs read only B from shared memory.
// The only memory traffic left is the features (HBM) and 4 bytes per pixel out.
// Roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-9 = layer-1 epilogue, warps 10-13 =
// layer-2 epilogue (separate warps, so the logit reduction of tile t overlaps the layer-1 epilogue of tile t+1:
// layer 1 of tile t+1 is issued before layer 2 of tile t and has its own pair of TMEM accumulators, so with one
// hidden-tile slot per SM the per-tile cycle is: hidden-tile stores -> layer-2 MMAs -> next tile's stores).
// SM resources at C = 128: 128 KB weights + 3 x 32 KB feature slots of shared memory; TMEM columns 0 / 128 (layer-1
// accumulators), 256 (layer 2), 384.. (hidden tile, C/2 packed columns per plane).
#include <cstdlib>
#include "tma.cuh"

namespace sl {
namespace tcs {
using namespace sl::tc;

constexpr int BLOCK_M = 128, BLOCK_K = 64, UMMA_K = 16;
constexpr int KB_TILE = BLOCK_M * BLOCK_K * 2;          // 16 KB: one k-block of a 128-row operand tile
constexpr int EPI_WARPS = 8;                           // layer-1 epilogue: TMEM -> relu -> bf16 hi/lo -> shared memory
constexpr int EPI2_WARPS = 4;                          // layer-2 epilogue: TMEM -> w3 . relu(.) -> logit
constexpr int THREADS = 64 + 32 * (EPI_WARPS + EPI2_WARPS);
constexpr int MAX_XS = 4;
constexpr int BAR_BYTES = 128;

// D[tmem] (+)= A[tmem] * B[smem]: the A operand (128 rows x 16 bf16 = 8 packed 32-bit columns) comes from tensor memory
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
// 32 lanes x 16 columns: thread = lane = row, 16 consecutive 32-bit columns
__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}

struct SmallParams {
  int C, kblocks, xs;            // channels, k-blocks (1 or 2), feature ring slots
  int m_tiles, tiles_per_image, N, Ktot, ch;
  const float* w3;
  float* logits;
  long long* dbg;                // optional per-tile timestamps of CTA 0 (profiling only)
};
struct SmallMaps { CUtensorMap x, w1h, w1l, w2h, w2l; };

// KB = k-blocks (64 channels each), KSL = 16-channel k-steps of the last one: C = 64 (KB - 1) + 16 KSL.  Compile-time so
// that the single MMA-issuing thread's loops unroll to one descriptor add per tcgen05.mma: at N = C <= 128 an MMA
// executes in ~50 cycles, and a runtime-indexed issue loop (~200 cycles per MMA for one thread) was the bottleneck.
template <int KB, int KSL>
__global__ void __launch_bounds__(THREADS, 1) bg_small_kernel(const __grid_constant__ SmallMaps maps, SmallParams p) {
  constexpr int C = 64 * (KB - 1) + 16 * KSL;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  if ((base & 1023u) != 0) asm volatile("trap;");
  constexpr int kblocks = KB;
  constexpr uint32_t w_plane = static_cast<uint32_t>(kblocks) * C * 128u;          // one weight plane: kblocks x [C rows x 128 B]
  // layout: weights [4 planes] | X ring [xs][kblocks][16 KB] | barriers.  The hidden tile lives in TENSOR MEMORY
  // (columns 384.. : bf16 pairs packed per 32-bit column, hi plane then lo plane), written by the layer-1 epilogue
  // with tcgen05.st and read by layer 2 as a TMEM A operand -- no shared-memory round trip, no proxy fence, and the
  // layer-2 MMAs do not compete with their own A reads for shared-memory bandwidth.
  const uint32_t w_base = base;
  const uint32_t x_base = (w_base + 4 * w_plane + 1023u) & ~1023u;
  constexpr uint32_t H_HI_COL = 384, H_LO_COL = 384 + 64;
  const uint32_t bar0 = x_base + static_cast<uint32_t>(p.xs) * kblocks * KB_TILE;
  uint8_t* gen = smem + (bar0 - base);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + 8 * (2 * MAX_XS + 8));
  auto xfull_bar = [&](int s) { return bar0 + 8u * s; };
  auto xempty_bar = [&](int s) { return bar0 + 8u * (MAX_XS + s); };
  // accumulators: layer 1 double-buffered (TMEM columns 0 / 128), layer 2 single (columns 256)
  auto tfull1_bar = [&](int b) { return bar0 + 8u * (2 * MAX_XS + b); };
  auto tempty1_bar = [&](int b) { return bar0 + 8u * (2 * MAX_XS + 2 + b); };
  const uint32_t tfull2_bar = bar0 + 8u * (2 * MAX_XS + 4);
  const uint32_t tempty2_bar = bar0 + 8u * (2 * MAX_XS + 5);
  const uint32_t w_bar = bar0 + 8u * (2 * MAX_XS + 6);
  const uint32_t h1_bar = bar0 + 8u * (2 * MAX_XS + 7);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto stamp = [&](int s, int slot_) { if (p.dbg != nullptr && blockIdx.x == 0 && s < 64) p.dbg[s * 16 + slot_] = clock64(); };
  const int n_my = (p.m_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  auto tile_of = [&](int s) { return static_cast<int>(blockIdx.x) + s * static_cast<int>(gridDim.x); };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.x); tma_prefetch_desc(&maps.w1h); tma_prefetch_desc(&maps.w1l);
    tma_prefetch_desc(&maps.w2h); tma_prefetch_desc(&maps.w2l);
    for (int s = 0; s < MAX_XS; ++s) { mbar_init(xfull_bar(s), 1); mbar_init(xempty_bar(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tfull1_bar(b), 1); mbar_init(tempty1_bar(b), EPI_WARPS); }
    mbar_init(tfull2_bar, 1); mbar_init(tempty2_bar, EPI2_WARPS);
    mbar_init(w_bar, 1);
    mbar_init(h1_bar, EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      // resident weights: plane order W1'hi, W1'lo, W2hi, W2lo; each k-block is a [C rows x 64 k] SWIZZLE_128B box
      mbar_expect_tx(w_bar, 4u * w_plane);
      const CUtensorMap* wm[4] = {&maps.w1h, &maps.w1l, &maps.w2h, &maps.w2l};
      for (int pl = 0; pl < 4; ++pl)
        for (int kb = 0; kb < kblocks; ++kb)
          tma_load_2d(w_base + pl * w_plane + kb * C * 128u, wm[pl], w_bar, kb * BLOCK_K, 0, L2_EVICT_LAST);
      int slot = 0; uint32_t phase = 0;
      for (int s = 0; s < n_my; ++s) {
        const int mt = tile_of(s);
        const int img = mt / p.tiles_per_image, n0 = (mt - img * p.tiles_per_image) * BLOCK_M;
        mbar_wait(xempty_bar(slot), phase ^ 1u);
        mbar_expect_tx(xfull_bar(slot), static_cast<uint32_t>(kblocks) * KB_TILE);
        for (int kb = 0; kb < kblocks; ++kb) {
          const uint32_t sa = x_base + (slot * kblocks + kb) * KB_TILE;
          tma_load_3d(sa, &maps.x, xfull_bar(slot), n0, kb * BLOCK_K, img, L2_EVICT_FIRST);
          tma_load_3d(sa + KB_TILE / 2, &maps.x, xfull_bar(slot), n0 + 64, kb * BLOCK_K, img, L2_EVICT_FIRST);
        }
        if (++slot == p.xs) { slot = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_g1 = make_idesc(BLOCK_M, C, true, 1u, 1u);
      constexpr uint32_t idesc_g2 = make_idesc(BLOCK_M, C, false, 1u, 1u);
      mbar_wait(w_bar, 0);
      // descriptor bases: the start-address field is the low 14 bits (>> 4), so stepping is an integer add
      const uint64_t db_w1 = make_desc(w_base, 16, 1024);
      int slot = 0; uint32_t phase = 0;
      // Issue order G1(0) | G1(1) G2(0) | G1(2) G2(1) | ...: layer 1 of the next tile runs on the tensor core while the
      // layer-1 epilogue of the current one converts its accumulators; layer 2 follows as soon as the hidden tile is in
      // shared memory.
      auto issue_g1 = [&](int s) {
        const int b = s & 1;
        stamp(s, 0);
        mbar_wait(tempty1_bar(b), static_cast<uint32_t>(((s >> 1) & 1) ^ 1));
        stamp(s, 1);
        mbar_wait(xfull_bar(slot), phase);
        stamp(s, 2);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(b * 128);
        const uint64_t da0 = make_desc(x_base + slot * kblocks * KB_TILE, KB_TILE / 2, 1024);   // + byte offset >> 4
#pragma unroll
        for (int pass = 0; pass < 2; ++pass)
#pragma unroll
          for (int kb = 0; kb < KB; ++kb)
#pragma unroll
            for (int k = 0; k < (kb == KB - 1 ? KSL : 4); ++k)
              tc_mma(d_tmem, da0 + ((kb * KB_TILE + k * (UMMA_K * 128)) >> 4),
                     db_w1 + ((pass * w_plane + kb * C * 128 + k * (UMMA_K * 2)) >> 4), idesc_g1, (pass | kb | k) ? 1u : 0u);
        tc_commit(xempty_bar(slot));                         // the feature slot is free once these MMAs retire
        tc_commit(tfull1_bar(b));
        if (++slot == p.xs) { slot = 0; phase ^= 1u; }
      };
      if (n_my > 0) issue_g1(0);
      for (int s = 0; s < n_my; ++s) {
        if (s + 1 < n_my) issue_g1(s + 1);
        // ---- layer 2: D = (H1hi + H1lo) W2hi^T + H1hi W2lo^T, A = the hidden tile the epilogue just wrote
        mbar_wait(tempty2_bar, static_cast<uint32_t>((s & 1) ^ 1));
        stamp(s, 3);
        mbar_wait(h1_bar, static_cast<uint32_t>(s & 1));
        stamp(s, 4);
        tc_fence_after();
#pragma unroll
        for (int pass = 0; pass < 3; ++pass)
#pragma unroll
          for (int kb = 0; kb < KB; ++kb)
#pragma unroll
            for (int k = 0; k < (kb == KB - 1 ? KSL : 4); ++k)
              tc_mma_ts(tmem_base + 256u, tmem_base + (pass == 1 ? H_LO_COL : H_HI_COL) + (kb * BLOCK_K + k * UMMA_K) / 2,
                        db_w1 + ((((pass == 2 ? 3 : 2) * w_plane) + kb * C * 128 + k * (UMMA_K * 2)) >> 4), idesc_g2,
                        (pass | kb | k) ? 1u : 0u);
        tc_commit(tfull2_bar);
        stamp(s, 5);
      }
    }
  } else if (warp < 2 + EPI_WARPS) {
    // ===================================================================== layer-1 epilogue (8 warps)
    // accumulators -> relu -> bf16 hi/lo (registers) -> shared memory (K-major SWIZZLE_128B tiles = layer 2's A
    // operand).  The conversion runs while layer 2 of the previous tile still reads the hidden tile; only the
    // shared-memory stores wait for it.
    const int sub = warp & 3;
    const int row = sub * 32 + lane;
    const int half = (warp - 2) >> 2;
    constexpr int n_chunks = C / 32;                         // <= 4: at most 2 chunks per warp
    constexpr int c_split = (n_chunks + 1) / 2;
    const int c_begin = half == 0 ? 0 : c_split, c_end = half == 0 ? c_split : n_chunks;
    for (int s = 0; s < n_my; ++s) {
      const int b = s & 1;
      mbar_wait(tfull1_bar(b), static_cast<uint32_t>((s >> 1) & 1));
      if (warp == 2 && lane == 0) stamp(s, 6);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(sub * 32) << 16) + static_cast<uint32_t>(b * 128);
      uint32_t hi[2][16], lo[2][16];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        if (c_begin + i < c_end) {
          uint32_t r[32];
          tc_ld32(taddr + (c_begin + i) * 32, r);
          tc_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float a = fmaxf(__uint_as_float(r[2 * j]), 0.f), bb = fmaxf(__uint_as_float(r[2 * j + 1]), 0.f);
            const __nv_bfloat162 h2 = __floats2bfloat162_rn(a, bb);
            const uint32_t hb = *reinterpret_cast<const uint32_t*>(&h2);
            const __nv_bfloat162 l2 = __floats2bfloat162_rn(a - __uint_as_float(hb << 16), bb - __uint_as_float(hb & 0xffff0000u));
            hi[i][j] = hb;
            lo[i][j] = *reinterpret_cast<const uint32_t*>(&l2);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty1_bar(b));            // accumulator drained: layer 1 of tile s+2 may overwrite it
      if (warp == 2 && lane == 0) stamp(s, 7);
      if (s > 0) mbar_wait(tfull2_bar, static_cast<uint32_t>((s - 1) & 1));   // layer 2 of tile s-1 has read the hidden tile
      if (warp == 2 && lane == 0) stamp(s, 8);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        if (c_begin + i < c_end) {
          const uint32_t col = static_cast<uint32_t>((c_begin + i) * 16);       // 32 channels = 16 packed columns
          const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(sub * 32) << 16);
          tc_st16(lane_addr + H_HI_COL + col, hi[i]);
          tc_st16(lane_addr + H_LO_COL + col, lo[i]);
        }
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(h1_bar);
    }
  } else {
    // ===================================================================== layer-2 epilogue (4 warps)
    const int sub = warp & 3;
    const int row = sub * 32 + lane;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(sub * 32) << 16) + 256u;
    for (int s = 0; s < n_my; ++s) {
      mbar_wait(tfull2_bar, static_cast<uint32_t>(s & 1));
      if (warp == 10 && lane == 0) stamp(s, 10);
      tc_fence_after();
      float logit = 0.f;
#pragma unroll
      for (int c0 = 0; c0 < C; c0 += 32) {
        uint32_t r[32];
        tc_ld32(taddr + c0, r);
        tc_ld_wait();
        const float4* wv = reinterpret_cast<const float4*>(p.w3 + c0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 w4 = __ldg(wv + j);
          logit = fmaf(w4.x, fmaxf(__uint_as_float(r[4 * j + 0]), 0.f), logit);
          logit = fmaf(w4.y, fmaxf(__uint_as_float(r[4 * j + 1]), 0.f), logit);
          logit = fmaf(w4.z, fmaxf(__uint_as_float(r[4 * j + 2]), 0.f), logit);
          logit = fmaf(w4.w, fmaxf(__uint_as_float(r[4 * j + 3]), 0.f), logit);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (warp == 10 && lane == 0) stamp(s, 11);
      if (lane == 0) mbar_arrive(tempty2_bar);
      const int mt = tile_of(s);
      const int img = mt / p.tiles_per_image;
      const int n = (mt - img * p.tiles_per_image) * BLOCK_M + row;
      if (n < p.N) p.logits[(static_cast<size_t>(img) * p.Ktot + p.ch) * p.N + n] = logit;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

}  // namespace tcs
}  // namespace sl

// Launch for C <= 128 (precise mode, background only).  Returns 0 on success.
int sl_pop_bg_small_launch(const uint16_t* feat, int B, int C, int N, const uint16_t* W1p_hi, const uint16_t* W1p_lo,
                           const uint16_t* W2_hi, const uint16_t* W2_lo, const float* w3_bg, float* logits, int Ktot, int ch,
                           cudaStream_t st) {
  using namespace sl::tcs;
  SmallParams p;
  p.C = C;
  p.kblocks = (C + BLOCK_K - 1) / BLOCK_K;
  p.tiles_per_image = (N + BLOCK_M - 1) / BLOCK_M;
  p.m_tiles = B * p.tiles_per_image;
  p.N = N; p.Ktot = Ktot; p.ch = ch; p.w3 = w3_bg; p.logits = logits;
  p.dbg = reinterpret_cast<long long*>(sl::env().small_dbg);   // SL_SMALL_DBG: device pointer of a 64 x 16 int64 buffer
  const size_t w_bytes = (static_cast<size_t>(4) * p.kblocks * C * 128 + 1023) / 1024 * 1024;
  const size_t h_bytes = 0;                                             // the hidden tile lives in tensor memory
  const size_t x_slot = static_cast<size_t>(p.kblocks) * KB_TILE;
  const size_t budget = 232448 - 1024 - BAR_BYTES;                      // minus alignment slack and barriers
  int xs = static_cast<int>((budget - w_bytes - h_bytes) / x_slot);
  if (xs < 1) return static_cast<int>(cudaErrorInvalidConfiguration);
  p.xs = xs > MAX_XS ? MAX_XS : xs;
  // at least 117 KB so that two CTAs can never share an SM: each allocates all 512 TMEM columns
  size_t smem = w_bytes + h_bytes + p.xs * x_slot + BAR_BYTES + 1024;
  if (smem < 117 * 1024) smem = 117 * 1024;
  SmallMaps m;
  int rc;
  {
    cuuint64_t dims[3] = {static_cast<cuuint64_t>(N), static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(B)};
    cuuint32_t box[3] = {64, BLOCK_K, 1};
    if ((rc = sl::tc::make_map(&m.x, feat, 3, dims, box))) return rc;
  }
  {
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(C)};
    cuuint32_t box[2] = {BLOCK_K, static_cast<cuuint32_t>(C)};
    if ((rc = sl::tc::make_map(&m.w1h, W1p_hi, 2, dims, box))) return rc;
    if ((rc = sl::tc::make_map(&m.w1l, W1p_lo, 2, dims, box))) return rc;
    if ((rc = sl::tc::make_map(&m.w2h, W2_hi, 2, dims, box))) return rc;
    if ((rc = sl::tc::make_map(&m.w2l, W2_lo, 2, dims, box))) return rc;
  }
  const int grid = p.m_tiles < sl::num_sms() ? p.m_tiles : sl::num_sms();
  auto launch = [&](auto kern) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
    kern<<<grid, THREADS, smem, st>>>(m, p);
    return SL_LAUNCH_RESULT();
  };
  switch (C) {
    case 32: return launch(bg_small_kernel<1, 2>);
    case 64: return launch(bg_small_kernel<1, 4>);
    case 96: return launch(bg_small_kernel<2, 2>);
    case 128: return launch(bg_small_kernel<2, 4>);
    default: return SL_EINVAL;
  }
}
