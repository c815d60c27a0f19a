// sl_pop_bg_tc: background (class 0) logit of the POP head on tcgen05 tensor cores.
//   reference: networks/pspnet_pop.py:112,118 (out_bg) through classifier / classifier_n
//   (:154-157, :178-182):   logit_0 = w3 . relu(W2 relu(W1' q)),  W1' = W1 (I - S^T S).
//
// Precision: plain bf16 operands miss the 1e-3 parity bound (3e-3 measured), so both layers run
// as split-bf16 products with fp32 accumulation in TMEM:
//   layer 1:  q is exactly bf16          ->  q.W1'hi + q.W1'lo                         (2 passes)
//   layer 2:  h = relu(.) = h_hi + h_lo  ->  h_hi.W2hi + h_lo.W2hi + h_hi.W2lo         (3 passes)
// which lands within ~5e-6 of the fp32 reference.
//
// Two launches of one warp-specialised persistent GEMM kernel (sm_100a):
//   MODE 1  H1[pixels, C] = relu(X^T W1'^T): A = feature tile straight from the NCHW tensor
//           (MN-major UMMA operand, pixels contiguous), epilogue splits h into bf16 hi/lo and
//           writes them K-major for layer 2.
//   MODE 2  logit[pixel] = sum_n w3[n] relu(H1 W2^T)[pixel, n]: epilogue reduces over the n-tile
//           in registers; one CTA owns every n-tile of its 128-pixel tile so no atomics.
// Roles: warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread tcgen05.mma issuer,
// warps 2.. = epilogue (tcgen05.ld, one TMEM lane = one pixel per thread).  128 x NT fp32
// accumulators are double-buffered in TMEM (2 x 256 columns) so the epilogue of one n-tile
// overlaps the MMAs of the next.  Operands arrive by TMA (SWIZZLE_128B) through a 4-stage
// mbarrier ring: 16 KB (A) + up to 32 KB (B) per stage.
#include <cuda.h>
#include <cstdlib>
#include "common.cuh"

namespace sl {
namespace tc {

constexpr int BLOCK_M = 128;     // pixels per tile (= TMEM lanes)
constexpr int BLOCK_K = 64;      // bf16 elements per k-block = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int STAGES = 4;
constexpr int MAX_NT = 256;      // n-tile width (UMMA N), C or C/2
constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;          // 16 KB
constexpr int B_BYTES_MAX = MAX_NT * BLOCK_K * 2;       // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES_MAX;
constexpr int TMEM_COLS = 512;
// warps: 0 = TMA producer, 1 = MMA issuer, then the epilogue warps.  Layer 1's epilogue converts and
// stores 128 x NT x {hi,lo} bf16 per n-tile and is the critical path there, so it gets 8 warps (two per
// TMEM sub-partition, each taking half of the columns); layer 2's epilogue is a dot product: 4 warps.
template <int MODE> struct Cfg {
  static constexpr int EPI_WARPS = MODE == 1 ? 8 : 4;
  static constexpr int THREADS = 64 + 32 * EPI_WARPS;
};
// After the operand ring: a 32 KB region that holds w3 (MODE 2) or the epilogue's TMA-store staging
// (MODE 1: 8 warps x {hi,lo} x [32 rows][64 B]), then the barriers.
constexpr int AUX_BYTES = 32768;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + AUX_BYTES + 256 /*barriers*/;

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// smem tile -> global (bulk async group); the source must be visible to the async proxy first.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 x bf16 -> fp32, issued by one thread for the CTA.
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
// 32 lanes x 32 columns of fp32 accumulators -> 32 registers per thread (thread = lane = row).
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout), SWIZZLE_128B:
//   [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout=2
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return static_cast<uint64_t>((saddr & 0x3ffff) >> 4) | (static_cast<uint64_t>(lbo_bytes >> 4) << 16) |
         (static_cast<uint64_t>(sbo_bytes >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): fp32 accumulate, bf16 A/B.
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, bool a_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn_major ? 1u : 0u) << 15) | (0u << 16) |
         (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

struct Params {
  int C;            // channels (K of both layers and N of both layers)
  int NT;           // n-tile width
  int n_tiles;      // C / NT
  int m_tiles;      // B*N / 128
  int tiles_per_image;  // N / 128
  int N;            // pixels per image
  int Ktot, ch;     // logits layout
  uint16_t* h_hi;   // MODE 1 out: [B*N][C] bf16
  uint16_t* h_lo;
  const float* w3;  // MODE 2
  float* logits;    // MODE 2 out
  int debug;        // SL_TC_DEBUG experiments: 1 = no epilogue stores, 2 = no epilogue work at all
};

template <int MODE>
__global__ void __launch_bounds__(Cfg<MODE>::THREADS, 1)
bg_gemm_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
               const __grid_constant__ CUtensorMap map_b0, const __grid_constant__ CUtensorMap map_b1,
               const __grid_constant__ CUtensorMap map_st0, const __grid_constant__ CUtensorMap map_st1, Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;            // SWIZZLE_128B wants 1024-byte tiles
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  float* w3s = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);     // [512]  (MODE 2)
  const uint32_t stage_out = base + STAGES * STAGE_BYTES;                 // store staging (MODE 1)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + AUX_BYTES);
  // bars: [0,S) full, [S,2S) empty, [2S,2S+2) tmem_full, [2S+2,2S+4) tmem_empty; then the TMEM base slot
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bar0 + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bar0 + 8u * (2 * STAGES + 2 + s); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int PASSES = MODE == 1 ? 2 : 3;
  const int kblocks = p.C / BLOCK_K;
  const uint32_t b_bytes = static_cast<uint32_t>(p.NT) * BLOCK_K * 2;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a0); tma_prefetch_desc(&map_b0); tma_prefetch_desc(&map_b1);
    if (MODE == 2) tma_prefetch_desc(&map_a1);
    if (MODE == 1) { tma_prefetch_desc(&map_st0); tma_prefetch_desc(&map_st1); }
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), Cfg<MODE>::EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (MODE == 2)
    for (int i = threadIdx.x; i < p.C; i += Cfg<MODE>::THREADS) w3s[i] = p.w3[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int mt = blockIdx.x; mt < p.m_tiles; mt += gridDim.x) {
        for (int nt = 0; nt < p.n_tiles; ++nt) {
          for (int pass = 0; pass < PASSES; ++pass) {
            const CUtensorMap* ma = (MODE == 2 && pass == 1) ? &map_a1 : &map_a0;
            const CUtensorMap* mb = (MODE == 1 ? pass == 1 : pass == 2) ? &map_b1 : &map_b0;
            for (int kb = 0; kb < kblocks; ++kb) {
              mbar_wait(empty_bar(stage), phase ^ 1u);
              const uint32_t sa = base + stage * STAGE_BYTES, sb = sa + A_BYTES;
              mbar_expect_tx(full_bar(stage), A_BYTES + b_bytes);
              if (MODE == 1) {
                // feature tile: 64 channels x 128 pixels as two 64-pixel (128-byte) column blocks
                const int img = mt / p.tiles_per_image, n0 = (mt - img * p.tiles_per_image) * BLOCK_M;
                tma_load_3d(sa, ma, full_bar(stage), n0, kb * BLOCK_K, img);
                tma_load_3d(sa + A_BYTES / 2, ma, full_bar(stage), n0 + 64, kb * BLOCK_K, img);
              } else {
                tma_load_2d(sa, ma, full_bar(stage), kb * BLOCK_K, mt * BLOCK_M);
              }
              tma_load_2d(sb, mb, full_bar(stage), kb * BLOCK_K, nt * p.NT);
              if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc(BLOCK_M, p.NT, MODE == 1);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int mt = blockIdx.x; mt < p.m_tiles; mt += gridDim.x) {
        for (int nt = 0; nt < p.n_tiles; ++nt) {
          mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * MAX_NT);
          uint32_t accumulate = 0;
          for (int it = 0; it < PASSES * kblocks; ++it) {
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            const uint32_t sa = base + stage * STAGE_BYTES, sb = sa + A_BYTES;
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              uint64_t da, db;
              if (MODE == 1)   // MN-major: 16 k-rows of 128 B = 2 KB per UMMA_K; 64-pixel blocks 8 KB apart
                da = make_desc(sa + k * (UMMA_K * 128), A_BYTES / 2, 1024);
              else             // K-major: 32 bytes along the 128-byte row per UMMA_K
                da = make_desc(sa + k * (UMMA_K * 2), 16, 1024);
              db = make_desc(sb + k * (UMMA_K * 2), 16, 1024);
              tc_mma(d_tmem, da, db, idesc, accumulate);
              accumulate = 1;
            }
            tc_commit(empty_bar(stage));          // frees the smem stage once these MMAs retire
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          }
          tc_commit(tfull_bar(acc));              // accumulator complete -> epilogue
          if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
      }
    }
  } else {
    // ===================================================================== epilogue
    const int sub = warp & 3;                     // TMEM sub-partition this warp may read
    const int row = sub * 32 + lane;              // row of the 128-pixel tile
    const int half = (warp - 2) >> 2;             // MODE 1: which half of the n-tile's 32-column chunks
    const int n_chunks = p.NT / 32;
    const int c_begin = Cfg<MODE>::EPI_WARPS == 8 ? (half == 0 ? 0 : (n_chunks + 1) / 2) : 0;
    const int c_end = Cfg<MODE>::EPI_WARPS == 8 ? (half == 0 ? (n_chunks + 1) / 2 : n_chunks) : n_chunks;
    const uint32_t sbuf = stage_out + static_cast<uint32_t>((warp - 2) * 4096);
    int acc = 0; uint32_t acc_phase = 0;
    for (int mt = blockIdx.x; mt < p.m_tiles; mt += gridDim.x) {
      float logit = 0.f;
      for (int nt = 0; nt < p.n_tiles; ++nt) {
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(sub * 32) << 16) + static_cast<uint32_t>(acc * MAX_NT);
        for (int ch = c_begin; ch < c_end; ++ch) {
          if (p.debug & 2) break;
          const int c0 = ch * 32;
          uint32_t r[32];
          tc_ld32(taddr + c0, r);
          tc_ld_wait();
          if (MODE == 1) {
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float a = fmaxf(__uint_as_float(r[2 * j]), 0.f), b = fmaxf(__uint_as_float(r[2 * j + 1]), 0.f);
              const __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);          // one cvt.rn.bf16x2.f32
              const uint32_t hb = *reinterpret_cast<const uint32_t*>(&h2);
              const __nv_bfloat162 l2 = __floats2bfloat162_rn(a - __uint_as_float(hb << 16),
                                                              b - __uint_as_float(hb & 0xffff0000u));
              hi[j] = hb;
              lo[j] = *reinterpret_cast<const uint32_t*>(&l2);
            }
            // stage the 32x32 bf16 chunk (64 B per row, SWIZZLE_64B: 16-byte chunk ^= (row>>1)&3) and
            // hand it to TMA: the store engine writes full lines, the LSU only sees shared memory.
            if (p.debug & 1) continue;
            if (lane == 0) tma_store_wait_read<0>();          // the previous store has drained this buffer
            __syncwarp();
            const uint32_t rbase = sbuf + static_cast<uint32_t>(lane) * 64u;
            const uint32_t sw = static_cast<uint32_t>((lane >> 1) & 3);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint32_t off16 = ((static_cast<uint32_t>(q) ^ sw) << 4);
              st_shared_v4(rbase + off16, hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]);
              st_shared_v4(rbase + 2048u + off16, lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
              const int col = nt * p.NT + c0;
              const int row0 = mt * BLOCK_M + sub * 32;
              tma_store_2d(&map_st0, sbuf, col, row0);
              tma_store_2d(&map_st1, sbuf + 2048u, col, row0);
              tma_store_commit();
            }
          } else {
            const float* wv = w3s + nt * p.NT + c0;
#pragma unroll
            for (int j = 0; j < 32; ++j) logit = fmaf(wv[j], fmaxf(__uint_as_float(r[j]), 0.f), logit);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(acc));   // one arrival per epilogue warp
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
      if (MODE == 2) {
        const int img = mt / p.tiles_per_image;
        const int n = (mt - img * p.tiles_per_image) * BLOCK_M + row;
        p.logits[(static_cast<size_t>(img) * p.Ktot + p.ch) * p.N + n] = logit;
      }
    }
    if (MODE == 1) {
      if (lane == 0) tma_store_wait_read<0>();         // staging must outlive the CTA's last stores
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------ host: tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;           // benign race: every thread resolves the same pointer
  if (fn == nullptr) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

// bf16 tensor of `rank` dims (innermost first), 128-byte swizzled box.
static int make_map(CUtensorMap* m, const void* ptr, int rank, const cuuint64_t* dims, const cuuint32_t* box,
                    CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return static_cast<int>(cudaErrorNotSupported);
  cuuint64_t strides[2];
  strides[0] = dims[0] * 2;
  if (rank == 3) strides[1] = dims[0] * dims[1] * 2;
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : static_cast<int>(cudaErrorInvalidValue);
}

}  // namespace tc
}  // namespace sl

extern "C" size_t sl_pop_bg_tc_ws_bytes(int B, int C, int N) {
  if (B < 1 || C < 1 || N < 1) return 0;
  return static_cast<size_t>(2) * B * N * C * sizeof(uint16_t);
}

extern "C" int sl_pop_bg_tc(const uint16_t* feat, int B, int C, int N, const uint16_t* W1p_hi, const uint16_t* W1p_lo,
                            const uint16_t* W2_hi, const uint16_t* W2_lo, const float* w3_bg, uint16_t* h1_ws,
                            float* logits, int Ktot, int ch, void* stream) {
  using namespace sl::tc;
  SL_CHECK_PTR(feat); SL_CHECK_PTR(W1p_hi); SL_CHECK_PTR(W1p_lo); SL_CHECK_PTR(W2_hi); SL_CHECK_PTR(W2_lo);
  SL_CHECK_PTR(w3_bg); SL_CHECK_PTR(h1_ws); SL_CHECK_PTR(logits);
  SL_CHECK_ARG(B >= 1 && C >= 64 && C <= 512 && C % 64 == 0 && N >= 128 && N % 128 == 0);
  SL_CHECK_ARG(Ktot >= 1 && Ktot <= SL_MAX_CLASSES && ch >= 0 && ch < Ktot);
  SL_CHECK_ARG(static_cast<long long>(B) * N / BLOCK_M < (1ll << 30));
  SL_CHECK_ALIGN(feat, 16); SL_CHECK_ALIGN(h1_ws, 16);
  SL_CHECK_ALIGN(W1p_hi, 16); SL_CHECK_ALIGN(W1p_lo, 16); SL_CHECK_ALIGN(W2_hi, 16); SL_CHECK_ALIGN(W2_lo, 16);

  Params p;
  p.C = C;
  p.NT = C <= MAX_NT ? C : C / 2;
  p.n_tiles = C / p.NT;
  p.m_tiles = static_cast<int>(static_cast<long long>(B) * N / BLOCK_M);
  p.tiles_per_image = N / BLOCK_M;
  p.N = N; p.Ktot = Ktot; p.ch = ch;
  p.h_hi = h1_ws;
  p.h_lo = h1_ws + static_cast<size_t>(B) * N * C;
  p.w3 = w3_bg;
  p.logits = logits;
  {
    const char* dbg = getenv("SL_TC_DEBUG");
    p.debug = dbg ? atoi(dbg) : 0;
  }

  CUtensorMap m_x, m_w1h, m_w1l, m_hh, m_hl, m_w2h, m_w2l, m_sh, m_sl;
  int rc;
  {  // features [B][C][N]: box = 64 pixels x 64 channels
    cuuint64_t dims[3] = {static_cast<cuuint64_t>(N), static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(B)};
    cuuint32_t box[3] = {64, BLOCK_K, 1};
    if ((rc = make_map(&m_x, feat, 3, dims, box))) return rc;
  }
  {  // weights [C_out][C_in]: box = 64 k x NT rows
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(C)};
    cuuint32_t box[2] = {BLOCK_K, static_cast<cuuint32_t>(p.NT)};
    if ((rc = make_map(&m_w1h, W1p_hi, 2, dims, box))) return rc;
    if ((rc = make_map(&m_w1l, W1p_lo, 2, dims, box))) return rc;
    if ((rc = make_map(&m_w2h, W2_hi, 2, dims, box))) return rc;
    if ((rc = make_map(&m_w2l, W2_lo, 2, dims, box))) return rc;
  }
  {  // hidden activations [B*N][C]: box = 64 k x 128 pixels
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(B) * N};
    cuuint32_t box[2] = {BLOCK_K, BLOCK_M};
    if ((rc = make_map(&m_hh, p.h_hi, 2, dims, box))) return rc;
    if ((rc = make_map(&m_hl, p.h_lo, 2, dims, box))) return rc;
    cuuint32_t sbox[2] = {32, 32};   // epilogue store: 32 channels (64 B) x 32 pixels, SWIZZLE_64B staging
    if ((rc = make_map(&m_sh, p.h_hi, 2, dims, sbox, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
    if ((rc = make_map(&m_sl, p.h_lo, 2, dims, sbox, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
  }
  cudaError_t e = cudaFuncSetAttribute(bg_gemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e != cudaSuccess) return static_cast<int>(e);
  e = cudaFuncSetAttribute(bg_gemm_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e != cudaSuccess) return static_cast<int>(e);
  const int grid = p.m_tiles < sl::kNumSMs ? p.m_tiles : sl::kNumSMs;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  bg_gemm_kernel<1><<<grid, Cfg<1>::THREADS, SMEM_BYTES, st>>>(m_x, m_x, m_w1h, m_w1l, m_sh, m_sl, p);
  e = cudaGetLastError();
  if (e != cudaSuccess) return static_cast<int>(e);
  bg_gemm_kernel<2><<<grid, Cfg<2>::THREADS, SMEM_BYTES, st>>>(m_hh, m_hl, m_w2h, m_w2l, m_sh, m_sl, p);
  return SL_LAUNCH_RESULT();
}
