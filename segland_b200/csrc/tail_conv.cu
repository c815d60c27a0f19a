// sl_tail_bn_relu_conv, fused form: PSPModule.bottleneck[1:4] -- inference BatchNorm2d ->
// ReLU -> 1x1 convolution + bias (networks/pspnet_pop.py:19-22; PSP_Plus_Decoder.fc, networks/pspplus_pop.py:44-47)
// -- as ONE tcgen05 kernel that reads the 3x3 convolution's fp32 output once and writes the head's bf16 NCHW features:
// 4 B read + 2 B written per element, no intermediate planes.
//
// The convolution is the split-bf16 product r_hi W_hi + r_lo W_hi + r_hi W_lo (fp32 accumulation in TMEM, ~1e-5 of
// fp32), r = relu(bn(x)).  tcgen05.mma takes its operands from shared memory, and TMA cannot convert, so the A
// operand is produced on the SM: eight "transform" warps load the fp32 tile of a k-block (64 channels x 128 pixels)
// straight from global memory into registers (one channel row per warp-wide 512-byte load, issued two k-blocks
// ahead), apply alpha = w / sqrt(var + eps), x * alpha + (b - mean * alpha), ReLU, split into bf16 hi / lo and store
// both planes in the MN-major SWIZZLE_128B layout a TMA box load would have produced (16-byte chunk index XOR
// channel-row & 7), then fence.proxy.async + mbarrier arrive.  Warp 0 streams the W_hi / W_lo tiles of the k-block by
// TMA; warp 1's elected thread issues the twelve MMAs of the stage (3 passes x 4 k-slices) from ONE copy of each
// operand, so a 96 KB stage feeds 1536 tensor cycles (the two-kernel path loaded 144 KB for the same work); four
// epilogue warps read the double-buffered 128 x 256 accumulators, add the bias and store bf16.
// Work item = (128-pixel tile, 256-channel n-tile); a CTA walks every n-tile of its pixel tiles, so the second
// n-tile's x comes from L2.
#include "tma.cuh"

namespace sl {
namespace tailconv {
using namespace sl::tc;

constexpr int BLOCK_M = 128, BLOCK_K = 64, UMMA_K = 16, NT = 256, STAGES = 2;
constexpr int A_PLANE = BLOCK_M * BLOCK_K * 2;            // 16 KB: one bf16 plane of the activation tile
constexpr int B_PLANE = NT * BLOCK_K * 2;                 // 32 KB: one bf16 plane of the weight tile
constexpr int STAGE_BYTES = 2 * A_PLANE + 2 * B_PLANE;    // 96 KB
constexpr int XF_WARPS = 8, EPI_WARPS = 4;
constexpr int XF_THREADS = 32 * XF_WARPS;
constexpr int THREADS = 64 + XF_THREADS + 32 * EPI_WARPS;   // 448
constexpr int ROWS_PER_WARP = BLOCK_K / XF_WARPS;          // 8 channel rows of a k-block per transform warp
constexpr int MAX_CIN = 1024;
constexpr int MAX_COUT = 2048;
constexpr int BN_BYTES = 2 * MAX_CIN * 4;
constexpr int BIAS_BYTES = MAX_COUT * 4;
constexpr int BAR_BYTES = 128;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BN_BYTES + BIAS_BYTES + BAR_BYTES;
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB per-CTA shared memory limit");
static_assert(EPI_WARPS == 4 && ((2 + XF_WARPS) & 3) == 2, "epilogue warps must cover the four TMEM lane quarters");

struct Params {
  const float* x; int B, Cin, N;
  const float* bn_w; const float* bn_b; const float* bn_m; const float* bn_v; float eps; int relu;
  const float* bias; int Cout; uint16_t* out;
  int m_tiles, m_tiles_per_img, n_tiles;
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ void st_shared_v2(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}

__global__ void __launch_bounds__(THREADS, 1)
tail_conv_kernel(const __grid_constant__ CUtensorMap map_whi, const __grid_constant__ CUtensorMap map_wlo, Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  if ((base & 1023u) != 0) asm volatile("trap;");
  float* bn_s = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);          // alpha[MAX_CIN], shift[MAX_CIN]
  float* bias_s = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + BN_BYTES);   // [roundup32(Cout)], zero padded
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + BN_BYTES + BIAS_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 4);
  const uint32_t bar0 = smem_u32(bars);
  auto b_full = [&](int s) { return bar0 + 8u * s; };
  auto a_ready = [&](int s) { return bar0 + 8u * (STAGES + s); };
  auto s_empty = [&](int s) { return bar0 + 8u * (2 * STAGES + s); };
  auto tfull_bar = [&](int s) { return bar0 + 8u * (3 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bar0 + 8u * (3 * STAGES + 2 + s); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // A CTA owns whole pixel tiles (every n-tile of tile blockIdx.x + i * gridDim.x, n-tile inner): the second n-tile's
  // rows of x are L2 hits issued by the same SM a moment after the first pass fetched them from HBM.
  const int my_mtiles = (p.m_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const int my_items = my_mtiles * p.n_tiles;
  const int kbs = (p.Cin + BLOCK_K - 1) / BLOCK_K;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_whi); tma_prefetch_desc(&map_wlo);
    for (int s = 0; s < STAGES; ++s) { mbar_init(b_full(s), 1); mbar_init(a_ready(s), XF_THREADS); mbar_init(s_empty(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int c = threadIdx.x; c < p.Cin; c += THREADS) {          // inference batch-norm as ATen evaluates it
    float alpha = 1.f, shift = 0.f;
    if (p.bn_w != nullptr) {
      alpha = __ldg(p.bn_w + c) * (1.f / sqrtf(__ldg(p.bn_v + c) + p.eps));
      shift = __ldg(p.bn_b + c) - __ldg(p.bn_m + c) * alpha;
    }
    bn_s[c] = alpha;
    bn_s[MAX_CIN + c] = shift;
  }
  for (int c = p.Cin + threadIdx.x; c < kbs * BLOCK_K; c += THREADS) { bn_s[c] = 0.f; bn_s[MAX_CIN + c] = 0.f; }
  for (int c = threadIdx.x; c < ((p.Cout + 31) & ~31); c += THREADS)
    bias_s[c] = (p.bias != nullptr && c < p.Cout) ? __ldg(p.bias + c) : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================================================== TMA producer: weight planes
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int i = 0; i < my_items; ++i) {
        const int n0 = (i % p.n_tiles) * NT;
        for (int kb = 0; kb < kbs; ++kb) {
          mbar_wait(s_empty(stage), phase ^ 1u);
          mbar_expect_tx(b_full(stage), 2 * B_PLANE);
          const uint32_t sb = base + stage * STAGE_BYTES + 2 * A_PLANE;
          tma_load_2d(sb, &map_whi, b_full(stage), kb * BLOCK_K, n0, L2_EVICT_LAST);
          tma_load_2d(sb + B_PLANE, &map_wlo, b_full(stage), kb * BLOCK_K, n0, L2_EVICT_LAST);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc(BLOCK_M, NT, true, 1u, 1u, false);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int i = 0; i < my_items; ++i) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * NT);
        uint32_t accumulate = 0;
        for (int kb = 0; kb < kbs; ++kb) {
          mbar_wait(b_full(stage), phase);
          mbar_wait(a_ready(stage), phase);
          tc_fence_after();
          const uint32_t sa_hi = base + stage * STAGE_BYTES, sa_lo = sa_hi + A_PLANE;
          const uint32_t sb_hi = sa_hi + 2 * A_PLANE, sb_lo = sb_hi + B_PLANE;
#pragma unroll
          for (int pass = 0; pass < 3; ++pass) {                 // r_hi W_hi + r_lo W_hi + r_hi W_lo
            const uint32_t sa = pass == 1 ? sa_lo : sa_hi, sb = pass == 2 ? sb_lo : sb_hi;
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              tc_mma(d_tmem, make_desc(sa + k * (UMMA_K * 128), 8192, 1024), make_desc(sb + k * (UMMA_K * 2), 16, 1024),
                     idesc, accumulate);
              accumulate = 1;
            }
          }
          tc_commit(s_empty(stage));                             // both operands of the stage are free once these retire
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        tc_commit(tfull_bar(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else if (warp < 2 + XF_WARPS) {
    // ===================================================================== transform warps: fp32 -> BN/ReLU -> bf16 hi/lo
    const int xw = warp - 2;
    // lane = pixels [4*lane, 4*lane+4) of the 128-pixel tile: box = lane/16 (64-pixel SWIZZLE_128B box), 16-byte chunk
    // (lane%16)/2, 8-byte half lane%2; channel row r of the k-block lives at r*128 with its chunks XORed by r&7
    const uint32_t lane_off = static_cast<uint32_t>((lane >> 4) * 8192 + (lane & 1) * 8);
    const uint32_t chunk = static_cast<uint32_t>((lane & 15) >> 1);
    // Two k-blocks of fp32 rows are kept in flight per thread (2 x 8 x 16 B).  The loop is written for instruction
    // count: the first version spent ~700 instructions per warp and k-block (integer divisions for the work-item
    // decoding, per-row predicates) and the eight transform warps were issue-bound at ~2800 cycles per k-block against
    // 1536 of tensor work (ncu source page).  Now: running cursors (a division only when the work item changes),
    // unpredicated loads from clamped addresses (rows past Cin read row Cin-1 and are zeroed by alpha = shift = 0 in
    // the table, pixels past N read the tile's last valid pixels and land in accumulator rows that are never stored).
    float4 preA[ROWS_PER_WARP], preB[ROWS_PER_WARP];
    const int total = my_items * kbs;
    const size_t row_stride = static_cast<size_t>(p.N);
    // load cursor
    int ld_item = 0, ld_kb = 0;
    const float* ld_base = nullptr;                              // x[img][0][clamped pixel of this lane]
    auto ld_item_setup = [&]() {
      const int mt = blockIdx.x + (ld_item / p.n_tiles) * gridDim.x;
      const int img = mt / p.m_tiles_per_img;
      const int px = min((mt - img * p.m_tiles_per_img) * BLOCK_M + 4 * lane, p.N - 4);
      ld_base = p.x + static_cast<size_t>(img) * p.Cin * row_stride + px;
    };
    auto issue_loads = [&](float4 (&pre)[ROWS_PER_WARP]) {
      const int c0 = ld_kb * BLOCK_K + xw * ROWS_PER_WARP;
      if (c0 + ROWS_PER_WARP <= p.Cin) {
        const float* src = ld_base + static_cast<size_t>(c0) * row_stride;
#pragma unroll
        for (int j = 0; j < ROWS_PER_WARP; ++j) pre[j] = __ldg(reinterpret_cast<const float4*>(src + j * row_stride));
      } else {
#pragma unroll
        for (int j = 0; j < ROWS_PER_WARP; ++j)
          pre[j] = __ldg(reinterpret_cast<const float4*>(ld_base + static_cast<size_t>(min(c0 + j, p.Cin - 1)) * row_stride));
      }
      if (++ld_kb == kbs) { ld_kb = 0; ++ld_item; if (ld_item < my_items) ld_item_setup(); }
    };
    int stage = 0; uint32_t phase = 0;
    int kb_proc = 0;
    const bool relu = p.relu != 0;
    auto do_step = [&](float4 (&pre)[ROWS_PER_WARP], bool more) {
      const int c0 = kb_proc * BLOCK_K + xw * ROWS_PER_WARP;
      if (++kb_proc == kbs) kb_proc = 0;
      float al[ROWS_PER_WARP], sh[ROWS_PER_WARP];
      {
        const float4 a0 = *reinterpret_cast<const float4*>(bn_s + c0), a1 = *reinterpret_cast<const float4*>(bn_s + c0 + 4);
        const float4 s0 = *reinterpret_cast<const float4*>(bn_s + MAX_CIN + c0);
        const float4 s1 = *reinterpret_cast<const float4*>(bn_s + MAX_CIN + c0 + 4);
        al[0] = a0.x; al[1] = a0.y; al[2] = a0.z; al[3] = a0.w; al[4] = a1.x; al[5] = a1.y; al[6] = a1.z; al[7] = a1.w;
        sh[0] = s0.x; sh[1] = s0.y; sh[2] = s0.z; sh[3] = s0.w; sh[4] = s1.x; sh[5] = s1.y; sh[6] = s1.z; sh[7] = s1.w;
      }
      uint32_t hi[ROWS_PER_WARP][2], lo[ROWS_PER_WARP][2];
#pragma unroll
      for (int j = 0; j < ROWS_PER_WARP; ++j) {
        float v[4] = {fmaf(pre[j].x, al[j], sh[j]), fmaf(pre[j].y, al[j], sh[j]), fmaf(pre[j].z, al[j], sh[j]),
                      fmaf(pre[j].w, al[j], sh[j])};
        if (relu) {
#pragma unroll
          for (int e = 0; e < 4; ++e) v[e] = fmaxf(v[e], 0.f);
        }
        hi[j][0] = pack_bf16(v[0], v[1]);
        hi[j][1] = pack_bf16(v[2], v[3]);
        lo[j][0] = pack_bf16(v[0] - bf16lo(hi[j][0]), v[1] - bf16hi(hi[j][0]));
        lo[j][1] = pack_bf16(v[2] - bf16lo(hi[j][1]), v[3] - bf16hi(hi[j][1]));
      }
      mbar_wait(s_empty(stage), phase ^ 1u);
      const uint32_t sa = base + stage * STAGE_BYTES + lane_off;
#pragma unroll
      for (int j = 0; j < ROWS_PER_WARP; ++j) {
        const uint32_t r = static_cast<uint32_t>(xw * ROWS_PER_WARP + j);
        const uint32_t off = r * 128u + ((chunk ^ (r & 7u)) << 4);
        st_shared_v2(sa + off, hi[j][0], hi[j][1]);
        st_shared_v2(sa + A_PLANE + off, lo[j][0], lo[j][1]);
      }
      fence_async_smem();                                        // generic-proxy stores -> visible to tcgen05.mma
      mbar_arrive(a_ready(stage));
      // the proxy fence / release above wait for the thread's outstanding loads, so the refill goes after them
      if (more) issue_loads(pre);
      if (++stage == STAGES) { stage = 0; phase ^= 1u; }
    };
    if (my_items > 0) ld_item_setup();
    if (total > 0) issue_loads(preA);
    if (total > 1) issue_loads(preB);
    for (int step = 0; step < total; step += 2) {
      do_step(preA, step + 2 < total);
      if (step + 1 < total) do_step(preB, step + 3 < total);
    }
  } else {
    // ===================================================================== epilogue: + bias, bf16, NCHW
    const int sub = warp & 3;
    const int row = sub * 32 + lane;
    int acc = 0; uint32_t acc_phase = 0;
    for (int i = 0; i < my_items; ++i) {
      const int nt = i % p.n_tiles, mt = blockIdx.x + (i / p.n_tiles) * gridDim.x;
      const int img = mt / p.m_tiles_per_img;
      const int n_in_img = (mt - img * p.m_tiles_per_img) * BLOCK_M + row;
      const bool row_ok = n_in_img < p.N;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(sub * 32) << 16) + static_cast<uint32_t>(acc * NT);
      for (int ch = 0; ch < NT / 32; ++ch) {
        const int col0 = nt * NT + ch * 32;
        if (col0 >= p.Cout) break;
        const int ncols = min(32, p.Cout - col0);
        uint32_t r[32];
        tc_ld32(taddr + ch * 32, r);
        tc_ld_wait();
        // bias from shared memory (uniform 128-bit reads), all 32 conversions first, then the stores back to back:
        // the first version (a __ldg of the bias, a branch and a dependent add per element) kept the four epilogue
        // warps busy for 49 k cycles per item, four times the MMA time (ncu source page, profiles/r1_ncu_tailconv.txt)
        uint32_t packed[16];
        const float4* bv = reinterpret_cast<const float4*>(bias_s + col0);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 b4 = bv[q];
          packed[2 * q] = pack_bf16(__uint_as_float(r[4 * q]) + b4.x, __uint_as_float(r[4 * q + 1]) + b4.y);
          packed[2 * q + 1] = pack_bf16(__uint_as_float(r[4 * q + 2]) + b4.z, __uint_as_float(r[4 * q + 3]) + b4.w);
        }
        if (row_ok) {
          uint16_t* o = p.out + (static_cast<size_t>(img) * p.Cout + col0) * p.N + n_in_img;
          const size_t nn = static_cast<size_t>(p.N);
          if (ncols == 32) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              o[(2 * j) * nn] = static_cast<uint16_t>(packed[j] & 0xffffu);
              o[(2 * j + 1) * nn] = static_cast<uint16_t>(packed[j] >> 16);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < ncols) o[j * nn] = static_cast<uint16_t>((j & 1) ? (packed[j >> 1] >> 16) : (packed[j >> 1] & 0xffffu));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

}  // namespace tailconv
}  // namespace sl

// 0 = launched; SL_EINVAL = shape outside this kernel's range (the caller falls back to the two-kernel path)
int sl_tail_conv_fused_run(const float* x, int B, int Cin, int N, const float* bn_w, const float* bn_b, const float* bn_m,
                           const float* bn_v, float eps, int relu, const uint16_t* W_hi, const uint16_t* W_lo,
                           const float* bias, int Cout, uint16_t* out, cudaStream_t st) {
  using namespace sl::tailconv;
  if (Cin > MAX_CIN || Cout > MAX_COUT || Cin % 8 != 0 || N % 8 != 0) return SL_EINVAL;
  Params p{};
  p.x = x; p.B = B; p.Cin = Cin; p.N = N;
  p.bn_w = bn_w; p.bn_b = bn_b; p.bn_m = bn_m; p.bn_v = bn_v; p.eps = eps; p.relu = relu;
  p.bias = bias; p.Cout = Cout; p.out = out;
  p.m_tiles_per_img = (N + BLOCK_M - 1) / BLOCK_M;
  p.m_tiles = B * p.m_tiles_per_img;
  p.n_tiles = (Cout + NT - 1) / NT;
  CUtensorMap mh, ml;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(Cin), static_cast<cuuint64_t>(Cout)};
  cuuint32_t box[2] = {BLOCK_K, NT};
  int rc;
  if ((rc = make_map(&mh, W_hi, 2, dims, box)) != 0) return rc;
  if ((rc = make_map(&ml, W_lo, 2, dims, box)) != 0) return rc;
  cudaError_t e = cudaFuncSetAttribute(tail_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e != cudaSuccess) return static_cast<int>(e);
  if (static_cast<long long>(p.m_tiles) * p.n_tiles >= (1ll << 31)) return SL_EINVAL;
  const int grid = p.m_tiles < sl::kNumSMs ? p.m_tiles : sl::kNumSMs;
  tail_conv_kernel<<<grid, THREADS, SMEM_BYTES, st>>>(mh, ml, p);
  return SL_LAUNCH_RESULT();
}
