// sl_tail_bn_relu_conv, fused form: PSPModule.bottleneck[1:4] -- inference BatchNorm2d ->
// ReLU -> 1x1 convolution + bias (networks/pspnet_pop.py:19-22; PSP_Plus_Decoder.fc, networks/pspplus_pop.py:44-47)
// -- as ONE tcgen05 kernel that reads the 3x3 convolution's fp32 output once and writes the head's bf16 NCHW features:
// 4 B read + 2 B written per element, no intermediate planes.
//
// The convolution is the split-bf16 product r_hi W_hi + r_lo W_hi + r_hi W_lo (fp32 accumulation in TMEM, ~1e-5 of
// fp32), r = relu(bn(x)).  tcgen05.mma takes its operands from shared memory, and TMA cannot convert, so the A
// operand is produced on the SM: eight "transform" warps load the fp32 tile of a k-block (64 channels x 128 pixels)
// straight from global memory into registers (one channel row per warp-wide 512-byte load, four half k-blocks in
// flight), apply alpha = w / sqrt(var + eps), x * alpha + (b - mean * alpha), ReLU, split into bf16 hi / lo and store
// both planes in the MN-major SWIZZLE_128B layout a TMA box load would have produced (16-byte chunk index XOR
// channel-row & 7), then fence.proxy.async + mbarrier arrive.  Warp 0 streams the W_hi / W_lo tiles of the k-block by
// TMA; warp 1's elected thread issues the twelve MMAs of the stage (3 passes x 4 k-slices) from ONE copy of each
// operand, so 96 KB of operands feed 1536 tensor cycles (the two-kernel path loaded 144 KB for the same work); four
// epilogue warps read the double-buffered 128 x 256 accumulators, add the bias and store bf16.
// Work item = (128-pixel tile, 256-channel n-tile); a CTA walks every n-tile of its pixel tiles, so the second
// n-tile's x comes from L2.
#include "tma.cuh"

namespace sl {
namespace tailconv {
using namespace sl::tc;

constexpr int BLOCK_M = 128, BLOCK_K = 64, UMMA_K = 16, NT = 256;
constexpr int B_STAGES = 2;                               // weight ring: k-blocks of 64 channels
constexpr int B_PLANE = NT * BLOCK_K * 2;                 // 32 KB: one bf16 plane of the weight tile
constexpr int B_STAGE_BYTES = 2 * B_PLANE;                // W_hi + W_lo
constexpr int HALF_K = 32;                                // activation ring: half k-blocks of 32 channels
constexpr int A_SLOTS = 4;
constexpr int A_BOX = HALF_K * 128;                       // 4 KB: 32 channel rows x 64 pixels (one SWIZZLE_128B box)
constexpr int A_PLANE = 2 * A_BOX;                        // 8 KB: one bf16 plane of a 128-pixel half k-block
constexpr int A_SLOT_BYTES = 2 * A_PLANE;                 // hi + lo
constexpr int OPERAND_BYTES = B_STAGES * B_STAGE_BYTES + A_SLOTS * A_SLOT_BYTES;   // 128 + 64 = 192 KB
constexpr int XF_WARPS = 8, EPI_WARPS = 4;
constexpr int XF_THREADS = 32 * XF_WARPS;
constexpr int THREADS = 64 + XF_THREADS + 32 * EPI_WARPS;   // 448
constexpr int ROWS_PER_WARP = HALF_K / XF_WARPS;           // 4 channel rows of a half k-block per transform warp
constexpr int N_BUF = 4;                                   // half k-blocks of fp32 rows in flight per thread
constexpr int MAX_CIN = 1024;
constexpr int MAX_COUT = 2048;
constexpr int BN_BYTES = 2 * MAX_CIN * 4;
constexpr int BIAS_BYTES = MAX_COUT * 4;
constexpr int BAR_BYTES = 256;
constexpr int SMEM_BYTES = OPERAND_BYTES + BN_BYTES + BIAS_BYTES + BAR_BYTES;
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
  float* bn_s = reinterpret_cast<float*>(smem + OPERAND_BYTES);                 // alpha[MAX_CIN], shift[MAX_CIN]
  float* bias_s = reinterpret_cast<float*>(smem + OPERAND_BYTES + BN_BYTES);          // [roundup32(Cout)], zero padded
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OPERAND_BYTES + BN_BYTES + BIAS_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * B_STAGES + 2 * A_SLOTS + 4);
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t a_base = base + B_STAGES * B_STAGE_BYTES;
  auto b_full = [&](int s) { return bar0 + 8u * s; };
  auto b_empty = [&](int s) { return bar0 + 8u * (B_STAGES + s); };
  auto a_ready = [&](int q) { return bar0 + 8u * (2 * B_STAGES + q); };
  auto a_empty = [&](int q) { return bar0 + 8u * (2 * B_STAGES + A_SLOTS + q); };
  auto tfull_bar = [&](int s) { return bar0 + 8u * (2 * B_STAGES + 2 * A_SLOTS + s); };
  auto tempty_bar = [&](int s) { return bar0 + 8u * (2 * B_STAGES + 2 * A_SLOTS + 2 + s); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // A CTA owns whole pixel tiles (every n-tile of tile blockIdx.x + i * gridDim.x, n-tile inner): the second n-tile's
  // rows of x are L2 hits issued by the same SM a moment after the first pass fetched them from HBM.
  const int my_mtiles = (p.m_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const int my_items = my_mtiles * p.n_tiles;
  const int kbs = (p.Cin + BLOCK_K - 1) / BLOCK_K;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_whi); tma_prefetch_desc(&map_wlo);
    for (int s = 0; s < B_STAGES; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
    for (int q = 0; q < A_SLOTS; ++q) { mbar_init(a_ready(q), XF_THREADS); mbar_init(a_empty(q), 1); }
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
          mbar_wait(b_empty(stage), phase ^ 1u);
          mbar_expect_tx(b_full(stage), 2 * B_PLANE);
          const uint32_t sb = base + stage * B_STAGE_BYTES;
          tma_load_2d(sb, &map_whi, b_full(stage), kb * BLOCK_K, n0, L2_EVICT_LAST);
          tma_load_2d(sb + B_PLANE, &map_wlo, b_full(stage), kb * BLOCK_K, n0, L2_EVICT_LAST);
          if (++stage == B_STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc(BLOCK_M, NT, true, 1u, 1u, false);
      int stage = 0; uint32_t phase = 0;
      int aq = 0; uint32_t a_phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int i = 0; i < my_items; ++i) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * NT);
        uint32_t accumulate = 0;
        for (int kb = 0; kb < kbs; ++kb) {
          mbar_wait(b_full(stage), phase);
          const uint32_t sb_hi = base + stage * B_STAGE_BYTES, sb_lo = sb_hi + B_PLANE;
#pragma unroll
          for (int h = 0; h < BLOCK_K / HALF_K; ++h) {           // the activation tile arrives in half k-blocks
            mbar_wait(a_ready(aq), a_phase);
            tc_fence_after();
            const uint32_t sa_hi = a_base + aq * A_SLOT_BYTES, sa_lo = sa_hi + A_PLANE;
#pragma unroll
            for (int pass = 0; pass < 3; ++pass) {               // r_hi W_hi + r_lo W_hi + r_hi W_lo
              const uint32_t sa = pass == 1 ? sa_lo : sa_hi, sb = pass == 2 ? sb_lo : sb_hi;
#pragma unroll
              for (int k = 0; k < HALF_K / UMMA_K; ++k) {
                tc_mma(d_tmem, make_desc(sa + k * (UMMA_K * 128), A_BOX, 1024),
                       make_desc(sb + (h * (HALF_K / UMMA_K) + k) * (UMMA_K * 2), 16, 1024), idesc, accumulate);
                accumulate = 1;
              }
            }
            tc_commit(a_empty(aq));                              // the half-slot is free once these six MMAs retire
            if (++aq == A_SLOTS) { aq = 0; a_phase ^= 1u; }
          }
          tc_commit(b_empty(stage));
          if (++stage == B_STAGES) { stage = 0; phase ^= 1u; }
        }
        tc_commit(tfull_bar(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else if (warp < 2 + XF_WARPS) {
    // ===================================================================== transform warps: fp32 -> BN/ReLU -> bf16 hi/lo
    const int xw = warp - 2;
    // lane = pixels [4*lane, 4*lane+4) of the 128-pixel tile: box = lane/16 (64-pixel SWIZZLE_128B box), 16-byte chunk
    // (lane%16)/2, 8-byte half lane%2; channel row r of the half k-block lives at r*128 with its chunks XORed by r&7
    const uint32_t lane_off = static_cast<uint32_t>((lane >> 4) * A_BOX + (lane & 1) * 8);
    const uint32_t chunk = static_cast<uint32_t>((lane & 15) >> 1);
    // The unit of work is a HALF k-block (32 channels, 4 rows per warp = 16 registers per thread), and four of them
    // are in flight per thread: with whole k-blocks only two fit under the 128-register ceiling, i.e. one HBM round
    // trip per ~1,500 tensor cycles, and the first use of a loaded row was the transform warps' top stall (tensor pipe
    // 73 % active).  The loop is written for instruction count: running cursors (a division only when the work item
    // changes), unpredicated loads from clamped addresses (rows past Cin read row Cin-1 and are zeroed by
    // alpha = shift = 0 in the table, pixels past N read the tile's last valid pixels and land in accumulator rows
    // that are never stored).
    float4 pre0[ROWS_PER_WARP], pre1[ROWS_PER_WARP], pre2[ROWS_PER_WARP], pre3[ROWS_PER_WARP];
    const int total = my_items * kbs * (BLOCK_K / HALF_K);       // half-steps of this CTA
    const size_t row_stride = static_cast<size_t>(p.N);
    // load cursor
    int ld_item = 0, ld_c = 0;                                   // ld_c: first channel of the next half k-block
    const int c_end = kbs * BLOCK_K;
    const float* ld_base = nullptr;                              // x[img][0][clamped pixel of this lane]
    auto ld_item_setup = [&]() {
      const int mt = blockIdx.x + (ld_item / p.n_tiles) * gridDim.x;
      const int img = mt / p.m_tiles_per_img;
      const int px = min((mt - img * p.m_tiles_per_img) * BLOCK_M + 4 * lane, p.N - 4);
      ld_base = p.x + static_cast<size_t>(img) * p.Cin * row_stride + px;
    };
    auto issue_loads = [&](float4 (&pre)[ROWS_PER_WARP]) {
      const int c0 = ld_c + xw * ROWS_PER_WARP;
      if (c0 + ROWS_PER_WARP <= p.Cin) {
        const float* src = ld_base + static_cast<size_t>(c0) * row_stride;
#pragma unroll
        for (int j = 0; j < ROWS_PER_WARP; ++j) pre[j] = __ldg(reinterpret_cast<const float4*>(src + j * row_stride));
      } else {
#pragma unroll
        for (int j = 0; j < ROWS_PER_WARP; ++j)
          pre[j] = __ldg(reinterpret_cast<const float4*>(ld_base + static_cast<size_t>(min(c0 + j, p.Cin - 1)) * row_stride));
      }
      ld_c += HALF_K;
      if (ld_c == c_end) { ld_c = 0; ++ld_item; if (ld_item < my_items) ld_item_setup(); }
    };
    int aq = 0; uint32_t a_phase = 0;
    int c_proc = 0;
    const bool relu = p.relu != 0;
    auto do_step = [&](float4 (&pre)[ROWS_PER_WARP], bool more) {
      const int c0 = c_proc + xw * ROWS_PER_WARP;
      c_proc += HALF_K;
      if (c_proc == c_end) c_proc = 0;
      const float4 al4 = *reinterpret_cast<const float4*>(bn_s + c0), sh4 = *reinterpret_cast<const float4*>(bn_s + MAX_CIN + c0);
      const float al[ROWS_PER_WARP] = {al4.x, al4.y, al4.z, al4.w}, sh[ROWS_PER_WARP] = {sh4.x, sh4.y, sh4.z, sh4.w};
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
      mbar_wait(a_empty(aq), a_phase ^ 1u);
      const uint32_t sa = a_base + aq * A_SLOT_BYTES + lane_off;
#pragma unroll
      for (int j = 0; j < ROWS_PER_WARP; ++j) {
        const uint32_t r = static_cast<uint32_t>(xw * ROWS_PER_WARP + j);
        const uint32_t off = r * 128u + ((chunk ^ (r & 7u)) << 4);
        st_shared_v2(sa + off, hi[j][0], hi[j][1]);
        st_shared_v2(sa + A_PLANE + off, lo[j][0], lo[j][1]);
      }
      fence_async_smem();                                        // generic-proxy stores -> visible to tcgen05.mma
      mbar_arrive(a_ready(aq));
      if (more) issue_loads(pre);                                // refill after the fence (it waits for outstanding loads)
      if (++aq == A_SLOTS) { aq = 0; a_phase ^= 1u; }
    };
    if (my_items > 0) ld_item_setup();
    if (total > 0) issue_loads(pre0);
    if (total > 1) issue_loads(pre1);
    if (total > 2) issue_loads(pre2);
    if (total > 3) issue_loads(pre3);
    for (int step = 0; step < total; step += N_BUF) {            // total is even; the tail is guarded
      do_step(pre0, step + 4 < total);
      if (step + 1 < total) do_step(pre1, step + 5 < total);
      if (step + 2 < total) do_step(pre2, step + 6 < total);
      if (step + 3 < total) do_step(pre3, step + 7 < total);
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
  const int grid = p.m_tiles < sl::num_sms() ? p.m_tiles : sl::num_sms();
  tail_conv_kernel<<<grid, THREADS, SMEM_BYTES, st>>>(mh, ml, p);
  return SL_LAUNCH_RESULT();
}
