// Tensor-core (tcgen05) implementation of the POP-head backward, sl_pop_head_bwd with SL_BWD_TC.
//   reference: autograd through networks/pspnet_pop.py:95-121 (orthogonal_decompose) and :46-63 (classifier /
//   classifier_n) as called by forward_novel (:199-219) and forward_base (:169-182).
// The five (six with d_feat) GEMMs of pop_bwd.cu run as split-bf16 products with fp32 accumulation in TMEM,
// the same numerics as the forward's "precise" mode (~1e-5 of fp32):
//   G1  h1  = relu(q W1'^T)                      q exact bf16        -> 2 passes   A = features (MN-major, 3-D map)
//   G2  z2  = h1 W2^T;  dz2 = g0 w3 [z2>0]       h1 = hi + lo        -> 3 passes   + dw3 = sum_px g0 relu(z2)
//   G3  dz1 = (dz2 W2) [h1>0]                    dz2 = hi + lo       -> 3 passes   B = W2^T (K-major)
//   G4  dW2 = dz2^T h1          (K = pixels)     both MN-major       -> 3 passes   split-K, fp32 atomics
//   G5  dW1'= dz1^T q           (K = pixels)     A MN-major, B = features (K-major, 3-D map) -> 2 passes
//   G6  d_q = dz1 W1' + sum_k gp_k s_hat_k       optional            -> 3 passes   transposed fp32 store
// One generic persistent warp-specialised kernel (warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2-9 = epilogue) walks work items (m-tile, n-tile, k-chunk); operands arrive through a 4-stage
// mbarrier ring of 16 KB (A) + up to 32 KB (B) SWIZZLE_128B tiles; 128 x NT accumulators are double-buffered in
// TMEM.  Hidden activations travel between the GEMMs as bf16 hi/lo pairs [pixel][C] in the workspace.
#include "tma.cuh"

namespace sl {
namespace tcg {
using namespace sl::tc;

constexpr int BLOCK_M = 128, BLOCK_K = 64, UMMA_K = 16, STAGES = 4, MAX_NT = 256;
constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
constexpr int STAGE_BYTES = A_BYTES + MAX_NT * BLOCK_K * 2;
constexpr int EPI_WARPS = 8;
constexpr int THREADS = 64 + 32 * EPI_WARPS;
constexpr int DW3_BYTES = EPI_WARPS * 512 * 4;
constexpr int BAR_BYTES = 128;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + DW3_BYTES + BAR_BYTES;
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB per-CTA shared memory limit");
constexpr uint64_t L2_EVICT_NORMAL = 0x1000000000000000ull;

enum LoadKind { LD_K2D = 0, LD_MN2D = 1, LD_MN3D = 2, LD_K3D = 3 };
enum EpiKind { EPI_RELU_SPLIT = 0, EPI_LAYER2 = 1, EPI_MASK_SPLIT = 2, EPI_RED = 3, EPI_DFEAT = 4, EPI_TAIL = 5 };

struct GemmMaps { CUtensorMap a[3], b[3]; };

struct GemmParams {
  int passes, a_kind, b_kind;
  int m_is_px, k_is_px, k_flat;
  int C, N_img, B;
  int NT, n_tiles, m_tiles, m_tiles_per_img;
  int k_chunks, chunks_per_img, chunk_kb;
  int n_valid, m_valid;
  unsigned long long a_policy, b_policy;
  uint16_t* out_hi; uint16_t* out_lo; uint16_t* out_lo2;   // lo2: third bf16 term (G1 only; 24 mantissa bits in all)
  int* flag_count; int* flag_list; int flag_cap;            // G2: elements whose sign(z2) is within rounding noise
  const uint16_t* mask_hi;
  const float* w3; const float* g; int Ktot, ch;
  float* dw3;
  float* red_out;
  float* d_feat; const float* gp; const float* s_hat; int K;
  uint16_t* tail_out; const float* tail_bias; int C_out;    // EPI_TAIL: bf16 NCHW features = acc + bias (decoder tail, f-4)
};

__device__ __forceinline__ void split_bf16(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
  const uint32_t hb = *reinterpret_cast<const uint32_t*>(&h2);
  const __nv_bfloat162 l2 = __floats2bfloat162_rn(a - __uint_as_float(hb << 16), b - __uint_as_float(hb & 0xffff0000u));
  hi = hb;
  lo = *reinterpret_cast<const uint32_t*>(&l2);
}

template <int EPI>
__global__ void __launch_bounds__(THREADS, 1) tc_gemm_kernel(const __grid_constant__ GemmMaps maps, GemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  if ((base & 1023u) != 0) asm volatile("trap;");
  float* dw3_s = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + DW3_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bar0 + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bar0 + 8u * (2 * STAGES + 2 + s); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int items = p.m_tiles * p.n_tiles * p.k_chunks;
  const uint32_t b_bytes = static_cast<uint32_t>(p.NT) * BLOCK_K * 2;

  struct Item { int nt, mt, kc, kbs, k0, b_k; };
  auto decode = [&](int w) {
    Item it;
    it.nt = w % p.n_tiles;
    const int r = w / p.n_tiles;
    it.mt = r % p.m_tiles;
    it.kc = r / p.m_tiles;
    if (p.k_is_px) {
      const int extent = p.k_flat ? p.B * p.N_img : p.N_img;
      const int c = p.k_flat ? it.kc : it.kc % p.chunks_per_img;
      it.b_k = p.k_flat ? 0 : it.kc / p.chunks_per_img;
      it.k0 = c * p.chunk_kb * BLOCK_K;
      const int k1 = min(extent, it.k0 + p.chunk_kb * BLOCK_K);
      it.kbs = (k1 - it.k0 + BLOCK_K - 1) / BLOCK_K;
    } else {
      it.b_k = 0; it.k0 = 0;
      it.kbs = (p.C + BLOCK_K - 1) / BLOCK_K;
    }
    return it;
  };

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < p.passes; ++i) { tma_prefetch_desc(&maps.a[i]); tma_prefetch_desc(&maps.b[i]); }
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (EPI == EPI_LAYER2)
    for (int i = threadIdx.x; i < EPI_WARPS * 512; i += blockDim.x) dw3_s[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int w = blockIdx.x; w < items; w += gridDim.x) {
        const Item it = decode(w);
        const int n0 = it.nt * p.NT;
        int m_row = 0, m_img = 0, m_n0 = 0;
        if (p.m_is_px) {
          m_img = it.mt / p.m_tiles_per_img;
          m_n0 = (it.mt - m_img * p.m_tiles_per_img) * BLOCK_M;
          m_row = m_img * p.N_img + m_n0;
        } else {
          m_row = it.mt * BLOCK_M;
        }
        for (int kb = 0; kb < it.kbs; ++kb) {
          const int kk = it.k0 + kb * BLOCK_K;                       // K coordinate (within the image when !k_flat)
          const int kk_row = it.b_k * p.N_img + kk;                  // row of a [pixel][C] operand
          for (int pass = 0; pass < p.passes; ++pass) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            mbar_expect_tx(full_bar(stage), A_BYTES + b_bytes);
            const uint32_t sa = base + stage * STAGE_BYTES, sb = sa + A_BYTES;
            const CUtensorMap* ma = &maps.a[pass];
            const CUtensorMap* mb = &maps.b[pass];
            if (p.a_kind == LD_K2D) {
              tma_load_2d(sa, ma, full_bar(stage), kk, m_row, p.a_policy);
            } else if (p.a_kind == LD_MN3D) {
              tma_load_3d(sa, ma, full_bar(stage), m_n0, kk, m_img, p.a_policy);
              tma_load_3d(sa + A_BYTES / 2, ma, full_bar(stage), m_n0 + 64, kk, m_img, p.a_policy);
            } else {  // LD_MN2D: [pixel][C] with M = channel
              tma_load_2d(sa, ma, full_bar(stage), m_row, kk_row, p.a_policy);
              tma_load_2d(sa + A_BYTES / 2, ma, full_bar(stage), m_row + 64, kk_row, p.a_policy);
            }
            if (p.b_kind == LD_K2D) {
              tma_load_2d(sb, mb, full_bar(stage), kk, n0, p.b_policy);
            } else if (p.b_kind == LD_K3D) {
              tma_load_3d(sb, mb, full_bar(stage), kk, n0, it.b_k, p.b_policy);
            } else {  // LD_MN2D
              for (int h = 0; h < p.NT / 64; ++h)
                tma_load_2d(sb + h * 8192, mb, full_bar(stage), n0 + 64 * h, kk_row, p.b_policy);
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      const bool a_mn = p.a_kind != LD_K2D, b_mn = p.b_kind == LD_MN2D;
      const uint32_t idesc = make_idesc(BLOCK_M, p.NT, a_mn, 1u, 1u, b_mn);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int w = blockIdx.x; w < items; w += gridDim.x) {
        const Item it = decode(w);
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * MAX_NT);
        uint32_t accumulate = 0;
        const int iters = it.kbs * p.passes;
        for (int i = 0; i < iters; ++i) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = base + stage * STAGE_BYTES, sb = sa + A_BYTES;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t da = a_mn ? make_desc(sa + k * (UMMA_K * 128), 8192, 1024) : make_desc(sa + k * (UMMA_K * 2), 16, 1024);
            const uint64_t db = b_mn ? make_desc(sb + k * (UMMA_K * 128), 8192, 1024) : make_desc(sb + k * (UMMA_K * 2), 16, 1024);
            tc_mma(d_tmem, da, db, idesc, accumulate);
            accumulate = 1;
          }
          tc_commit(empty_bar(stage));
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        tc_commit(tfull_bar(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // ===================================================================== epilogue (8 warps)
    const int sub = warp & 3;
    const int row = sub * 32 + lane;
    const int half = (warp - 2) >> 2;
    const int ew = warp - 2;
    const int n_chunks = p.NT / 32;
    const int c_split = (n_chunks + 1) / 2;
    int acc = 0; uint32_t acc_phase = 0;
    for (int w = blockIdx.x; w < items; w += gridDim.x) {
      const Item it = decode(w);
      // row bookkeeping
      bool row_ok; long long out_row; int img = 0, n_in_img = 0;
      if (p.m_is_px) {
        img = it.mt / p.m_tiles_per_img;
        n_in_img = (it.mt - img * p.m_tiles_per_img) * BLOCK_M + row;
        row_ok = n_in_img < p.N_img;
        out_row = static_cast<long long>(img) * p.N_img + n_in_img;
      } else {
        out_row = it.mt * BLOCK_M + row;
        row_ok = out_row < p.m_valid;
      }
      float g0 = 0.f;
      if (EPI == EPI_LAYER2 && row_ok) g0 = p.g[(static_cast<long long>(img) * p.Ktot + p.ch) * p.N_img + n_in_img];

      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(sub * 32) << 16) + static_cast<uint32_t>(acc * MAX_NT);
      const int c_begin = half == 0 ? 0 : c_split, c_end = half == 0 ? c_split : n_chunks;
      for (int ch = c_begin; ch < c_end; ++ch) {
        const int col0 = it.nt * p.NT + ch * 32;
        if (col0 >= p.n_valid) break;
        const int ncols = min(32, p.n_valid - col0);          // multiple of 8
        uint32_t r[32];
        tc_ld32(taddr + ch * 32, r);
        tc_ld_wait();
        if (EPI == EPI_RELU_SPLIT || EPI == EPI_LAYER2 || EPI == EPI_MASK_SPLIT) {
          float v[32];
          if (EPI == EPI_RELU_SPLIT) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(__uint_as_float(r[j]), 0.f);
          } else if (EPI == EPI_LAYER2) {
            float t[32];
            // sign(z2) decides the ReLU mask; the split-bf16 product carries ~1e-5 relative noise, so elements
            // within 2.5e-4 of the local scale are queued for the exact fp32 recomputation (mask_fixup_kernel)
            float ss = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) ss = fmaf(__uint_as_float(r[j]), __uint_as_float(r[j]), ss);
            const float tau = 2.5e-4f * sqrtf(ss * (1.f / 32.f));
            if (row_ok) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < ncols && fabsf(__uint_as_float(r[j])) < tau) {
                  const int slot = atomicAdd(p.flag_count, 1);
                  if (slot < p.flag_cap) p.flag_list[slot] = static_cast<int>(out_row) * p.C + col0 + j;
                }
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float z = __uint_as_float(r[j]);
              const float w3v = (j < ncols) ? __ldg(p.w3 + col0 + j) : 0.f;
              v[j] = z > 0.f ? g0 * w3v : 0.f;
              t[j] = g0 * fmaxf(z, 0.f);                        // rows outside the image have g0 = 0
            }
            // transpose-reduce: lane j ends up with sum over the warp's 32 rows of t[j]
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
              const bool up = (lane & off) != 0;
#pragma unroll
              for (int i = 0; i < off; ++i) {
                const float send = up ? t[i] : t[i + off];
                const float keep = up ? t[i + off] : t[i];
                t[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
              }
            }
            if (lane < ncols) dw3_s[ew * 512 + col0 + lane] += t[0];
          } else {
            const uint4* mrow = reinterpret_cast<const uint4*>(p.mask_hi + out_row * p.C + col0);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint4 m = make_uint4(0, 0, 0, 0);
              if (row_ok && q * 8 < ncols) m = __ldg(mrow + q);
              const uint32_t mw[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                v[q * 8 + 2 * e] = (mw[e] & 0x7fffu) ? __uint_as_float(r[q * 8 + 2 * e]) : 0.f;
                v[q * 8 + 2 * e + 1] = (mw[e] & 0x7fff0000u) ? __uint_as_float(r[q * 8 + 2 * e + 1]) : 0.f;
              }
            }
          }
          if (row_ok) {
            uint4* oh = reinterpret_cast<uint4*>(p.out_hi + out_row * p.C + col0);
            uint4* ol = reinterpret_cast<uint4*>(p.out_lo + out_row * p.C + col0);
            uint4* ol2 = reinterpret_cast<uint4*>(p.out_lo2 + out_row * p.C + col0);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (q * 8 < ncols) {
                uint32_t hi[4], lo[4], lo2[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float a = v[q * 8 + 2 * e], b = v[q * 8 + 2 * e + 1];
                  split_bf16(a, b, hi[e], lo[e]);
                  if (EPI == EPI_RELU_SPLIT) {
                    const float ra = (a - __uint_as_float(hi[e] << 16)) - __uint_as_float(lo[e] << 16);
                    const float rb = (b - __uint_as_float(hi[e] & 0xffff0000u)) - __uint_as_float(lo[e] & 0xffff0000u);
                    const __nv_bfloat162 t2 = __floats2bfloat162_rn(ra, rb);
                    lo2[e] = *reinterpret_cast<const uint32_t*>(&t2);
                  }
                }
                oh[q] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                ol[q] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                if (EPI == EPI_RELU_SPLIT) ol2[q] = make_uint4(lo2[0], lo2[1], lo2[2], lo2[3]);
              }
            }
          }
        } else if (EPI == EPI_RED) {
          if (row_ok) {
            float* o = p.red_out + out_row * p.C + col0;
#pragma unroll
            for (int q = 0; q < 8; ++q)
              if (q * 4 < ncols)
                atomicAdd(reinterpret_cast<float4*>(o) + q,
                          make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]),
                                      __uint_as_float(r[4 * q + 3])));
          }
        } else if (EPI == EPI_TAIL) {  // features[img][col][n] = bf16(acc + bias[col]): the head's input layout
          if (row_ok) {
            uint16_t* o = p.tail_out + (static_cast<long long>(img) * p.C_out + col0) * p.N_img + n_in_img;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < ncols) {
                const float bj = p.tail_bias != nullptr ? __ldg(p.tail_bias + col0 + j) : 0.f;
                o[static_cast<long long>(j) * p.N_img] = f32_to_bf16_rn(__uint_as_float(r[j]) + bj);
              }
          }
        } else {  // EPI_DFEAT: d_q[img][col][n] = acc (the projection term is added by dfeat_proj_kernel)
          if (row_ok) {
            float* o = p.d_feat + (static_cast<long long>(img) * p.C + col0) * p.N_img + n_in_img;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < ncols) o[static_cast<long long>(j) * p.N_img] = __uint_as_float(r[j]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
    if (EPI == EPI_LAYER2) {
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const int t = threadIdx.x - 64;
      for (int c = t; c < p.n_valid; c += 256) {
        float s = 0.f;
#pragma unroll
        for (int wv = 0; wv < EPI_WARPS; ++wv) s += dw3_s[wv * 512 + c];
        atomicAdd(p.dw3 + c, s);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

// fp32 [C][C] -> bf16 hi/lo of the matrix and of its transpose
__global__ void split_weights_kernel(const float* __restrict__ W, int C, uint16_t* hi, uint16_t* lo, uint16_t* hi_t,
                                     uint16_t* lo_t, uint16_t* lo2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C * C) return;
  const float v = W[i];
  const uint16_t h = f32_to_bf16_rn(v);
  const uint16_t l = f32_to_bf16_rn(v - bf16_bits_to_f32(h));
  hi[i] = h; lo[i] = l;
  if (lo2 != nullptr) lo2[i] = f32_to_bf16_rn((v - bf16_bits_to_f32(h)) - bf16_bits_to_f32(l));
  const int r = i / C, c = i - r * C;
  hi_t[c * C + r] = h; lo_t[c * C + r] = l;
}

// Exact fp32 re-evaluation of the queued z2 elements from the 24-bit hidden activations: one warp per element.
//   z2[px][i] = sum_j W2[i][j] h1[px][j];   dz2[px][i] = g0[px] w3[i] [z2 > 0]  (rewritten as bf16 hi/lo)
__global__ void __launch_bounds__(256) mask_fixup_kernel(const int* __restrict__ count, const int* __restrict__ list, int cap,
                                                         const float* __restrict__ W2, const uint16_t* __restrict__ h1h,
                                                         const uint16_t* __restrict__ h1l, const uint16_t* __restrict__ h1l2,
                                                         const float* __restrict__ w3, const float* __restrict__ g, int C,
                                                         int N, int Ktot, int ch, uint16_t* dz2h, uint16_t* dz2l) {
  const int n = min(*count, cap);
  const int lane = threadIdx.x & 31;
  for (int e = blockIdx.x * 8 + (threadIdx.x >> 5); e < n; e += gridDim.x * 8) {
    const int idx = list[e];
    const int px = idx / C, i = idx - px * C;
    const long long ro = static_cast<long long>(px) * C;
    float z = 0.f;
    for (int j = lane; j < C; j += 32) {
      const float h = (bf16_bits_to_f32(h1h[ro + j]) + bf16_bits_to_f32(h1l[ro + j])) + bf16_bits_to_f32(h1l2[ro + j]);
      z = fmaf(__ldg(W2 + static_cast<long long>(i) * C + j), h, z);
    }
    z = warp_sum(z);
    if (lane == 0) {
      const int b = px / N, pn = px - b * N;
      const float v = z > 0.f ? g[(static_cast<long long>(b) * Ktot + ch) * N + pn] * w3[i] : 0.f;
      const uint16_t hb = f32_to_bf16_rn(v);
      dz2h[ro + i] = hb;
      dz2l[ro + i] = f32_to_bf16_rn(v - bf16_bits_to_f32(hb));
    }
  }
}

static int map2d(CUtensorMap* m, const void* ptr, long long inner, long long outer, int box_inner, int box_outer) {
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(inner), static_cast<cuuint64_t>(outer)};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_inner), static_cast<cuuint32_t>(box_outer)};
  return make_map(m, ptr, 2, dims, box);
}
static int map3d(CUtensorMap* m, const void* ptr, int N, int C, int B, int box_inner, int box_mid) {
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(N), static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(B)};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(box_inner), static_cast<cuuint32_t>(box_mid), 1};
  return make_map(m, ptr, 3, dims, box);
}

template <int EPI>
static int launch(const GemmMaps& m, const GemmParams& p, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(tc_gemm_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e != cudaSuccess) return static_cast<int>(e);
  const int items = p.m_tiles * p.n_tiles * p.k_chunks;
  tc_gemm_kernel<EPI><<<items < sl::num_sms() ? items : sl::num_sms(), THREADS, SMEM_BYTES, st>>>(m, p);
  return SL_LAUNCH_RESULT();
}

}  // namespace tcg
}  // namespace sl

// Decoder tail: the final 1x1 convolution of PSPModule.bottleneck (networks/pspnet_pop.py:22)
// as a split-bf16 GEMM whose epilogue adds the bias and stores the head's bf16 NCHW features directly.
//   act_hi/lo [B,Cin,N] bf16 planes of relu(bn(x)) (tails.cu), W_hi/lo [Cout][Cin]; 3 passes, ~1e-5 of fp32.
int sl_tail_gemm_run(const uint16_t* act_hi, const uint16_t* act_lo, int B, int Cin, int N, const uint16_t* W_hi,
                     const uint16_t* W_lo, const float* bias, int Cout, uint16_t* feat_out, cudaStream_t st) {
  using namespace sl::tcg;
  GemmParams p{};
  p.C = Cin; p.N_img = N; p.B = B; p.C_out = Cout;
  p.NT = (Cout + 63) / 64 * 64 < MAX_NT ? (Cout + 63) / 64 * 64 : MAX_NT;
  p.n_tiles = (Cout + p.NT - 1) / p.NT;
  p.n_valid = Cout; p.m_valid = Cout;
  p.m_tiles_per_img = (N + BLOCK_M - 1) / BLOCK_M;
  p.m_is_px = 1; p.k_is_px = 0; p.k_flat = 0;
  p.m_tiles = B * p.m_tiles_per_img; p.k_chunks = 1; p.chunks_per_img = 1; p.chunk_kb = 0;
  p.a_policy = L2_EVICT_FIRST; p.b_policy = L2_EVICT_LAST;
  p.passes = 3; p.a_kind = LD_MN3D; p.b_kind = LD_K2D;
  p.tail_out = feat_out; p.tail_bias = bias;
  GemmMaps m{};
  int rc;
  const uint16_t* as[3] = {act_hi, act_lo, act_hi};
  const uint16_t* bs[3] = {W_hi, W_hi, W_lo};
  for (int i = 0; i < 3; ++i) {
    if ((rc = map3d(&m.a[i], as[i], N, Cin, B, 64, BLOCK_K)) != 0) return rc;
    if ((rc = map2d(&m.b[i], bs[i], Cin, Cout, BLOCK_K, p.NT)) != 0) return rc;
  }
  return launch<EPI_TAIL>(m, p, st);
}

// capacity of the sign-fix-up queue: ~0.1 % of the elements are expected, 1.6 % fit (the rest stay as computed)
int sl_pop_bwd_flag_cap(long long px, int C) {
  const long long cap = px * C / 64;
  return static_cast<int>(cap < 4096 ? 4096 : (cap > (1 << 26) ? (1 << 26) : cap));
}

// Workspace layout (bytes): see sl_pop_head_bwd_ws_bytes.  Called by sl_pop_head_bwd after the foreground part.
int sl_pop_bwd_tc_run(const uint16_t* feat, int B, int C, int N, const float* s_hat, int K, const float* W1p, const float* W2,
                      const float* w3, const float* g_logits, int Ktot, int bg_ch, const float* gp, float* dW1p,
                      float* dW2, float* dw3, float* d_feat, uint16_t* act, uint16_t* wsplit, cudaStream_t st) {
  using namespace sl::tcg;
  const long long px = static_cast<long long>(B) * N;
  uint16_t* h1h = act;            uint16_t* h1l = h1h + px * C;
  uint16_t* z2h = h1l + px * C;   uint16_t* z2l = z2h + px * C;     // dz2
  uint16_t* z1h = z2l + px * C;   uint16_t* z1l = z1h + px * C;     // dz1
  uint16_t* h1l2 = z1l + px * C;                                    // third term of h1 (24 mantissa bits in all)
  const long long CC = static_cast<long long>(C) * C;
  uint16_t* w1h = wsplit;        uint16_t* w1l = w1h + CC; uint16_t* w1ht = w1l + CC; uint16_t* w1lt = w1ht + CC;
  uint16_t* w2h = w1lt + CC;     uint16_t* w2l = w2h + CC; uint16_t* w2ht = w2l + CC; uint16_t* w2lt = w2ht + CC;
  uint16_t* w1l2 = w2lt + CC;
  int* flag_count = reinterpret_cast<int*>(w1l2 + CC);
  int* flag_list = flag_count + 4;
  const int flag_cap = sl_pop_bwd_flag_cap(px, C);
  cudaMemsetAsync(flag_count, 0, sizeof(int), st);
  split_weights_kernel<<<(C * C + 255) / 256, 256, 0, st>>>(W1p, C, w1h, w1l, w1ht, w1lt, w1l2);
  split_weights_kernel<<<(C * C + 255) / 256, 256, 0, st>>>(W2, C, w2h, w2l, w2ht, w2lt, nullptr);

  GemmParams base{};
  base.C = C; base.N_img = N; base.B = B;
  base.NT = (C + 63) / 64 * 64 < MAX_NT ? (C + 63) / 64 * 64 : MAX_NT;
  base.n_tiles = (C + base.NT - 1) / base.NT;
  base.n_valid = C; base.m_valid = C;
  base.m_tiles_per_img = (N + BLOCK_M - 1) / BLOCK_M;
  base.Ktot = Ktot; base.ch = bg_ch; base.K = K;
  const int NT = base.NT;
  int rc;
#define SL_TRY(x) do { if ((rc = (x)) != 0) return rc; } while (0)

  // pixel-major GEMMs: M = pixels (per image), K = channels
  auto px_gemm = [&](GemmParams& p) {
    p.m_is_px = 1; p.k_is_px = 0; p.k_flat = 0;
    p.m_tiles = B * base.m_tiles_per_img; p.k_chunks = 1; p.chunks_per_img = 1; p.chunk_kb = 0;
    p.a_policy = L2_EVICT_FIRST; p.b_policy = L2_EVICT_LAST;
  };
  {  // G1: h1 = relu(q W1'^T); three weight terms = 24 mantissa bits, so sign(z1) is decided at fp32 precision
    GemmMaps m{}; GemmParams p = base; px_gemm(p);
    p.passes = 3; p.a_kind = LD_MN3D; p.b_kind = LD_K2D;
    for (int i = 0; i < 3; ++i) SL_TRY(map3d(&m.a[i], feat, N, C, B, 64, BLOCK_K));
    SL_TRY(map2d(&m.b[0], w1h, C, C, BLOCK_K, NT));
    SL_TRY(map2d(&m.b[1], w1l, C, C, BLOCK_K, NT));
    SL_TRY(map2d(&m.b[2], w1l2, C, C, BLOCK_K, NT));
    p.out_hi = h1h; p.out_lo = h1l; p.out_lo2 = h1l2;
    SL_TRY(launch<EPI_RELU_SPLIT>(m, p, st));
  }
  {  // G2: z2 = h1 W2^T -> dz2 (hi/lo), dw3
    GemmMaps m{}; GemmParams p = base; px_gemm(p);
    p.passes = 3; p.a_kind = LD_K2D; p.b_kind = LD_K2D;
    SL_TRY(map2d(&m.a[0], h1h, C, px, BLOCK_K, BLOCK_M)); SL_TRY(map2d(&m.b[0], w2h, C, C, BLOCK_K, NT));
    SL_TRY(map2d(&m.a[1], h1l, C, px, BLOCK_K, BLOCK_M)); SL_TRY(map2d(&m.b[1], w2h, C, C, BLOCK_K, NT));
    SL_TRY(map2d(&m.a[2], h1h, C, px, BLOCK_K, BLOCK_M)); SL_TRY(map2d(&m.b[2], w2l, C, C, BLOCK_K, NT));
    p.out_hi = z2h; p.out_lo = z2l; p.w3 = w3; p.g = g_logits; p.dw3 = dw3;
    p.flag_count = flag_count; p.flag_list = flag_list; p.flag_cap = flag_cap;
    SL_TRY(launch<EPI_LAYER2>(m, p, st));
    mask_fixup_kernel<<<2 * sl::num_sms(), 256, 0, st>>>(flag_count, flag_list, flag_cap, W2, h1h, h1l, h1l2, w3, g_logits, C, N,
                                                      Ktot, bg_ch, z2h, z2l);
  }
  {  // G3: dz1 = (dz2 W2) [h1 > 0]   (B = W2^T, K-major)
    GemmMaps m{}; GemmParams p = base; px_gemm(p);
    p.passes = 3; p.a_kind = LD_K2D; p.b_kind = LD_K2D;
    SL_TRY(map2d(&m.a[0], z2h, C, px, BLOCK_K, BLOCK_M)); SL_TRY(map2d(&m.b[0], w2ht, C, C, BLOCK_K, NT));
    SL_TRY(map2d(&m.a[1], z2l, C, px, BLOCK_K, BLOCK_M)); SL_TRY(map2d(&m.b[1], w2ht, C, C, BLOCK_K, NT));
    SL_TRY(map2d(&m.a[2], z2h, C, px, BLOCK_K, BLOCK_M)); SL_TRY(map2d(&m.b[2], w2lt, C, C, BLOCK_K, NT));
    p.out_hi = z1h; p.out_lo = z1l; p.mask_hi = h1h;
    SL_TRY(launch<EPI_MASK_SPLIT>(m, p, st));
  }
  // channel-major GEMMs: M = N = channels, K = pixels, split over pixel chunks, fp32 atomics
  const int mt_c = (C + BLOCK_M - 1) / BLOCK_M;
  auto ch_gemm = [&](GemmParams& p, bool flat) {
    p.m_is_px = 0; p.k_is_px = 1; p.k_flat = flat ? 1 : 0;
    p.m_tiles = mt_c;
    const long long extent = flat ? px : N;
    const int kb_total = static_cast<int>((extent + BLOCK_K - 1) / BLOCK_K);
    const int units = flat ? 1 : B;
    int want = 2 * sl::num_sms() / (mt_c * p.n_tiles * units);          // chunks per image (or in total when flat)
    if (want < 1) want = 1;
    if (want > kb_total) want = kb_total;
    p.chunk_kb = (kb_total + want - 1) / want;
    p.chunks_per_img = (kb_total + p.chunk_kb - 1) / p.chunk_kb;
    p.k_chunks = units * p.chunks_per_img;
    p.a_policy = L2_EVICT_NORMAL; p.b_policy = L2_EVICT_NORMAL;
  };
  {  // G4: dW2[i][j] = sum_px dz2[px][i] h1[px][j]
    GemmMaps m{}; GemmParams p = base; ch_gemm(p, true);
    p.passes = 3; p.a_kind = LD_MN2D; p.b_kind = LD_MN2D;
    SL_TRY(map2d(&m.a[0], z2h, C, px, 64, BLOCK_K)); SL_TRY(map2d(&m.b[0], h1h, C, px, 64, BLOCK_K));
    SL_TRY(map2d(&m.a[1], z2l, C, px, 64, BLOCK_K)); SL_TRY(map2d(&m.b[1], h1h, C, px, 64, BLOCK_K));
    SL_TRY(map2d(&m.a[2], z2h, C, px, 64, BLOCK_K)); SL_TRY(map2d(&m.b[2], h1l, C, px, 64, BLOCK_K));
    p.red_out = dW2;
    SL_TRY(launch<EPI_RED>(m, p, st));
  }
  {  // G5: dW1'[i][j] = sum_px dz1[px][i] q[px][j]
    GemmMaps m{}; GemmParams p = base; ch_gemm(p, false);
    p.passes = 2; p.a_kind = LD_MN2D; p.b_kind = LD_K3D;
    SL_TRY(map2d(&m.a[0], z1h, C, px, 64, BLOCK_K)); SL_TRY(map3d(&m.b[0], feat, N, C, B, BLOCK_K, NT));
    SL_TRY(map2d(&m.a[1], z1l, C, px, 64, BLOCK_K)); SL_TRY(map3d(&m.b[1], feat, N, C, B, BLOCK_K, NT));
    p.red_out = dW1p;
    SL_TRY(launch<EPI_RED>(m, p, st));
  }
  if (d_feat != nullptr) {  // G6: d_q = dz1 W1' (+ projections); B = W1'^T, K-major
    GemmMaps m{}; GemmParams p = base; px_gemm(p);
    p.passes = 3; p.a_kind = LD_K2D; p.b_kind = LD_K2D;
    SL_TRY(map2d(&m.a[0], z1h, C, px, BLOCK_K, BLOCK_M)); SL_TRY(map2d(&m.b[0], w1ht, C, C, BLOCK_K, NT));
    SL_TRY(map2d(&m.a[1], z1l, C, px, BLOCK_K, BLOCK_M)); SL_TRY(map2d(&m.b[1], w1ht, C, C, BLOCK_K, NT));
    SL_TRY(map2d(&m.a[2], z1h, C, px, BLOCK_K, BLOCK_M)); SL_TRY(map2d(&m.b[2], w1lt, C, C, BLOCK_K, NT));
    p.d_feat = d_feat; p.gp = gp; p.s_hat = s_hat;
    SL_TRY(launch<EPI_DFEAT>(m, p, st));
  }
#undef SL_TRY
  return 0;
}
