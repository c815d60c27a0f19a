// Pruned up-sampling + argmax (+ confusion) for the prediction-only path of sl_upsample_argmax
//   F.interpolate(bilinear, align_corners=True) -> argmax -> get_confusion_matrix
//   (eval_base.py:168-178, eval_ft.py:168-183, ft_pop.py:327-331).
//
// The row-cached kernel in postproc.cu evaluates all K classes at every output pixel and is bound by instruction
// issue (ncu: 10.7 lane-instructions per pixel-class, DRAM 1.6 % busy).  Most of that work cannot change the result:
// inside one source cell (the output pixels between four neighbouring low-res pixels) every class value is
//     v_k = l0y*(l0x*a_k + l1x*b_k) + l1y*(l0x*c_k + l1x*d_k),   weights >= 0,
// and fp32 multiply / fused multiply-add are monotone in each operand, so the order of two classes at the four
// corners carries over to every pixel of the cell:
//   * c < j and corner values of c >= those of j at all four corners  =>  v_c >= v_j everywhere, and on a tie the
//     first-maximum rule (np.argmax) picks c: class j can never be the argmax in this cell;
//   * c > j needs a strict inequality at every pixel; it is guaranteed when c exceeds j at all four corners by a
//     margin of 2^-19 of the largest magnitude in the cell (>= 4x the worst-case rounding error of the three-level
//     expression, |error| <= 4 * 2^-24 * max|corner| per class; the weights sum to 1 within 2^-23).
// "c precedes j at every pixel" is a strict order per pixel, so dropping every class that is preceded by some other
// class keeps the true argmax among the survivors, and evaluating the survivors in ascending index order with a
// strict compare returns it: the prediction map is IDENTICAL to evaluating all K classes.  Cells holding a non-finite
// (or > 1e37) value keep every class and use the NaN-aware compare.
//
// Layout: one CTA = one band (the output rows sharing a top source row) x 1024 output columns.  The two source rows
// are staged in shared memory, one thread per cell computes the survivor mask, then one thread per 4-pixel-wide strip
// either writes its single surviving class for all rows (homogeneous regions: no arithmetic at all) or queues the strip
// on a work list bucketed by survivor count; the lists are then worked off with full warps, so the divergence a
// per-warp skip would suffer (a warp spans 16 cells, some of them always on a class boundary) does not arise.
// The confusion matrix is counted in the same pass from the packed label / prediction words (run-length per thread).
#include "common.cuh"

namespace sl {

constexpr int PR_THREADS = 256;
constexpr int PR_ROWS = 5;            // output rows per work item
constexpr int PR_MAX_BAND = 10;       // rows per band (up-sampling factors below 10)
constexpr uint32_t PR_NF = 0x80000000u;

__device__ __forceinline__ float2 pr_mul2(const float2 a, const float2 b) {
  float2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<const unsigned long long&>(a)), "l"(reinterpret_cast<const unsigned long long&>(b)));
  return d;
}
__device__ __forceinline__ float2 pr_fma2(const float2 a, const float2 b, const float2 c) {
  float2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<const unsigned long long&>(a)), "l"(reinterpret_cast<const unsigned long long&>(b)),
        "l"(reinterpret_cast<const unsigned long long&>(c)));
  return d;
}

// first output coordinate whose source index (src_coord(scale, ., in_size).i0) is >= s
__device__ __forceinline__ int first_dst(float scale, int s, int in_size, int out_size) {
  if (s <= 0) return 0;
  if (s >= in_size) return out_size;
  int d = scale > 0.f ? static_cast<int>(ceilf(static_cast<float>(s) / scale)) : out_size;
  d = max(0, min(out_size, d));
  while (d > 0 && src_coord(scale, d - 1, in_size).i0 >= s) --d;
  while (d < out_size && src_coord(scale, d, in_size).i0 < s) ++d;
  return d;
}

// Confusion counting of one packed word of four (label, prediction) pairs with a per-thread run-length accumulator.
struct CmRun {
  int bin;
  unsigned int cnt;
  __device__ __forceinline__ void flush(unsigned int* hist) { if (cnt) atomicAdd(&hist[bin], cnt); cnt = 0; }
  __device__ __forceinline__ void add_word(unsigned int* hist, uint32_t lw, uint32_t pw, int K, int ignore_label) {
    const uint32_t l0 = lw & 0xffu, p0 = pw & 0xffu;
    if (lw == l0 * 0x01010101u && pw == p0 * 0x01010101u) {
      if (static_cast<int>(l0) == ignore_label || static_cast<int>(l0) >= K) return;
      const int b = static_cast<int>(l0) * K + static_cast<int>(p0);
      if (b == bin) { cnt += 4u; return; }
      flush(hist);
      bin = b; cnt = 4u;
      return;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int l = static_cast<int>((lw >> (8 * j)) & 0xffu), p = static_cast<int>((pw >> (8 * j)) & 0xffu);
      if (l == ignore_label || l >= K) continue;
      const int b = l * K + p;
      if (b == bin) ++cnt; else atomicAdd(&hist[b], 1u);
    }
  }
};

// Classes that can be the argmax somewhere in a cell, from the cell's four corner values get(corner, k) (see the
// header comment for the argument); bit 31 flags a cell with a non-finite / huge value (all classes kept).
template <int KT, class Get>
__device__ __forceinline__ uint32_t survivor_mask(int K, Get get) {
  constexpr int KU = KT > 0 ? KT : 1;
  const int Kn = KT > 0 ? KT : K;
  float mx = 0.f, z = 0.f, lead_v = -INFINITY;
  int lead = 0;                                                 // first maximum at corner 0
#pragma unroll(KU)
  for (int k = 0; k < Kn; ++k) {
    const float a = get(0, k), b = get(1, k), c = get(2, k), d = get(3, k);
    mx = fmaxf(mx, fmaxf(fmaxf(fabsf(a), fabsf(b)), fmaxf(fabsf(c), fabsf(d))));
    z += (a - a) + (b - b) + (c - c) + (d - d);                 // 0 for finite values, NaN for NaN / inf (fmaxf drops NaNs)
    if (a > lead_v) { lead_v = a; lead = k; }
  }
  const uint32_t all = (1u << Kn) - 1u;                         // K <= 31
  if (!(z == 0.f) || !(mx < 1e37f)) return all | PR_NF;
  const float margin = fmaxf(mx * 1.9073486328125e-6f /* 2^-19 */, 1e-30f);
  auto beats = [&](int c, int j) {                              // c precedes j at every pixel of the cell
    bool r = true;
#pragma unroll
    for (int q = 0; q < 4; ++q) r = r && (c < j ? get(q, c) >= get(q, j) : get(q, c) - get(q, j) >= margin);
    return r;
  };
  // common case first: the leading class precedes every other class
  bool single = true;
  if constexpr (KT > 0) {
    // `lead` is a run-time index: select its corner values once so the register arrays keep static indices
    float lv[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      lv[q] = get(q, 0);
#pragma unroll
      for (int k = 1; k < KT; ++k) lv[q] = (k == lead) ? get(q, k) : lv[q];
    }
#pragma unroll
    for (int j = 0; j < KT; ++j) {
      bool r = true;
#pragma unroll
      for (int q = 0; q < 4; ++q) r = r && (lead < j ? lv[q] >= get(q, j) : lv[q] - get(q, j) >= margin);
      single = single && (j == lead || r);
    }
  } else {
    for (int j = 0; j < Kn; ++j) single = single && (j == lead || beats(lead, j));
  }
  if (single) return 1u << lead;
  uint32_t mask = 0u;
#pragma unroll(KU)
  for (int j = 0; j < Kn; ++j) {
    bool dead = false;
#pragma unroll(KU)
    for (int c = 0; c < Kn; ++c)
      if (c != j) dead = dead || beats(c, j);
    if (!dead) mask |= 1u << j;
  }
  return mask;
}

// KT > 0: compile-time class count (loops unrolled, corner values in registers); KT == 0: any K <= SL_MAX_CLASSES.
template <int KT>
__global__ void __launch_bounds__(PR_THREADS, 3) upsample_prune_kernel(
    const float* __restrict__ logits_lr, int K_rt, int h, int w, int H, int W, float sy, float sx,
    const uint8_t* __restrict__ label, int ignore_label, uint8_t* __restrict__ pred,
    unsigned long long* __restrict__ cm, int ncols_alloc) {
  const int K = KT > 0 ? KT : K_rt;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // layout: raw [2][K][ncols_alloc] fp32 | mask [ncols_alloc] u32 | lists [4][2*PR_THREADS] u16 | hist [K*K] u32
  float* raw = reinterpret_cast<float*>(smem_raw);
  uint32_t* cmask = reinterpret_cast<uint32_t*>(raw + 2 * K * ncols_alloc);
  uint16_t* lists = reinterpret_cast<uint16_t*>(cmask + ncols_alloc);
  unsigned int* hist = reinterpret_cast<unsigned int*>(lists + 4 * 2 * PR_THREADS);
  __shared__ int list_n[4];
  __shared__ float row_l0[PR_MAX_BAND], row_l1[PR_MAX_BAND];

  const int tid = threadIdx.x;
  const int b = blockIdx.z, band = blockIdx.y;
  const bool do_cm = cm != nullptr;
  const int X0 = blockIdx.x * (PR_THREADS * 4);
  const int X1 = min(W, X0 + PR_THREADS * 4);
  const int y_lo = first_dst(sy, band, h, H), y_hi = first_dst(sy, band + 1, h, H);   // rows [y_lo, y_hi)
  const int R = y_hi - y_lo;
  if (R <= 0) return;
  const int row1 = min(band + 1, h - 1);
  const int c_lo = src_coord(sx, X0, w).i0;
  const SrcCoord c_last = src_coord(sx, X1 - 1, w);
  const int c_hi = c_last.i0 + c_last.step;
  const int ncols = c_hi - c_lo + 1;                            // <= ncols_alloc by construction of the launch

  if (tid < 4) list_n[tid] = 0;
  if (tid < R) {
    const SrcCoord cy = src_coord(sy, y_lo + tid, h);
    row_l0[tid] = cy.l0; row_l1[tid] = cy.l1;
  }
  if (do_cm) for (int i = tid; i < K * K; i += PR_THREADS) hist[i] = 0u;
  {  // stage the two source rows of every class (coalesced; the low-res logits are L2-resident)
    const size_t hw = static_cast<size_t>(h) * w;
    const float* p0 = logits_lr + static_cast<size_t>(b) * K * hw + static_cast<size_t>(band) * w + c_lo;
    const float* p1 = logits_lr + static_cast<size_t>(b) * K * hw + static_cast<size_t>(row1) * w + c_lo;
    for (int k = 0; k < K; ++k, p0 += hw, p1 += hw)
      for (int c = tid; c < ncols; c += PR_THREADS) {
        raw[(0 * K + k) * ncols_alloc + c] = __ldg(p0 + c);
        raw[(1 * K + k) * ncols_alloc + c] = __ldg(p1 + c);
      }
  }
  __syncthreads();

  // ---- survivor mask per cell (cell c = source columns c, min(c + 1, c_hi))
  for (int c = tid; c < ncols; c += PR_THREADS) {
    const int c1 = min(c + 1, ncols - 1);
    const float* q0 = raw + c;
    const float* q1 = raw + c1;
    if constexpr (KT > 0) {
      float v[4][KT];                                           // corner values in registers
#pragma unroll
      for (int k = 0; k < KT; ++k) {
        v[0][k] = q0[(0 * KT + k) * ncols_alloc]; v[1][k] = q1[(0 * KT + k) * ncols_alloc];
        v[2][k] = q0[(1 * KT + k) * ncols_alloc]; v[3][k] = q1[(1 * KT + k) * ncols_alloc];
      }
      cmask[c] = survivor_mask<KT>(KT, [&](int corner, int k) { return v[corner][k]; });
    } else {
      cmask[c] = survivor_mask<0>(K, [&](int corner, int k) {
        return ((corner & 1) ? q1 : q0)[((corner >> 1) * K + k) * ncols_alloc];
      });
    }
  }
  __syncthreads();

  // ---- one strip (4 output columns x the band's rows) per thread
  const int x0 = X0 + tid * 4;
  const bool col_ok = x0 < W;                                   // W % 4 == 0
  CmRun run; run.bin = 0; run.cnt = 0;
  const int nchunks = (R + PR_ROWS - 1) / PR_ROWS;
  if (col_ok) {
    const int ca = src_coord(sx, x0, w).i0 - c_lo, cb = src_coord(sx, x0 + 3, w).i0 - c_lo;
    uint32_t m = 0u;
    for (int c = ca; c <= cb; ++c) m |= cmask[c];
    const int nc = __popc(m & ~PR_NF);
    if (nc == 1 && !(m & PR_NF)) {
      const uint32_t pw = static_cast<uint32_t>(__ffs(m) - 1) * 0x01010101u;
      const size_t pix = (static_cast<size_t>(b) * H + y_lo) * W + x0;
      uint32_t lw[PR_MAX_BAND];
      if (do_cm) {
#pragma unroll
        for (int r = 0; r < PR_MAX_BAND; ++r)
          if (r < R) lw[r] = __ldg(reinterpret_cast<const uint32_t*>(label + pix + static_cast<size_t>(r) * W));
      }
#pragma unroll
      for (int r = 0; r < PR_MAX_BAND; ++r)
        if (r < R) *reinterpret_cast<uint32_t*>(pred + pix + static_cast<size_t>(r) * W) = pw;
      if (do_cm) {
#pragma unroll
        for (int r = 0; r < PR_MAX_BAND; ++r)
          if (r < R) run.add_word(hist, lw[r], pw, K, ignore_label);
      }
    } else {
      const int bucket = (m & PR_NF) ? 3 : nc <= 2 ? 0 : nc <= 4 ? 1 : 2;
      const int slot = atomicAdd(&list_n[bucket], nchunks);
      for (int q = 0; q < nchunks; ++q) lists[bucket * 2 * PR_THREADS + slot + q] = static_cast<uint16_t>(tid | (q << 8));
    }
  }
  __syncthreads();

  // ---- queued strips, one (strip, row chunk) item per thread and turn; items of one bucket run together
  for (int bucket = 0; bucket < 4; ++bucket) {
    const int n_items = list_n[bucket];
    for (int it = tid; it < n_items; it += PR_THREADS) {
      const int item = lists[bucket * 2 * PR_THREADS + it];
      const int st = item & 0xff, r0 = (item >> 8) * PR_ROWS;
      const int nr = min(PR_ROWS, R - r0);
      const int xs0 = X0 + st * 4;
      int xi[4], xn[4];
      float xl0[4], xl1[4];
      uint32_t m = 0u;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const SrcCoord c = src_coord(sx, xs0 + j, w);
        xi[j] = c.i0 - c_lo; xn[j] = xi[j] + c.step; xl0[j] = c.l0; xl1[j] = c.l1;
      }
      for (int c = xi[0]; c <= xi[3]; ++c) m |= cmask[c];
      float best[PR_ROWS][4];
      uint32_t idx[PR_ROWS];                                    // four class indices per row, one byte each
#pragma unroll
      for (int r = 0; r < PR_ROWS; ++r) {
        idx[r] = 0u;
#pragma unroll
        for (int j = 0; j < 4; ++j) best[r][j] = -INFINITY;
      }
      const bool nf = (m & PR_NF) != 0u;
      uint32_t todo = nf ? (K >= 32 ? 0xffffffffu : ((1u << K) - 1u)) : m;
      bool first = !nf;
      while (todo) {
        const int k = __ffs(todo) - 1;
        todo &= todo - 1u;
        const uint32_t kk = static_cast<uint32_t>(k) * 0x01010101u;
        const float* t = raw + (0 * K + k) * ncols_alloc;
        const float* u = raw + (1 * K + k) * ncols_alloc;
        float top[4], bot[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          // l0x*a + l1x*b as nvcc contracts it in the row-cached kernel: fma(l0x, a, l1x*b)
          top[j] = __fmaf_rn(xl0[j], t[xi[j]], __fmul_rn(xl1[j], t[xn[j]]));
          bot[j] = __fmaf_rn(xl0[j], u[xi[j]], __fmul_rn(xl1[j], u[xn[j]]));
        }
#pragma unroll
        for (int r = 0; r < PR_ROWS; ++r) {
          if (r < nr) {
            const float l0 = row_l0[r0 + r], l1 = row_l1[r0 + r];
            // l0y*top + l1y*bot on packed pairs: mul then fma, the same products and sums as the row-cached kernel
            float2 v01 = pr_mul2(make_float2(l0, l0), make_float2(top[0], top[1]));
            float2 v23 = pr_mul2(make_float2(l0, l0), make_float2(top[2], top[3]));
            v01 = pr_fma2(make_float2(l1, l1), make_float2(bot[0], bot[1]), v01);
            v23 = pr_fma2(make_float2(l1, l1), make_float2(bot[2], bot[3]), v23);
            const float v[4] = {v01.x, v01.y, v23.x, v23.y};
            if (first) {
              idx[r] = kk;
#pragma unroll
              for (int j = 0; j < 4; ++j) best[r][j] = v[j];
            } else if (!nf) {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (v[j] > best[r][j]) { best[r][j] = v[j]; idx[r] = (idx[r] & ~(0xffu << (8 * j))) | (kk & (0xffu << (8 * j))); }
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j)                      // np.argmax: first maximum, NaN counts as maximal
                if (v[j] > best[r][j] || (v[j] != v[j] && best[r][j] == best[r][j])) {
                  best[r][j] = v[j];
                  idx[r] = (idx[r] & ~(0xffu << (8 * j))) | (kk & (0xffu << (8 * j)));
                }
            }
          }
        }
        first = false;
      }
      const size_t pix = (static_cast<size_t>(b) * H + y_lo + r0) * W + xs0;
#pragma unroll
      for (int r = 0; r < PR_ROWS; ++r) {
        if (r < nr) {
          const uint32_t pw = idx[r];
          *reinterpret_cast<uint32_t*>(pred + pix + static_cast<size_t>(r) * W) = pw;
          if (do_cm)
            run.add_word(hist, __ldg(reinterpret_cast<const uint32_t*>(label + pix + static_cast<size_t>(r) * W)), pw, K,
                         ignore_label);
        }
      }
    }
  }
  if (do_cm) {
    run.flush(hist);
    __syncthreads();
    for (int i = tid; i < K * K; i += PR_THREADS)
      if (hist[i]) atomicAdd(&cm[i], static_cast<unsigned long long>(hist[i]));
  }
}

// Eligibility + launch.  Returns -100 when the shape is outside this kernel's range (the caller then uses the
// row-cached kernel), otherwise the launch result.
int launch_upsample_prune(const float* logits_lr, int B, int K, int h, int w, int H, int W, float sy, float sx,
                          const uint8_t* label, int ignore_label, uint8_t* pred, unsigned long long* cm,
                          cudaStream_t st) {
  // up-sampling by 2x .. <10x in both directions, whole 4-pixel strips, a prediction map to write
  if (pred == nullptr || W % 4 != 0 || B > 65535 || h > 65535 || K < 2 || K > 31) return -100;
  if (!(sx > 0.f && sx <= 0.5f && sy > 0.1f && sy <= 0.5f)) return -100;
  if (static_cast<int>(1.f / sy) + 2 > sl::PR_MAX_BAND) return -100;
  const int ncols_alloc = static_cast<int>(PR_THREADS * 4 * sx) + 4;
  const size_t smem = static_cast<size_t>(2) * K * ncols_alloc * 4 + static_cast<size_t>(ncols_alloc) * 4 +
                      4 * 2 * PR_THREADS * 2 + static_cast<size_t>(K) * K * 4;
  if (smem > 160 * 1024) return -100;
  const dim3 grid((W / 4 + PR_THREADS - 1) / PR_THREADS, h, B);
  auto launch = [&](auto kern) {
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
      if (e != cudaSuccess) return static_cast<int>(e);
    }
    kern<<<grid, PR_THREADS, smem, st>>>(logits_lr, K, h, w, H, W, sy, sx, label, ignore_label, pred, cm, ncols_alloc);
    return SL_LAUNCH_RESULT();
  };
  if (K == 8) return launch(upsample_prune_kernel<8>);
  if (K == 12) return launch(upsample_prune_kernel<12>);
  return launch(upsample_prune_kernel<0>);
}

}  // namespace sl
