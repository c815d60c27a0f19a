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
constexpr int PR_BUCKETS = 6;         // survivor counts 2, 3, 4, 5-6, >= 7, and non-finite cells

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
    // common cases without a per-byte loop: the four predictions are equal and the labels are one class, possibly with
    // ignore_label bytes sprinkled in (OEM tiles: ~1 % ignore pixels)
    const uint32_t p0 = pw & 0xffu;
    if (pw == p0 * 0x01010101u) {
      const uint32_t ign = __vcmpeq4(lw, static_cast<uint32_t>(ignore_label & 0xff) * 0x01010101u);   // 0xff per ignored byte
      if (ign == 0xffffffffu) return;
      const uint32_t l0 = (lw >> ((__ffs(~ign) - 1) & 24)) & 0xffu;        // first non-ignored label
      if ((((lw ^ (l0 * 0x01010101u)) & ~ign) == 0u)) {
        if (static_cast<int>(l0) >= K) return;
        const unsigned int n = 4u - (__popc(ign) >> 3);
        const int b = static_cast<int>(l0) * K + static_cast<int>(p0);
        if (b == bin) { cnt += n; return; }
        flush(hist);
        bin = b; cnt = n;
        return;
      }
    }
    slow(hist, lw, pw, K, ignore_label);
  }
  // mixed word: per byte, kept out of line so that the (many, unrolled) call sites stay small
  __device__ __noinline__ void slow(unsigned int* hist, uint32_t lw, uint32_t pw, int K, int ignore_label) {
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
//   pass 1, O(K): with m_k / M_k the smallest / largest corner of class k and c0 the class with the largest m, every
//           class whose M lies below m_c0 is preceded by c0 at every pixel (a special case of the corner rule);
//   pass 2, only when 2..4 classes are left: the corner rule itself between the remaining pairs.
// Any superset of the true survivors gives the same prediction, so both passes may stop early.
template <int KT, class Get>
__device__ __forceinline__ uint32_t survivor_mask(int K, Get get) {
  constexpr int KU = KT > 0 ? KT : 1;
  const int Kn = KT > 0 ? KT : K;
  float mx = 0.f, z = 0.f, lead_m = -INFINITY;
  int lead = 0;                                                 // first class with the largest minimum corner
#pragma unroll(KU)
  for (int k = 0; k < Kn; ++k) {
    const float a = get(0, k), b = get(1, k), c = get(2, k), d = get(3, k);
    const float hi = fmaxf(fmaxf(a, b), fmaxf(c, d)), lo = fminf(fminf(a, b), fminf(c, d));
    mx = fmaxf(mx, fmaxf(fabsf(hi), fabsf(lo)));
    z += (a - a) + (b - b) + (c - c) + (d - d);                 // 0 for finite values, NaN for NaN / inf (fmaxf drops NaNs)
    if (lo > lead_m) { lead_m = lo; lead = k; }
  }
  const uint32_t all = (1u << Kn) - 1u;                         // K <= 31
  if (!(z == 0.f) || !(mx < 1e37f)) return all | PR_NF;
  const float margin = fmaxf(mx * 1.9073486328125e-6f /* 2^-19 */, 1e-30f);
  uint32_t S = 0u;
#pragma unroll(KU)
  for (int k = 0; k < Kn; ++k) {
    const float hi = fmaxf(fmaxf(get(0, k), get(1, k)), fmaxf(get(2, k), get(3, k)));
    // kept unless the leading class precedes it: weakly when the leader has the lower index, by the margin otherwise
    const bool dead = k != lead && (lead < k ? lead_m >= hi : lead_m - hi >= margin);
    if (!dead) S |= 1u << k;
  }
  const int ns = __popc(S);
  if (ns >= 2 && ns <= 4) {
    const uint32_t S0 = S;
    for (uint32_t js = S0; js; js &= js - 1u) {
      const int j = __ffs(js) - 1;
      const float j0 = get(0, j), j1 = get(1, j), j2 = get(2, j), j3 = get(3, j);
      for (uint32_t cs = S0 & ~(1u << j); cs; cs &= cs - 1u) {
        const int c = __ffs(cs) - 1;
        const float c0 = get(0, c), c1 = get(1, c), c2 = get(2, c), c3 = get(3, c);
        const bool beats = c < j ? (c0 >= j0 && c1 >= j1 && c2 >= j2 && c3 >= j3)
                                 : (c0 - j0 >= margin && c1 - j1 >= margin && c2 - j2 >= margin && c3 - j3 >= margin);
        if (beats) { S &= ~(1u << j); break; }
      }
    }
  }
  return S;
}

// One work item of the queued-strip phase: 4 output columns x nr <= PR_ROWS rows, the surviving classes of `todo` evaluated in
// ascending index order with the same arithmetic as the row-cached kernel: top = fma(l0x, a, l1x*b) per source row,
// value = fma(l1y, bot, l0y*top).  NF: a non-finite value is around -> NaN-aware compare (np.argmax: first maximum,
// NaN counts as maximal).  Returns the four class indices of every row, one byte each.
template <bool NF>
__device__ __forceinline__ void eval_item(int nr, const float* __restrict__ raw, int K, int ncols_alloc, uint32_t todo,
                                          const int (&xi)[4], const int (&xn)[4], const float (&xl0)[4],
                                          const float (&xl1)[4], const float* __restrict__ rl0,
                                          const float* __restrict__ rl1, uint32_t (&idx)[PR_ROWS]) {
  constexpr int NR = PR_ROWS;                    // rows >= nr are computed on the (valid) weights of row nr-1 and never stored
  float best[NR][4];
  int ro[NR];                                    // row weight index (shared-memory broadcast loads in the loops)
#pragma unroll
  for (int r = 0; r < NR; ++r) { ro[r] = min(r, nr - 1); idx[r] = 0u; }
  auto L0 = [&](int r) { return rl0[ro[r]]; };
  auto L1 = [&](int r) { return rl1[ro[r]]; };
  bool first = !NF;
  if (NF) {
#pragma unroll
    for (int r = 0; r < NR; ++r)
#pragma unroll
      for (int j = 0; j < 4; ++j) best[r][j] = -INFINITY;
  }
  while (todo) {
    const int k = __ffs(todo) - 1;
    todo &= todo - 1u;
    const uint32_t kk = static_cast<uint32_t>(k) * 0x01010101u;
    const float* t = raw + k * ncols_alloc;
    const float* u = t + K * ncols_alloc;
    float top[4], bot[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      top[j] = __fmaf_rn(xl0[j], t[xi[j]], __fmul_rn(xl1[j], t[xn[j]]));
      bot[j] = __fmaf_rn(xl0[j], u[xi[j]], __fmul_rn(xl1[j], u[xn[j]]));
    }
    if (first) {
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        idx[r] = kk;
#pragma unroll
        for (int j = 0; j < 4; ++j) best[r][j] = __fmaf_rn(L1(r), bot[j], __fmul_rn(L0(r), top[j]));
      }
      first = false;
    } else {
#pragma unroll
      for (int r = 0; r < NR; ++r)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float v = __fmaf_rn(L1(r), bot[j], __fmul_rn(L0(r), top[j]));
          const bool take = NF ? (v > best[r][j] || (v != v && best[r][j] == best[r][j])) : (v > best[r][j]);
          if (take) { best[r][j] = v; idx[r] = (idx[r] & ~(0xffu << (8 * j))) | (kk & (0xffu << (8 * j))); }
        }
    }
  }
}

// Persistent kernel: a CTA walks work units u = blockIdx.x, blockIdx.x + gridDim.x, ...; unit = (image, band, column
// chunk of 1024 pixels).  The source rows of unit i+1 are copied global -> shared (cp.async) while unit i is processed;
// the confusion histogram lives in shared memory for the whole kernel and is flushed once per CTA.
// KT > 0: compile-time class count (loops unrolled); KT == 0: any K <= 31.
template <int KT>
__global__ void __launch_bounds__(PR_THREADS, 3) upsample_prune_kernel(
    const float* __restrict__ logits_lr, int K_rt, int h, int w, int H, int W, float sy, float sx,
    const uint8_t* __restrict__ label, int ignore_label, uint8_t* __restrict__ pred,
    unsigned long long* __restrict__ cm, int ncols_alloc, int n_chunks, int n_units) {
  const int K = KT > 0 ? KT : K_rt;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // layout: raw [2 buffers][2 rows][K][ncols_alloc] fp32 | mask [ncols_alloc] u32 | lists [PR_BUCKETS][2*PR_THREADS] u16 | hist [K*K] u32
  float* raw_all = reinterpret_cast<float*>(smem_raw);
  const int raw_stride = 2 * K * ncols_alloc;
  uint32_t* cmask = reinterpret_cast<uint32_t*>(raw_all + 2 * raw_stride);
  uint16_t* lists = reinterpret_cast<uint16_t*>(cmask + ncols_alloc);
  unsigned int* hist = reinterpret_cast<unsigned int*>(lists + PR_BUCKETS * 2 * PR_THREADS);
  __shared__ int list_n[PR_BUCKETS];
  __shared__ int band_rows[2];                                  // y_lo, y_hi of the current unit
  __shared__ float row_l0[PR_MAX_BAND], row_l1[PR_MAX_BAND];

  const int tid = threadIdx.x;
  const bool do_cm = cm != nullptr;
  const size_t hw = static_cast<size_t>(h) * w;
  if (do_cm) for (int i = tid; i < K * K; i += PR_THREADS) hist[i] = 0u;

  struct Unit { int b, band, X0, X1, c_lo, ncols; };
  auto decode = [&](int u) {
    Unit q;
    const int chunk = u % n_chunks;
    const int t = u / n_chunks;
    q.band = t % h; q.b = t / h;
    q.X0 = chunk * (PR_THREADS * 4);
    q.X1 = min(W, q.X0 + PR_THREADS * 4);
    q.c_lo = src_coord(sx, q.X0, w).i0;
    const SrcCoord c_last = src_coord(sx, q.X1 - 1, w);
    q.ncols = c_last.i0 + c_last.step - q.c_lo + 1;             // <= ncols_alloc by construction of the launch
    return q;
  };
  auto stage = [&](const Unit& q, int buf) {                    // the two source rows of every class -> shared (async)
    const int row1 = min(q.band + 1, h - 1);
    const float* p0 = logits_lr + static_cast<size_t>(q.b) * K * hw + static_cast<size_t>(q.band) * w + q.c_lo;
    const float* p1 = logits_lr + static_cast<size_t>(q.b) * K * hw + static_cast<size_t>(row1) * w + q.c_lo;
    const uint32_t d0 = static_cast<uint32_t>(__cvta_generic_to_shared(raw_all + buf * raw_stride));
    for (int k = 0; k < K; ++k, p0 += hw, p1 += hw)
      for (int c = tid; c < q.ncols; c += PR_THREADS) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d0 + 4u * ((0 * K + k) * ncols_alloc + c)), "l"(p0 + c) : "memory");
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d0 + 4u * ((1 * K + k) * ncols_alloc + c)), "l"(p1 + c) : "memory");
      }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  CmRun run; run.bin = 0; run.cnt = 0;
  int u = blockIdx.x, buf = 0;
  if (u < n_units) stage(decode(u), 0);
  for (; u < n_units; u += gridDim.x, buf ^= 1) {
    const Unit q = decode(u);
    const float* raw = raw_all + buf * raw_stride;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                                            // this unit's rows have landed; the previous unit is done
    if (u + gridDim.x < n_units) stage(decode(u + gridDim.x), buf ^ 1);
    if (tid >= PR_THREADS - 32) {       // the last warp (it rarely owns cells): the band's output rows and their weights
      const int ln = tid - (PR_THREADS - 32);
      int yl = 0, yh = 0;
      if (ln == 0) { yl = first_dst(sy, q.band, h, H); yh = first_dst(sy, q.band + 1, h, H); }
      yl = __shfl_sync(0xffffffffu, yl, 0); yh = __shfl_sync(0xffffffffu, yh, 0);
      if (ln == 0) { band_rows[0] = yl; band_rows[1] = yh; }
      if (ln < yh - yl && ln < PR_MAX_BAND) {
        const SrcCoord cy = src_coord(sy, yl + ln, h);
        row_l0[ln] = cy.l0; row_l1[ln] = cy.l1;
      }
      if (ln < PR_BUCKETS) list_n[ln] = 0;
    }
    // ---- survivor mask per cell (cell c = source columns c, min(c + 1, ncols - 1))
    for (int c = tid; c < q.ncols; c += PR_THREADS) {
      const float* q0 = raw + c;
      const float* q1 = raw + min(c + 1, q.ncols - 1);
      cmask[c] = survivor_mask<KT>(K, [&](int corner, int k) {
        return ((corner & 1) ? q1 : q0)[((corner >> 1) * K + k) * ncols_alloc];
      });
    }
    __syncthreads();
    const int y_lo = band_rows[0], R = band_rows[1] - band_rows[0];
    const int nchunks = (R + PR_ROWS - 1) / PR_ROWS;
    const int chunk_rows = (R + nchunks - 1) / nchunks;          // balanced: 8 rows -> 4 + 4, 9 -> 5 + 4
    if (R > 0) {
      // ---- one strip (4 output columns x the band's rows) per thread
      const int x0 = q.X0 + tid * 4;
      if (x0 < W) {                                             // W % 4 == 0
        const int ca = src_coord(sx, x0, w).i0 - q.c_lo, cb = src_coord(sx, x0 + 3, w).i0 - q.c_lo;
        uint32_t m = 0u;
        for (int c = ca; c <= cb; ++c) m |= cmask[c];
        const int nc = __popc(m & ~PR_NF);
        if (nc == 1 && !(m & PR_NF)) {
          const uint32_t pw = static_cast<uint32_t>(__ffs(m) - 1) * 0x01010101u;
          const size_t pix = (static_cast<size_t>(q.b) * H + y_lo) * W + x0;
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
          // one list per survivor-count class and one for non-finite cells: a warp only ever works on items of one
          // list, so its lanes run about the same number of candidate iterations
          const int bucket = (m & PR_NF) ? PR_BUCKETS - 1 : nc <= 4 ? nc - 2 : nc <= 6 ? 3 : 4;
          const int slot = atomicAdd(&list_n[bucket], nchunks);
          for (int qq = 0; qq < nchunks; ++qq)
            lists[bucket * 2 * PR_THREADS + slot + qq] = static_cast<uint16_t>(tid | (qq << 8));
        }
      }
    }
    __syncthreads();
    if (R <= 0) continue;

    // ---- queued strips, one (strip, row chunk) item per thread and turn; items of one bucket run together
    // the lists are cut into rounds of 32 items and the rounds dealt to the warps in turn, across list boundaries, so every
    // warp gets the same number of rounds (+-1) whatever the list sizes: nobody waits at the next barrier for a warp
    // that happened to own the long lists
    int round = 0;
    for (int bucket = 0; bucket < PR_BUCKETS; ++bucket) {
      const int n_items = list_n[bucket];
      const int n_rounds = (n_items + 31) >> 5;
      for (int rr = ((tid >> 5) - round) & (PR_THREADS / 32 - 1); rr < n_rounds; rr += PR_THREADS / 32) {
        const int it = rr * 32 + (tid & 31);
        if (it >= n_items) continue;
        const int item = lists[bucket * 2 * PR_THREADS + it];
        const int st = item & 0xff, r0 = (item >> 8) * chunk_rows;
        const int nr = min(chunk_rows, R - r0);
        const int xs0 = q.X0 + st * 4;
        int xi[4], xn[4];
        float xl0[4], xl1[4];
        uint32_t m = 0u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const SrcCoord c = src_coord(sx, xs0 + j, w);
          xi[j] = c.i0 - q.c_lo; xn[j] = xi[j] + c.step; xl0[j] = c.l0; xl1[j] = c.l1;
        }
        for (int c = xi[0]; c <= xi[3]; ++c) m |= cmask[c];
        uint32_t idx[PR_ROWS];                                  // four class indices per row, one byte each
        const float* rl0 = row_l0 + r0;
        const float* rl1 = row_l1 + r0;
        if (m & PR_NF) eval_item<true>(nr, raw, K, ncols_alloc, (1u << K) - 1u, xi, xn, xl0, xl1, rl0, rl1, idx);
        else eval_item<false>(nr, raw, K, ncols_alloc, m, xi, xn, xl0, xl1, rl0, rl1, idx);
        const size_t pix = (static_cast<size_t>(q.b) * H + y_lo + r0) * W + xs0;
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
      round += n_rounds;
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
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
  const size_t smem = static_cast<size_t>(4) * K * ncols_alloc * 4 + static_cast<size_t>(ncols_alloc) * 4 +
                      PR_BUCKETS * 2 * PR_THREADS * 2 + static_cast<size_t>(K) * K * 4;
  if (smem > 160 * 1024) return -100;
  const int n_chunks = (W / 4 + PR_THREADS - 1) / PR_THREADS;
  const long long units = static_cast<long long>(B) * h * n_chunks;
  if (units >= (1ll << 31)) return -100;
  const int per_sm = smem <= 72 * 1024 ? 3 : smem <= 110 * 1024 ? 2 : 1;
  const int grid = static_cast<int>(units < static_cast<long long>(num_sms()) * per_sm ? units : static_cast<long long>(num_sms()) * per_sm);
  auto launch = [&](auto kern) {
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
      if (e != cudaSuccess) return static_cast<int>(e);
    }
    kern<<<grid, PR_THREADS, smem, st>>>(logits_lr, K, h, w, H, W, sy, sx, label, ignore_label, pred, cm, ncols_alloc,
                                        n_chunks, static_cast<int>(units));
    return SL_LAUNCH_RESULT();
  };
  if (K == 8) return launch(upsample_prune_kernel<8>);
  if (K == 12) return launch(upsample_prune_kernel<12>);
  return launch(upsample_prune_kernel<0>);
}

}  // namespace sl
