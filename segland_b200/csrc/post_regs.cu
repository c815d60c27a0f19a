// sl_upsample_argmax for up-sampling by >= 2x with K = 8 or 12 classes (the eval hot path: eval_base.py:168-178,
// eval_ft.py:168-183): F.interpolate(bilinear, align_corners=True) -> argmax -> confusion counts, and at K = 12 also the
// soft-max outputs (confidence, probability maps; spec: this repo).
//
// The row-cached kernel (postproc.cu) keeps the horizontal lerps of the two live source rows in shared memory and
// re-reads them for every output row: at K = 8 its row loop is 258 instructions per warp and output row, 144 of them
// the class loop (2 LDS.128 + 4 packed flops + 12 compare/select per class), and it is bound by instruction issue
// (ncu: profiles/r1_ncu_post.txt).  Here the loop is turned inside out: the outer loop walks SOURCE-row intervals, the
// K x COLS horizontal lerps of the interval's two source rows live in REGISTERS for the ~1/sy output rows that use
// them, and the arg-max is a tournament (contiguous index groups, right side wins only when strictly greater = first
// maximum) instead of a K-long dependent chain.  No shared memory in the row loop, no per-row "is this row staged"
// bookkeeping.  Same expression tree as the row-cached kernel (v = fma(l1y, H1, l0y*H0), H = fma(l0x, a, l1x*b)), so
// every output is bit-identical to it (tests/test_gpu_parity.py, tests/test_gpu_post_regs.py).  Non-finite inputs
// are detected when a source row is loaded and routed to the np.argmax-compatible compare (first maximum, NaN wins).
#include "common.cuh"

namespace sl {

namespace {

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

// first maximum of v[LO..LO+N) on finite values: (value, index); ties keep the lower index
template <int LO, int N, int K>
__device__ __forceinline__ void tour(const float (&v)[K], float& best, int& idx) {
  if constexpr (N == 1) {
    best = v[LO];
    idx = LO;
  } else {
    constexpr int L = (N + 1) / 2;
    float bl, br;
    int il, ir;
    tour<LO, L, K>(v, bl, il);
    tour<LO + L, N - L, K>(v, br, ir);
    const bool p = br > bl;
    best = p ? br : bl;
    idx = p ? ir : il;
  }
}

}  // namespace

// A thread owns COLS consecutive output columns (COLS = 4: uchar4 label / pred accesses; 2 for K = 12 to stay under
// 128 registers) of one image and walks down a band of output rows.  grid (ceil(W / (COLS*THREADS)), bands, B).
// SOFT: also the soft-max outputs of the row-cached kernel's soft path (conf = max probability, probs = [B,K,H,W]), with
// its formulas -- e = ex2.approx(v*log2e - best*log2e), sums in class order, rcp.approx -- so the bits are the same.
template <int K, int COLS, int THREADS, bool SOFT>
__global__ void __launch_bounds__(THREADS, 512 / THREADS) upsample_regs_kernel(
    const float* __restrict__ logits_lr, int h, int w, int H, int W, int ivals_per_band, float sy, float sx,
    const uint8_t* __restrict__ label, int ignore_label, uint8_t* __restrict__ pred, float* __restrict__ conf,
    float* __restrict__ probs, unsigned long long* __restrict__ cm) {
  static_assert(COLS == 4 || COLS == 2, "COLS");
  constexpr int NP = COLS / 2;                                   // packed pairs per class
  __shared__ unsigned int hist[K * K];
  const bool do_cm = cm != nullptr;
  if (do_cm) {
    for (int i = threadIdx.x; i < K * K; i += THREADS) hist[i] = 0u;
    __syncthreads();
  }
  const int b = blockIdx.z;
  const int x0 = (blockIdx.x * THREADS + threadIdx.x) * COLS;
  // A band = the output rows of `ivals_per_band` consecutive source-row intervals, so a band of n intervals loads
  // exactly n + 1 source rows; first_y(r) = first output row whose upper source row is >= r.
  auto first_y = [&](int r) {
    if (r <= 0) return 0;
    int y = static_cast<int>(ceilf(static_cast<float>(r) / sy));
    y = max(0, min(H, y));
    while (y > 0 && src_coord(sy, y - 1, h).i0 >= r) --y;
    while (y < H && src_coord(sy, y, h).i0 < r) ++y;
    return y;
  };
  const int r_begin = blockIdx.y * ivals_per_band;
  const int y_begin = first_y(r_begin);
  const int y_end = (r_begin + ivals_per_band >= h - 1) ? H : first_y(r_begin + ivals_per_band);
  if (x0 < W && y_begin < y_end) {                               // W % COLS == 0: all columns of the thread or none
    // Up-sampling by >= COLS - 1 (host-checked): a thread's columns touch at most three source columns c0, c0+1, c0+2
    // (clamped to the row), so a source-row load is three loads per class instead of 2*COLS gathers with their 64-bit
    // address arithmetic (which made a row load cost as much as 3.6 output rows).
    float xl0[COLS], xl1[COLS];
    bool hiq[COLS];
    const int c0 = src_coord(sx, x0, w).i0;
#pragma unroll
    for (int j = 0; j < COLS; ++j) {
      const SrcCoord c = src_coord(sx, x0 + j, w);
      xl0[j] = c.l0; xl1[j] = c.l1;
      const int d = c.i0 - c0;
      hiq[j] = d == 1;
      if (d != 0 && d != 1) asm volatile("trap;");               // cannot happen for the scales the host lets through
    }
    const int hw = h * w;
    const float* plane = logits_lr + static_cast<size_t>(b) * K * hw;
    // Interior threads (c0 + 2 inside the row) evaluate H = fma(l0x, a, l1x*b) as a three-term form over the loaded
    // values s0, s1, s2 with per-column weights (l0, l1, 0) or (0, l0, l1): fma(w0, s0, fma(w1, s1, w2*s2)) gives the same
    // bits on finite inputs (a zero weight contributes an exact +-0) without selects, in packed pairs.  Rows with a
    // non-finite value, and the threads at the right edge, take the select form below.
    const bool interior = c0 + 2 <= w - 1;
    float2 w0[NP], w1[NP], w2[NP];
#pragma unroll
    for (int q = 0; q < NP; ++q) {
      w0[q] = make_float2(hiq[2 * q] ? 0.f : xl0[2 * q], hiq[2 * q + 1] ? 0.f : xl0[2 * q + 1]);
      w1[q] = make_float2(hiq[2 * q] ? xl0[2 * q] : xl1[2 * q], hiq[2 * q + 1] ? xl0[2 * q + 1] : xl1[2 * q + 1]);
      w2[q] = make_float2(hiq[2 * q] ? xl1[2 * q] : 0.f, hiq[2 * q + 1] ? xl1[2 * q + 1] : 0.f);
    }

    // horizontal lerps of source row r for the thread's columns; returns "the slice holds a non-finite value"
    auto load_row_select = [&](int r, float2 (&dst)[K][NP]) -> bool {
      const float* p0 = plane + r * w + c0;
      const float* p1 = plane + r * w + min(c0 + 1, w - 1);
      const float* p2 = plane + r * w + min(c0 + 2, w - 1);
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < K; ++k, p0 += hw, p1 += hw, p2 += hw) {
        const float s0 = __ldg(p0), s1 = __ldg(p1), s2 = __ldg(p2);
        float v[COLS];
#pragma unroll
        for (int j = 0; j < COLS; ++j) {
          v[j] = __fmaf_rn(xl0[j], hiq[j] ? s1 : s0, __fmul_rn(xl1[j], hiq[j] ? s2 : s1));
          acc += fabsf(v[j]);
        }
#pragma unroll
        for (int q = 0; q < NP; ++q) dst[k][q] = make_float2(v[2 * q], v[2 * q + 1]);
      }
      return !(acc < INFINITY);
    };
    auto load_row = [&](int r, float2 (&dst)[K][NP]) -> bool {
      if (interior) {
        const float* p0 = plane + r * w + c0;
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k, p0 += hw) {
          const float s0 = __ldg(p0), s1 = __ldg(p0 + 1), s2 = __ldg(p0 + 2);
          acc += fabsf(s0) + fabsf(s1) + fabsf(s2);
          const float2 d0 = make_float2(s0, s0), d1 = make_float2(s1, s1), d2 = make_float2(s2, s2);
#pragma unroll
          for (int q = 0; q < NP; ++q) dst[k][q] = pr_fma2(w0[q], d0, pr_fma2(w1[q], d1, pr_mul2(w2[q], d2)));
        }
        if (acc < INFINITY) return false;
      }
      return load_row_select(r, dst);
    };

    size_t pix = (static_cast<size_t>(b) * H + y_begin) * W + x0;
    // labels are fetched four rows ahead so the load latency hides behind several rows of arithmetic
    uint32_t lq0 = 0u, lq1 = 0u, lq2 = 0u, lq3 = 0u;
    auto load_label = [&](const uint8_t* p) -> uint32_t {
      if constexpr (COLS == 4) return *reinterpret_cast<const uint32_t*>(p);
      else return *reinterpret_cast<const uint16_t*>(p);
    };
    if (do_cm) {
      const uint8_t* lp = label + pix;
      if (y_begin + 0 < y_end) lq0 = load_label(lp);
      if (y_begin + 1 < y_end) lq1 = load_label(lp + W);
      if (y_begin + 2 < y_end) lq2 = load_label(lp + 2 * static_cast<size_t>(W));
      if (y_begin + 3 < y_end) lq3 = load_label(lp + 3 * static_cast<size_t>(W));
    }
    // Confusion counts: vertical run-length accumulation per thread over the packed (label, prediction) words of its
    // columns -- both maps are piecewise constant down a column -- with one flush (<= COLS shared atomics) per run.
    // Ignored pixels (scattered, ~1 % of the bench labels) do not break a run: a label byte equal to ignore_label is
    // taken as "same as the run" and counted in a per-byte counter that the flush subtracts.
    const uint32_t lane_mask = COLS == 4 ? 0xffffffffu : 0x0000ffffu;
    const bool has_ign = ignore_label >= 0 && ignore_label <= 255;
    const uint32_t ign4 = static_cast<uint32_t>(ignore_label & 0xff) * 0x01010101u;
    uint32_t run_l = 0u, run_p = 0u, run_n = 0u, run_ign = 0u;   // run_ign: four 8-bit counters (a band has < 255 rows)
    auto flush_run = [&]() {
      if (run_n == 0u) return;
      int bin[COLS];
      bool valid[COLS];
      bool uni = true;
#pragma unroll
      for (int j = 0; j < COLS; ++j) {
        const int lab = static_cast<int>((run_l >> (8 * j)) & 0xffu);
        valid[j] = lab != ignore_label && lab < K;
        bin[j] = lab * K + static_cast<int>((run_p >> (8 * j)) & 0xffu);
        uni = uni && valid[j] && bin[j] == bin[0];
      }
      if (uni) {
        atomicAdd(&hist[bin[0]], run_n * COLS - static_cast<uint32_t>(__dp4a(run_ign, 0x01010101u, 0u)));
      } else {
#pragma unroll
        for (int j = 0; j < COLS; ++j)
          if (valid[j]) atomicAdd(&hist[bin[j]], run_n - ((run_ign >> (8 * j)) & 0xffu));
      }
    };

    int y = y_begin;
    SrcCoord cy = src_coord(sy, y, h);

    // all output rows whose upper source row is r_top, with `top` / `bot` = the lerped rows r_top / r_top + step
    auto rows = [&](const int r_top, const float2 (&top)[K][NP], const float2 (&bot)[K][NP], const bool bad) {
      do {
        const float2 l0 = make_float2(cy.l0, cy.l0), l1 = make_float2(cy.l1, cy.l1);
        const uint32_t l4 = lq0;
        if (do_cm) {
          lq0 = lq1; lq1 = lq2; lq2 = lq3;
          if (y + 4 < y_end) lq3 = load_label(label + pix + 4 * static_cast<size_t>(W));
        }
        int idx[COLS];
        // soft-max of one packed pair of pixels from its K values and their maximum (row-cached kernel's formulas)
        auto soft_pair = [&](float (&va)[K], float (&vb)[K], float best_a, float best_b, int q) {
          constexpr float kLog2e = 1.4426950408889634f;
          const float nba = -best_a * kLog2e, nbb = -best_b * kLog2e;
          float sa = 0.f, sb = 0.f;
#pragma unroll
          for (int k = 0; k < K; ++k) {
            float ea, eb;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ea) : "f"(fmaf(va[k], kLog2e, nba)));
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(eb) : "f"(fmaf(vb[k], kLog2e, nbb)));
            va[k] = ea; vb[k] = eb;
            sa += ea; sb += eb;
          }
          float ia, ib;
          asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ia) : "f"(sa));
          asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ib) : "f"(sb));
          if (conf) *reinterpret_cast<float2*>(conf + pix + 2 * q) = make_float2(ia, ib);
          if (probs) {
            float* dst = probs + (static_cast<size_t>(b) * K) * (static_cast<size_t>(H) * W) + (pix - static_cast<size_t>(b) * H * W) + 2 * q;
            const float2 inv = make_float2(ia, ib);
#pragma unroll
            for (int k = 0; k < K; ++k) {
              __stcs(reinterpret_cast<float2*>(dst), pr_mul2(make_float2(va[k], vb[k]), inv));   // streaming: written once
              dst += static_cast<size_t>(H) * W;
            }
          }
        };
        if (!bad) {
#pragma unroll
          for (int q = 0; q < NP; ++q) {
            float va[K], vb[K];
#pragma unroll
            for (int k = 0; k < K; ++k) {
              const float2 t = pr_fma2(l1, bot[k][q], pr_mul2(l0, top[k][q]));
              va[k] = t.x;
              vb[k] = t.y;
            }
            float best_a, best_b;
            tour<0, K, K>(va, best_a, idx[2 * q]);
            tour<0, K, K>(vb, best_b, idx[2 * q + 1]);
            if constexpr (SOFT) soft_pair(va, vb, best_a, best_b, q);
          }
        } else {
#pragma unroll
          for (int q = 0; q < NP; ++q) {
            float va[K], vb[K];
            float best_a = -INFINITY, best_b = -INFINITY;
            idx[2 * q] = 0; idx[2 * q + 1] = 0;
#pragma unroll
            for (int k = 0; k < K; ++k) {
              va[k] = __fmaf_rn(cy.l1, bot[k][q].x, __fmul_rn(cy.l0, top[k][q].x));
              vb[k] = __fmaf_rn(cy.l1, bot[k][q].y, __fmul_rn(cy.l0, top[k][q].y));
              if (va[k] > best_a || (va[k] != va[k] && best_a == best_a)) { best_a = va[k]; idx[2 * q] = k; }
              if (vb[k] > best_b || (vb[k] != vb[k] && best_b == best_b)) { best_b = vb[k]; idx[2 * q + 1] = k; }
            }
            if constexpr (SOFT) soft_pair(va, vb, best_a, best_b, q);
          }
        }
        uint32_t p4 = static_cast<uint32_t>(idx[0]) | (static_cast<uint32_t>(idx[1]) << 8);
        if constexpr (COLS == 4) p4 |= (static_cast<uint32_t>(idx[2]) << 16) | (static_cast<uint32_t>(idx[3]) << 24);
        if (pred) {
          if constexpr (COLS == 4) *reinterpret_cast<uint32_t*>(pred + pix) = p4;
          else *reinterpret_cast<uint16_t*>(pred + pix) = static_cast<uint16_t>(p4);
        }
        if (do_cm) {
          // run-length over the packed (label, prediction) words of the thread's columns: both maps are piecewise
          // constant down a column, so the common row costs two compares and an add
          const uint32_t x = l4 ^ ign4;
          uint32_t t = ~(((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x | 0x7f7f7f7fu) & lane_mask;   // 0x80 in every byte == ignore_label
          if (!has_ign) t = 0u;
          const uint32_t ones = t >> 7, m = ones * 0xffu;
          if ((((l4 & ~m) | (run_l & m)) == run_l) && p4 == run_p) {
            ++run_n;
            run_ign += ones;
          } else {
            flush_run();
            run_l = l4; run_p = p4; run_n = 1u; run_ign = ones;
          }
        }
        ++y;
        pix += W;
        if (y >= y_end) return;
        cy = src_coord(sy, y, h);
      } while (cy.i0 == r_top);
    };

    float2 A[K][NP], Bq[K][NP];
    int rA = cy.i0, rB = -1;
    bool badA = load_row(rA, A), badB = false;
    while (y < y_end) {
      // upper row in A
      {
        const int r1 = cy.i0 + cy.step;
        if (r1 == rA) {
#pragma unroll
          for (int k = 0; k < K; ++k)
#pragma unroll
            for (int q = 0; q < NP; ++q) Bq[k][q] = A[k][q];
          badB = badA;
        } else if (rB != r1) {
          badB = load_row(r1, Bq);
        }
        rB = r1;
      }
      rows(rA, A, Bq, badA | badB);
      if (y >= y_end) break;
      if (cy.i0 != rB) {                                         // skipped a source row (not when up-sampling): restart
        rA = cy.i0;
        badA = load_row(rA, A);
        rB = -1;
        continue;
      }
      // upper row in Bq
      {
        const int r1 = cy.i0 + cy.step;
        if (r1 == rB) {
#pragma unroll
          for (int k = 0; k < K; ++k)
#pragma unroll
            for (int q = 0; q < NP; ++q) A[k][q] = Bq[k][q];
          badA = badB;
        } else {
          badA = load_row(r1, A);
        }
        rA = r1;
      }
      rows(rB, Bq, A, badA | badB);
      if (y >= y_end) break;
      if (cy.i0 != rA) {
        rA = cy.i0;
        badA = load_row(rA, A);
      }
      rB = -1;
    }
    if (do_cm) flush_run();
  }
  if (do_cm) {
    __syncthreads();
    for (int i = threadIdx.x; i < K * K; i += THREADS)
      if (hist[i]) atomicAdd(&cm[i], static_cast<unsigned long long>(hist[i]));
  }
}

// Returns -100 when the shape is outside this kernel's range (the caller falls back to the row-cached kernel).
int launch_upsample_regs(const float* logits_lr, int B, int K, int h, int w, int H, int W, float sy, float sx,
                         const uint8_t* label, int ignore_label, uint8_t* pred, float* conf, float* probs,
                         unsigned long long* cm, cudaStream_t st) {
  if ((conf != nullptr || probs != nullptr) && K != 12) return -100;   // K = 8 soft outputs: four columns per thread spill; row-cached kernel
  if (!(K == 8 || K == 12) || W % 4 != 0 || B > 65535 ||
      !(sy > 0.f && sy <= 0.5f && sx > 0.f && sx <= (K == 8 ? 0.3f : 0.5f)) || h < 2) return -100;
  // 128-thread CTAs: 16 K registers and 256 B of shared memory, so that one CTA fits on an SM NEXT TO a CTA of the
  // background-MLP pair kernel (12 warp slots x 128 registers + 225 KB) and this issue-bound kernel can run on a second
  // stream in the issue slots the tensor-bound kernel leaves idle (sweep.PipelinedTileEvaluator).  Alone, 4 such CTAs
  // per SM perform like 2 of 256 threads.
  constexpr int threads = 128;
  const int cols = K == 8 ? 4 : 2;
  const int gx = (W / cols + threads - 1) / threads;
  // band = n source-row intervals (n + 1 source rows loaded, each ~1.3 output rows of work), n by wave quantisation;
  // the per-byte ignore counters of the confusion run lengths need < 255 rows per band
  const long long slots = static_cast<long long>(512 / threads) * num_sms();
  const int intervals = h - 1;
  int best_n = 0;
  double best_cost = 1e30;
  for (int n : {1, 2, 3, 4, 6, 8}) {
    const int bands = (intervals + n - 1) / n;
    if (bands > 65535 || n / static_cast<double>(sy) + 2.0 >= 250.0) continue;
    const long long ctas = static_cast<long long>(gx) * bands * B;
    const long long waves = (ctas + slots - 1) / slots;
    const double cost = static_cast<double>(waves) * (n / static_cast<double>(sy) + 1.3 * (n + 1));
    if (cost < best_cost) { best_cost = cost; best_n = n; }
  }
  if (best_n == 0) return -100;
  const dim3 grid(gx, (intervals + best_n - 1) / best_n, B);
#define SL_REGS_LAUNCH(KK, CC, TT, SOFT) upsample_regs_kernel<KK, CC, TT, SOFT><<<grid, TT, 0, st>>>( \
      logits_lr, h, w, H, W, best_n, sy, sx, label, ignore_label, pred, conf, probs, cm)
  const bool soft = conf != nullptr || probs != nullptr;
  if (K == 8) SL_REGS_LAUNCH(8, 4, threads, false);            // (soft outputs at K = 8 were turned away above)
  else { if (soft) SL_REGS_LAUNCH(12, 2, threads, true); else SL_REGS_LAUNCH(12, 2, threads, false); }
#undef SL_REGS_LAUNCH
  return SL_LAUNCH_RESULT();
}

}  // namespace sl
