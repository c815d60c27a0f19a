// Dense post-processing behind the POP head:
//   sl_upsample_argmax  F.interpolate(bilinear, align_corners=True) -> argmax -> confusion matrix
//                       (eval_base.py:168-178, eval_ft.py:168-183, ft_pop.py:327-331)
//   sl_pseudo_label     pspnet_pop.py:221-231
//   sl_confusion        utils/pyt_utils.py:182-200 (get_confusion_matrix)
//   sl_inter_union      utils/pyt_utils.py:293-305 (intersectionAndUnionGPU)
//   sl_views_reduce     flip/multi-view aggregation at feature resolution (spec: this repo)
// The full-resolution logits are never written unless the caller asks for them (probs / logits_hr).
#include <cstdlib>
#include "common.cuh"

namespace sl {

// Interpolated value of channel plane `src` [h][w] at output pixel given by (cy, cx), in ATen's
// association: l0y*(l0x*a + l1x*b) + l1y*(l0x*c + l1x*d).
__device__ __forceinline__ float bilerp(const float* __restrict__ src, int w, const SrcCoord& cy, const SrcCoord& cx) {
  const float* r0 = src + cy.i0 * w + cx.i0;
  const float* r1 = r0 + cy.step * w;
  const float a = __ldg(r0), b = __ldg(r0 + cx.step), c = __ldg(r1), d = __ldg(r1 + cx.step);
  return cy.l0 * (cx.l0 * a + cx.l1 * b) + cy.l1 * (cx.l0 * c + cx.l1 * d);
}

// np.argmax / torch.argmax semantics: first maximum; NaN counts as maximal (first NaN wins).
__device__ __forceinline__ void argmax_step(float v, int k, float& best, int& idx) {
  if (v > best || (v != v && best == best)) { best = v; idx = k; }
}

constexpr int UP_PX = 4;  // consecutive x pixels per thread (uchar4 label/pred accesses)

// KP: compile-time bound on K so the per-pixel channel values stay in registers when the
// softmax outputs need them (KP == 0: values are not kept).
template <int KP>
__global__ void __launch_bounds__(256) upsample_argmax_kernel(
    const float* __restrict__ logits_lr, int B, int K, int h, int w, int H, int W, float sy, float sx,
    const uint8_t* __restrict__ label, int ignore_label, uint8_t* __restrict__ pred, float* __restrict__ conf,
    float* __restrict__ probs, float* __restrict__ logits_hr, unsigned long long* __restrict__ cm) {
  __shared__ unsigned int hist[SL_MAX_CLASSES * SL_MAX_CLASSES];
  const bool do_cm = cm != nullptr;
  if (do_cm) {
    for (int i = threadIdx.x; i < K * K; i += 256) hist[i] = 0u;
    __syncthreads();
  }
  const int groups_x = (W + UP_PX - 1) / UP_PX;
  const long long total = static_cast<long long>(B) * H * groups_x;
  const bool vec_ok = (W % UP_PX) == 0;
  const long long start = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  const long long stride = static_cast<long long>(gridDim.x) * 256;
  // every lane runs the same number of iterations so the warp-collective histogram stays converged
  const long long iters = (total + stride - 1) / stride;
  for (long long it = 0; it < iters; ++it) {
    const long long g = start + it * stride;
    const bool active = g < total;
    int idx[UP_PX];
    int lab[UP_PX];
#pragma unroll
    for (int j = 0; j < UP_PX; ++j) { idx[j] = 0; lab[j] = -1; }
    if (active) {
      const int gx = static_cast<int>(g % groups_x);
      const long long rowid = g / groups_x;
      const int y = static_cast<int>(rowid % H);
      const int b = static_cast<int>(rowid / H);
      const int x0 = gx * UP_PX;
      const SrcCoord cy = src_coord(sy, y, h);
      SrcCoord cx[UP_PX];
#pragma unroll
      for (int j = 0; j < UP_PX; ++j) cx[j] = src_coord(sx, min(x0 + j, W - 1), w);
      float best[UP_PX];
#pragma unroll
      for (int j = 0; j < UP_PX; ++j) best[j] = -INFINITY;
      float vals[UP_PX][KP > 0 ? KP : 1];
      const float* plane = logits_lr + static_cast<size_t>(b) * K * h * w;
      const size_t pix = (static_cast<size_t>(b) * H + y) * W + x0;           // index into [B,H,W]
      const size_t HW = static_cast<size_t>(H) * W;
      auto channel = [&](int k, float (&v)[UP_PX]) {
#pragma unroll
        for (int j = 0; j < UP_PX; ++j) {
          v[j] = bilerp(plane + static_cast<size_t>(k) * h * w, w, cy, cx[j]);
          argmax_step(v[j], k, best[j], idx[j]);
        }
        if (logits_hr) {
          float* dst = logits_hr + (static_cast<size_t>(b) * K + k) * HW + static_cast<size_t>(y) * W + x0;
          if (vec_ok) *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
          else
#pragma unroll
            for (int j = 0; j < UP_PX; ++j) if (x0 + j < W) dst[j] = v[j];
        }
      };
      if constexpr (KP > 0) {
#pragma unroll
        for (int k = 0; k < KP; ++k) {
          if (k < K) {
            float v[UP_PX];
            channel(k, v);
#pragma unroll
            for (int j = 0; j < UP_PX; ++j) vals[j][k] = v[j];
          }
        }
      } else {
#pragma unroll 2
        for (int k = 0; k < K; ++k) {
          float v[UP_PX];
          channel(k, v);
        }
      }
      if (KP > 0 && (conf || probs)) {
        float inv[UP_PX];
#pragma unroll
        for (int j = 0; j < UP_PX; ++j) {
          float s = 0.f;
#pragma unroll
          for (int k = 0; k < KP; ++k) if (k < K) { vals[j][k] = __expf(vals[j][k] - best[j]); s += vals[j][k]; }
          inv[j] = 1.f / s;
        }
        if (conf) {
          float* dst = conf + pix;
          if (vec_ok) *reinterpret_cast<float4*>(dst) = make_float4(inv[0], inv[1], inv[2], inv[3]);
          else
#pragma unroll
            for (int j = 0; j < UP_PX; ++j) if (x0 + j < W) dst[j] = inv[j];
        }
        if (probs) {
#pragma unroll
          for (int k = 0; k < KP; ++k) if (k < K) {
            float* dst = probs + (static_cast<size_t>(b) * K + k) * HW + static_cast<size_t>(y) * W + x0;
            if (vec_ok)
              *reinterpret_cast<float4*>(dst) = make_float4(vals[0][k] * inv[0], vals[1][k] * inv[1],
                                                            vals[2][k] * inv[2], vals[3][k] * inv[3]);
            else
#pragma unroll
              for (int j = 0; j < UP_PX; ++j) if (x0 + j < W) dst[j] = vals[j][k] * inv[j];
          }
        }
      }
      if (pred) {
        if (vec_ok) *reinterpret_cast<uchar4*>(pred + pix) = make_uchar4(idx[0], idx[1], idx[2], idx[3]);
        else
#pragma unroll
          for (int j = 0; j < UP_PX; ++j) if (x0 + j < W) pred[pix + j] = static_cast<uint8_t>(idx[j]);
      }
      if (do_cm) {
        if (vec_ok) {
          const uchar4 l4 = *reinterpret_cast<const uchar4*>(label + pix);
          lab[0] = l4.x; lab[1] = l4.y; lab[2] = l4.z; lab[3] = l4.w;
        } else {
#pragma unroll
          for (int j = 0; j < UP_PX; ++j) if (x0 + j < W) lab[j] = label[pix + j];
        }
      }
    }
    if (do_cm) {
#pragma unroll
      for (int j = 0; j < UP_PX; ++j) {
        const bool valid = lab[j] >= 0 && lab[j] != ignore_label && lab[j] < K;
        hist_add_warp(hist, valid ? lab[j] * K + idx[j] : 0, valid);
      }
    }
  }
  if (do_cm) {
    __syncthreads();
    for (int i = threadIdx.x; i < K * K; i += 256)
      if (hist[i]) atomicAdd(&cm[i], static_cast<unsigned long long>(hist[i]));
  }
}

// ------------------------------------------------------------------ fast path (W % 4 == 0)
// A thread owns 4 consecutive output columns of one image and walks down a band of output rows.
// ATen's association is horizontal-first -- val = l0y*(l0x*a + l1x*b) + l1y*(l0x*c + l1x*d) -- so
// the horizontal lerps Hrow[k][x] = l0x*in[k][r][x0] + l1x*in[k][r][x1] of a source row r serve every
// output row that touches r (8 rows at the x8 scale of the stride-8 models).  Each thread caches
// its own 4 columns of the two live source rows in shared memory (slot = r & 1, float4 per
// channel: private to the thread, so no block barrier); an output pixel-channel then costs two
// shared loads, a multiply, an FMA and a 3-instruction argmax step instead of four global loads
// and six flops.  Bit-identical to the generic kernel (same expression tree).  Non-finite inputs
// are detected when a source row is staged and routed to the NaN-aware (np.argmax) compare.

// Warp-aggregated histogram update for (bin, count) pairs: lanes holding the same bin are grouped
// with match.any, their counts summed with redux, and one shared-memory atomic is issued per
// distinct bin in the warp.
__device__ __forceinline__ void hist_add_grouped(unsigned int* hist, int bin, unsigned int cnt) {
  const unsigned peers = __match_any_sync(0xffffffffu, bin);
  const unsigned total = __reduce_add_sync(peers, cnt);
  if ((threadIdx.x & 31) == (__ffs(peers) - 1) && total) atomicAdd(&hist[bin], total);
}

__device__ __forceinline__ float2 mul2(const float2 a, const float2 b) {
  float2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<const unsigned long long&>(a)), "l"(reinterpret_cast<const unsigned long long&>(b)));
  return d;
}
__device__ __forceinline__ float2 fma2(const float2 a, const float2 b, const float2 c) {
  float2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<const unsigned long long&>(a)), "l"(reinterpret_cast<const unsigned long long&>(b)),
        "l"(reinterpret_cast<const unsigned long long&>(c)));
  return d;
}

// KP == 0: predictions / confusion only (the eval hot path).  KP > 0: also soft outputs (conf, probs,
// up-sampled logits) for K <= KP classes, whose interpolated values stay in registers so the softmax
// costs one interpolation and one exp per value.
template <int THREADS, int KP, bool EXACT = false, bool LG = true>      // LG: the up-sampled logits may be requested
__global__ void __launch_bounds__(THREADS, KP > 0 ? (THREADS == 256 ? 2 : 4) : (THREADS == 256 ? 3 : 8)) upsample_rows_kernel(
    const float* __restrict__ logits_lr, int K, int h, int w, int H, int W, int rows_per_band, float sy, float sx,
    const uint8_t* __restrict__ label, int ignore_label, uint8_t* __restrict__ pred, float* __restrict__ conf,
    float* __restrict__ probs, float* __restrict__ logits_hr, unsigned long long* __restrict__ cm, int coop) {
  constexpr bool EXTRA = KP > 0;
  if constexpr (KP < 0) K = -KP;                               // exact-K instantiation of the plain path: loops fully unrolled
  if constexpr (KP > 0 && EXACT) K = KP;                       // exact-K soft-output path: the k < K predicates fold away
  extern __shared__ __align__(16) float hrow[];                 // [2][K][THREADS] float4, then raw [K][ncols] (coop)
  __shared__ unsigned int hist[SL_MAX_CLASSES * SL_MAX_CLASSES];
  const bool do_cm = cm != nullptr;
  if (do_cm) {
    for (int i = threadIdx.x; i < K * K; i += THREADS) hist[i] = 0u;
    __syncthreads();
  }
  const int b = blockIdx.z;
  const int x0 = (blockIdx.x * THREADS + threadIdx.x) * 4;
  const bool col_ok = x0 < W;                                   // W % 4 == 0: all four columns or none
  const int y_begin = blockIdx.y * rows_per_band;
  const int y_end = min(H, y_begin + rows_per_band);
  float4* my = reinterpret_cast<float4*>(hrow) + threadIdx.x;   // element [slot][k] at my[(slot*K + k) * THREADS]
  int xi[4], xs[4];
  float xl0[4], xl1[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const SrcCoord c = src_coord(sx, min(x0 + j, W - 1), w);
    xi[j] = c.i0; xs[j] = c.i0 + c.step; xl0[j] = c.l0; xl1[j] = c.l1;
  }
  const int hw = h * w;
  const float* plane = logits_lr + static_cast<size_t>(b) * K * hw;
  const size_t HW = static_cast<size_t>(H) * W;
  int row_even = -1, row_odd = -1;                              // source rows held by slot 0 / slot 1
  bool bad_even = false, bad_odd = false;                       // slot holds a non-finite value

  // Cooperative staging of a source row (coop, up-sampling by >= 2x): the CTA's columns of row r, all K channels,
  // are copied global -> shared with cp.async by all threads (coalesced, no registers), one row AHEAD of its first
  // use, so the ~700-cycle L2 latency of the old per-thread gathers (35 % of the kernel's stall samples) is off the
  // critical path; the per-thread horizontal lerps then read shared memory.  Every thread of the CTA walks the same
  // rows, so the two barriers per new source row (one per ~1/sy output rows) are uniform.
  const int cta_x0 = blockIdx.x * THREADS * 4;
  const int c_lo = src_coord(sx, min(cta_x0, W - 1), w).i0;
  const SrcCoord c_last = src_coord(sx, min(cta_x0 + THREADS * 4 - 1, W - 1), w);
  const int ncols = c_last.i0 + c_last.step - c_lo + 1;
  float* raw = hrow + static_cast<size_t>(2) * K * THREADS * 4;
  int pf_row = -1;                                              // source row in flight / staged in `raw`
  auto stage_row = [&](int r) {
    const float* src = plane + r * w + c_lo;
    const uint32_t dst0 = static_cast<uint32_t>(__cvta_generic_to_shared(raw));
    for (int k = 0; k < K; ++k, src += hw)
      for (int c = threadIdx.x; c < ncols; c += THREADS)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst0 + 4u * (k * ncols + c)), "l"(src + c) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
    pf_row = r;
  };

  auto fill = [&](int r) {                                      // horizontal lerps of source row r
    const int slot = r & 1;
    if ((slot ? row_odd : row_even) == r) return;
    float4* dst = my + slot * K * THREADS;
    bool bad = false;
    if (coop) {
      if (pf_row != r) stage_row(r);
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();
      const float* rr = raw - c_lo;
#pragma unroll 2
      for (int k = 0; k < K; ++k, rr += ncols, dst += THREADS) {
        float4 v;
        v.x = xl0[0] * rr[xi[0]] + xl1[0] * rr[xs[0]];
        v.y = xl0[1] * rr[xi[1]] + xl1[1] * rr[xs[1]];
        v.z = xl0[2] * rr[xi[2]] + xl1[2] * rr[xs[2]];
        v.w = xl0[3] * rr[xi[3]] + xl1[3] * rr[xs[3]];
        bad |= !(fabsf(v.x) + fabsf(v.y) + fabsf(v.z) + fabsf(v.w) < INFINITY);
        *dst = v;
      }
      __syncthreads();                                          // everyone is done with `raw`
      if (r + 1 < h) stage_row(r + 1);                          // the next new row, one source row ahead
    } else {
      const float* rk = plane + r * w;
#pragma unroll 2
      for (int k = 0; k < K; ++k, rk += hw, dst += THREADS) {
        float4 v;
        v.x = xl0[0] * __ldg(rk + xi[0]) + xl1[0] * __ldg(rk + xs[0]);
        v.y = xl0[1] * __ldg(rk + xi[1]) + xl1[1] * __ldg(rk + xs[1]);
        v.z = xl0[2] * __ldg(rk + xi[2]) + xl1[2] * __ldg(rk + xs[2]);
        v.w = xl0[3] * __ldg(rk + xi[3]) + xl1[3] * __ldg(rk + xs[3]);
        bad |= !(fabsf(v.x) + fabsf(v.y) + fabsf(v.z) + fabsf(v.w) < INFINITY);
        *dst = v;
      }
    }
    if (slot) { row_odd = r; bad_odd = bad; } else { row_even = r; bad_even = bad; }
  };

  size_t pix = (static_cast<size_t>(b) * H + y_begin) * W + x0;
  // labels are fetched four rows ahead so the load latency hides behind several rows of arithmetic
  uint32_t lq0 = 0u, lq1 = 0u, lq2 = 0u, lq3 = 0u;              // four packed labels per row, kept as words (no byte shuffles)
  if (do_cm && col_ok) {
    const uint8_t* lp = label + pix;
    if (y_begin + 0 < y_end) lq0 = *reinterpret_cast<const uint32_t*>(lp);
    if (y_begin + 1 < y_end) lq1 = *reinterpret_cast<const uint32_t*>(lp + W);
    if (y_begin + 2 < y_end) lq2 = *reinterpret_cast<const uint32_t*>(lp + 2 * static_cast<size_t>(W));
    if (y_begin + 3 < y_end) lq3 = *reinterpret_cast<const uint32_t*>(lp + 3 * static_cast<size_t>(W));
  }
  // confusion counts: vertical run-length accumulation per thread (label and prediction maps are piecewise
  // constant, so a thread's four pixels usually stay in one bin for many rows): no warp collectives, one
  // shared-memory atomic per run.
  int run_bin = 0;
  unsigned int run_cnt = 0;
  for (int y = y_begin; y < y_end; ++y, pix += W) {
    int idx[4] = {0, 0, 0, 0};
    const uint32_t l4 = lq0;
    const SrcCoord cy = src_coord(sy, y, h);
    const int r1 = cy.i0 + cy.step;
    fill(cy.i0);                                                // uniform across the CTA (barriers inside when coop)
    fill(r1);
    if (col_ok) {
      if (do_cm) {
        lq0 = lq1; lq1 = lq2; lq2 = lq3;
        if (y + 4 < y_end) lq3 = *reinterpret_cast<const uint32_t*>(label + pix + 4 * static_cast<size_t>(W));
      }
      const float4* h0 = my + ((cy.i0 & 1) * K) * THREADS;
      const float4* h1 = my + ((r1 & 1) * K) * THREADS;
      const float l0 = cy.l0, l1 = cy.l1;
      float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      const bool bad = ((cy.i0 & 1) ? bad_odd : bad_even) | ((r1 & 1) ? bad_odd : bad_even);
      if constexpr (!EXTRA) {
        if (!bad) {
#pragma unroll (KP < 0 ? -KP : 4)
          for (int k = 0; k < K; ++k) {
            const float4 a = h0[k * THREADS], c = h1[k * THREADS];
            // l0*a + l1*c on packed fp32 pairs (FMUL2 + FFMA2): same products and sums as the scalar form
            float2 t01 = mul2(make_float2(l0, l0), make_float2(a.x, a.y)), t23 = mul2(make_float2(l0, l0), make_float2(a.z, a.w));
            t01 = fma2(make_float2(l1, l1), make_float2(c.x, c.y), t01);
            t23 = fma2(make_float2(l1, l1), make_float2(c.z, c.w), t23);
            const float v[4] = {t01.x, t01.y, t23.x, t23.y};
            if (KP < 0 && k == 0) {                              // finite values: v > -inf or v == -inf, either way (v, 0)
#pragma unroll
              for (int j = 0; j < 4; ++j) best[j] = v[j];
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (v[j] > best[j]) { best[j] = v[j]; idx[j] = k; }
            }
          }
        } else {
          for (int k = 0; k < K; ++k) {
            const float4 a = h0[k * THREADS], c = h1[k * THREADS];
            const float v[4] = {l0 * a.x + l1 * c.x, l0 * a.y + l1 * c.y, l0 * a.z + l1 * c.z, l0 * a.w + l1 * c.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) argmax_step(v[j], k, best[j], idx[j]);
          }
        }
      } else {
        float vals[4][KP > 0 ? KP : 1];
        const size_t plane_off = pix - static_cast<size_t>(b) * HW;
        // output pointers advance by one class plane per k (two adds) instead of a 64-bit multiply-add per store
        float* lg_out = (LG && logits_hr) ? logits_hr + static_cast<size_t>(b) * K * HW + plane_off : nullptr;
        if (!bad) {
#pragma unroll
          for (int k = 0; k < KP; ++k) {
            if (k < K) {
              const float4 a = h0[k * THREADS], c = h1[k * THREADS];
              // l0*a + l1*c on packed fp32 pairs (FMUL2 + FFMA2): same products and sums as the scalar form
              float2 t01 = mul2(make_float2(l0, l0), make_float2(a.x, a.y)), t23 = mul2(make_float2(l0, l0), make_float2(a.z, a.w));
              t01 = fma2(make_float2(l1, l1), make_float2(c.x, c.y), t01);
              t23 = fma2(make_float2(l1, l1), make_float2(c.z, c.w), t23);
              const float v[4] = {t01.x, t01.y, t23.x, t23.y};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                if (k == 0) { best[j] = v[j]; } else if (v[j] > best[j]) { best[j] = v[j]; idx[j] = k; }
                vals[j][k] = v[j];
              }
              if (LG && lg_out) { __stcs(reinterpret_cast<float4*>(lg_out), make_float4(v[0], v[1], v[2], v[3])); lg_out += HW; }
            }
          }
        } else {
#pragma unroll
          for (int k = 0; k < KP; ++k) {
            if (k < K) {
              const float4 a = h0[k * THREADS], c = h1[k * THREADS];
              const float v[4] = {l0 * a.x + l1 * c.x, l0 * a.y + l1 * c.y, l0 * a.z + l1 * c.z, l0 * a.w + l1 * c.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) { argmax_step(v[j], k, best[j], idx[j]); vals[j][k] = v[j]; }
              if (LG && lg_out) { __stcs(reinterpret_cast<float4*>(lg_out), make_float4(v[0], v[1], v[2], v[3])); lg_out += HW; }
            }
          }
        }
        if (conf || probs) {
          // softmax with exp2: e = 2^(v*log2e - best*log2e) (one FFMA + MUFU.EX2 per value), 1/sum by MUFU.RCP
          constexpr float kLog2e = 1.4426950408889634f;
          float s[4] = {0.f, 0.f, 0.f, 0.f};
          const float nb[4] = {-best[0] * kLog2e, -best[1] * kLog2e, -best[2] * kLog2e, -best[3] * kLog2e};
#pragma unroll
          for (int k = 0; k < KP; ++k)
            if (k < K) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float e;
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(vals[j][k], kLog2e, nb[j])));
                vals[j][k] = e;
                s[j] += e;
              }
            }
          float inv[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv[j]) : "f"(s[j]));
          if (conf) *reinterpret_cast<float4*>(conf + pix) = make_float4(inv[0], inv[1], inv[2], inv[3]);
          if (probs) {
            float* pr_out = probs + static_cast<size_t>(b) * K * HW + plane_off;
#pragma unroll
            for (int k = 0; k < KP; ++k)
              if (k < K) {
                const float2 p01 = mul2(make_float2(vals[0][k], vals[1][k]), make_float2(inv[0], inv[1]));
                const float2 p23 = mul2(make_float2(vals[2][k], vals[3][k]), make_float2(inv[2], inv[3]));
                __stcs(reinterpret_cast<float4*>(pr_out), make_float4(p01.x, p01.y, p23.x, p23.y));   // streaming: written once
                pr_out += HW;
              }
          }
        }
      }
      if (pred) *reinterpret_cast<uchar4*>(pred + pix) = make_uchar4(idx[0], idx[1], idx[2], idx[3]);
    }
    if (do_cm && col_ok) {
      const int lab[4] = {static_cast<int>(l4 & 0xffu), static_cast<int>((l4 >> 8) & 0xffu), static_cast<int>((l4 >> 16) & 0xffu),
                          static_cast<int>(l4 >> 24)};
      int bin[4];
      bool valid[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        valid[j] = lab[j] != ignore_label && lab[j] < K;
        bin[j] = lab[j] * K + idx[j];
      }
      const bool uni = valid[0] && valid[1] && valid[2] && valid[3] && bin[0] == bin[1] && bin[1] == bin[2] &&
                       bin[2] == bin[3];
      if (uni && bin[0] == run_bin) {
        run_cnt += 4u;                                          // the common case: 2 instructions
      } else if (uni) {
        if (run_cnt) atomicAdd(&hist[run_bin], run_cnt);
        run_bin = bin[0];
        run_cnt = 4u;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (valid[j]) {
            if (bin[j] == run_bin) ++run_cnt; else atomicAdd(&hist[bin[j]], 1u);
          }
      }
    }
  }
  if (do_cm && run_cnt) atomicAdd(&hist[run_bin], run_cnt);
  asm volatile("cp.async.wait_group 0;" ::: "memory");       // the last row's look-ahead copy may still be in flight
  if (do_cm) {
    __syncthreads();
    for (int i = threadIdx.x; i < K * K; i += THREADS)
      if (hist[i]) atomicAdd(&cm[i], static_cast<unsigned long long>(hist[i]));
  }
}

template <int THREADS>
static int launch_rows(const float* logits_lr, int B, int K, int h, int w, int H, int W, int rows_per_band, float sy,
                       float sx, const uint8_t* label, int ignore_label, uint8_t* pred, float* conf, float* probs,
                       float* logits_hr, unsigned long long* cm, cudaStream_t st) {
  const bool extra = conf || probs || logits_hr;
  // cooperative row staging for up-sampling by >= 2x: the CTA's source columns, all K channels, one row
  const int ncols_max = static_cast<int>(THREADS * 4 * sx) + 3;
  const size_t base_smem = static_cast<size_t>(2) * K * THREADS * sizeof(float4);
  const size_t raw_smem = static_cast<size_t>(K) * ncols_max * sizeof(float);
  // measured: it pays whenever a new source row arrives every ~4 output rows (x4 up-sampling: 10.9 -> 8.1 us/tile
  // at K = 12) and for the soft outputs at any scale (prob map 16.8 -> 14.1 us/tile at x8); the plain argmax path at
  // x8 refills only every 8 rows and the two barriers per refill cost what the hidden latency gains (0.131 vs 0.138 ms)
  const int coop = (sx > 0.f && sx <= 0.5f && (extra || sx > 0.2f) &&
                    base_smem + raw_smem <= (THREADS == 256 ? 110 * 1024 : 48 * 1024)) ? 1 : 0;
  const size_t smem = base_smem + (coop ? raw_smem : 0);
  auto kern = !extra ? (K == 8 ? upsample_rows_kernel<THREADS, -8> : K == 12 ? upsample_rows_kernel<THREADS, -12>
                                                                             : upsample_rows_kernel<THREADS, 0>)
              : K == 8 ? (logits_hr ? upsample_rows_kernel<THREADS, 8, true> : upsample_rows_kernel<THREADS, 8, true, false>)
              : K < 8 ? upsample_rows_kernel<THREADS, 8>
              : K == 12 ? (logits_hr ? upsample_rows_kernel<THREADS, 12, true> : upsample_rows_kernel<THREADS, 12, true, false>)
              : K < 12 ? upsample_rows_kernel<THREADS, 12>
              : K <= 16 ? upsample_rows_kernel<THREADS, 16> : upsample_rows_kernel<THREADS, 32>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  dim3 grid((W / 4 + THREADS - 1) / THREADS, (H + rows_per_band - 1) / rows_per_band, B);
  kern<<<grid, THREADS, smem, st>>>(logits_lr, K, h, w, H, W, rows_per_band, sy, sx, label, ignore_label, pred, conf,
                                    probs, logits_hr, cm, coop);
  return SL_LAUNCH_RESULT();
}

// ------------------------------------------------------------------ pseudo-labelling
__global__ void __launch_bounds__(256) pseudo_label_kernel(const float* __restrict__ preds2, int B, int K2, int h, int w,
                                                           int H, int W, float sy, float sx, int n_base,
                                                           long long* __restrict__ mask) {
  const long long total = static_cast<long long>(B) * H * W;
  for (long long g = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; g < total;
       g += static_cast<long long>(gridDim.x) * 256) {
    if (mask[g] != 0) continue;
    const int x = static_cast<int>(g % W);
    const long long rowid = g / W;
    const int y = static_cast<int>(rowid % H);
    const int b = static_cast<int>(rowid / H);
    const SrcCoord cy = src_coord(sy, y, h), cx = src_coord(sx, x, w);
    const float* plane = preds2 + static_cast<size_t>(b) * K2 * h * w;
    float best = -INFINITY;
    int idx = 0;
    for (int k = 0; k < K2; ++k) argmax_step(bilerp(plane + static_cast<size_t>(k) * h * w, w, cy, cx), k, best, idx);
    mask[g] = idx > 0 ? idx + n_base : 0;
  }
}

// Same result, bit for bit, with the gathers hoisted: a thread owns one source cell of one output row (the four
// corner values of every channel in registers) and visits the cell's output pixels; only pixels whose mask is 0
// are evaluated, with ATen's association l0y*(l0x*a + l1x*b) + l1y*(l0x*c + l1x*d) as in bilerp().
// grid (ceil(w/32), ceil(H/8), B), block 256 = 32 cells x 8 rows.  K2 <= KP.
template <int KP>
__global__ void __launch_bounds__(256) pseudo_label_cells_kernel(const float* __restrict__ preds2, int K2, int h, int w,
                                                                 int H, int W, float sy, float sx, int n_base,
                                                                 long long* __restrict__ mask) {
  const int j = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5), b = blockIdx.z;
  if (j >= w || y >= H) return;
  // output columns whose left source column is j (src_coord(sx, x, w).i0 == j): a contiguous range
  auto first_x = [&](int col) {
    if (col <= 0) return 0;
    if (col >= w) return W;
    int x = sx > 0.f ? static_cast<int>(ceilf(static_cast<float>(col) / sx)) : W;
    x = max(0, min(W, x));
    while (x > 0 && src_coord(sx, x - 1, w).i0 >= col) --x;
    while (x < W && src_coord(sx, x, w).i0 < col) ++x;
    return x;
  };
  const int x_lo = first_x(j), x_hi = first_x(j + 1) - 1;
  if (x_hi < x_lo) return;
  long long* mrow = mask + (static_cast<size_t>(b) * H + y) * W;
  bool any = false;
  for (int x = x_lo; x <= x_hi; ++x) any |= mrow[x] == 0;
  if (!any) return;
  const SrcCoord cy = src_coord(sy, y, h);
  const int hw = h * w, j1 = min(j + 1, w - 1);
  const float* r0 = preds2 + static_cast<size_t>(b) * K2 * hw + cy.i0 * w;
  const float* r1 = r0 + cy.step * w;
  float a[KP], bb[KP], c[KP], d[KP];
#pragma unroll
  for (int k = 0; k < KP; ++k)
    if (k < K2) {
      a[k] = __ldg(r0 + j); bb[k] = __ldg(r0 + j1); c[k] = __ldg(r1 + j); d[k] = __ldg(r1 + j1);
      r0 += hw; r1 += hw;
    }
  for (int x = x_lo; x <= x_hi; ++x) {
    if (mrow[x] != 0) continue;
    const SrcCoord cx = src_coord(sx, x, w);
    // cx.step == 0 only in the last column, where j1 == j already
    float best = -INFINITY;
    int idx = 0;
#pragma unroll
    for (int k = 0; k < KP; ++k)
      if (k < K2) argmax_step(cy.l0 * (cx.l0 * a[k] + cx.l1 * bb[k]) + cy.l1 * (cx.l0 * c[k] + cx.l1 * d[k]), k, best, idx);
    mrow[x] = idx > 0 ? idx + n_base : 0;
  }
}

// ------------------------------------------------------------------ confusion on label maps
__global__ void __launch_bounds__(256) confusion_kernel(const uint8_t* __restrict__ gt, const uint8_t* __restrict__ pr,
                                                        long long n, int K, int ignore_label,
                                                        unsigned long long* __restrict__ cm,
                                                        unsigned long long* __restrict__ n_bad) {
  __shared__ unsigned int hist[SL_MAX_CLASSES * SL_MAX_CLASSES];
  __shared__ unsigned int bad;
  for (int i = threadIdx.x; i < K * K; i += 256) hist[i] = 0u;
  if (threadIdx.x == 0) bad = 0u;
  __syncthreads();
  const bool vec = ((reinterpret_cast<uintptr_t>(gt) | reinterpret_cast<uintptr_t>(pr)) & 15) == 0;
  const long long nvec = vec ? n / 16 : 0;
  const long long stride = static_cast<long long>(gridDim.x) * 256;
  const long long start = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  unsigned int my_bad = 0;
  // run-length accumulation: consecutive pixels of a thread mostly fall in one bin
  int run_bin = 0;
  unsigned int run_cnt = 0;
  auto add = [&](int g, int p) {
    if (g == ignore_label) return;
    if (g >= K || p >= K) { ++my_bad; return; }
    const int bin = g * K + p;
    if (bin == run_bin) { ++run_cnt; return; }
    if (run_cnt) atomicAdd(&hist[run_bin], run_cnt);
    run_bin = bin;
    run_cnt = 1u;
  };
  for (long long v = start; v < nvec; v += stride) {
    const uint4 g4 = ld_stream_u4(gt + v * 16), p4 = ld_stream_u4(pr + v * 16);
    const uint32_t gw[4] = {g4.x, g4.y, g4.z, g4.w};
    const uint32_t pw[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      // fast path: four equal (gt, pred) byte pairs in this word pair
      const uint32_t g0 = gw[q] & 0xffu, p0 = pw[q] & 0xffu;
      if (gw[q] == g0 * 0x01010101u && pw[q] == p0 * 0x01010101u && static_cast<int>(g0) != ignore_label &&
          static_cast<int>(g0) < K && static_cast<int>(p0) < K && static_cast<int>(g0) * K + static_cast<int>(p0) == run_bin) {
        run_cnt += 4u;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) add((gw[q] >> (8 * j)) & 0xff, (pw[q] >> (8 * j)) & 0xff);
      }
    }
  }
  // scalar tail (and the whole array when the pointers are not 16-byte aligned)
  for (long long i = nvec * 16 + start; i < n; i += stride) add(gt[i], pr[i]);
  if (run_cnt) atomicAdd(&hist[run_bin], run_cnt);
  if (my_bad) atomicAdd(&bad, my_bad);
  __syncthreads();
  for (int i = threadIdx.x; i < K * K; i += 256)
    if (hist[i]) atomicAdd(&cm[i], static_cast<unsigned long long>(hist[i]));
  if (threadIdx.x == 0 && bad && n_bad) atomicAdd(n_bad, static_cast<unsigned long long>(bad));
}

// ------------------------------------------------------------------ intersection / union
// ws: [3][K] int64 = {intersection, area_output, area_target}.  One joint (output, target) histogram of (K+1)^2 bins
// per CTA (bin K = "outside [0,K)", which torch.histc ignores), filled by per-thread run-length accumulation like
// the confusion kernel (prediction / label maps are piecewise constant along a row), then folded into the three
// marginals: intersection = diagonal, area_output = row sums, area_target = column sums.
__global__ void __launch_bounds__(256) inter_union_kernel(long long* __restrict__ output,
                                                          const long long* __restrict__ target, long long n, int K,
                                                          long long ignore, unsigned long long* __restrict__ ws) {
  __shared__ unsigned int hist[(SL_MAX_CLASSES + 1) * (SL_MAX_CLASSES + 1)];
  const int K1 = K + 1;
  for (int i = threadIdx.x; i < K1 * K1; i += 256) hist[i] = 0u;
  __syncthreads();
  int run_bin = 0;
  unsigned int run_cnt = 0;
  auto add = [&](long long i, long long o, long long t) {
    if (t == ignore && o != ignore) { o = ignore; output[i] = ignore; }          // utils/pyt_utils.py:299
    const int oi = (o >= 0 && o < K) ? static_cast<int>(o) : K;
    const int ti = (t >= 0 && t < K) ? static_cast<int>(t) : K;
    const int bin = oi * K1 + ti;
    if (bin == run_bin) { ++run_cnt; return; }
    if (run_cnt) atomicAdd(&hist[run_bin], run_cnt);
    run_bin = bin;
    run_cnt = 1u;
  };
  const bool vec = ((reinterpret_cast<uintptr_t>(output) | reinterpret_cast<uintptr_t>(target)) & 15) == 0;
  const long long nvec = vec ? n / 2 : 0;
  const long long stride = static_cast<long long>(gridDim.x) * 256;
  const long long start = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  // a thread owns runs of 4 consecutive 16-byte pairs (8 pixels) so that its run-length counter sees neighbours
  for (long long v = start * 4; v < nvec; v += stride * 4) {
    longlong2 o2[4], t2[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (v + u < nvec) {
        o2[u] = *reinterpret_cast<const longlong2*>(output + 2 * (v + u));
        t2[u] = __ldg(reinterpret_cast<const longlong2*>(target + 2 * (v + u)));
      }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (v + u < nvec) {
        add(2 * (v + u), o2[u].x, t2[u].x);
        add(2 * (v + u) + 1, o2[u].y, t2[u].y);
      }
  }
  for (long long i = nvec * 2 + start; i < n; i += stride) add(i, output[i], target[i]);
  if (run_cnt) atomicAdd(&hist[run_bin], run_cnt);
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += 256) {
    unsigned long long row = 0, col = 0;
    for (int j = 0; j < K1; ++j) { row += hist[k * K1 + j]; col += hist[j * K1 + k]; }
    if (hist[k * K1 + k]) atomicAdd(&ws[k], static_cast<unsigned long long>(hist[k * K1 + k]));
    if (row) atomicAdd(&ws[K + k], row);
    if (col) atomicAdd(&ws[2 * K + k], col);
  }
}

__global__ void inter_union_finish_kernel(const unsigned long long* __restrict__ ws, int K, float* __restrict__ inter,
                                          float* __restrict__ uni, float* __restrict__ tgt) {
  const int k = threadIdx.x;
  if (k >= K) return;
  // torch.histc returns fp32 counts; union = area_output + area_target - intersection in fp32
  const float fi = static_cast<float>(ws[k]), fo = static_cast<float>(ws[K + k]), ft = static_cast<float>(ws[2 * K + k]);
  inter[k] = fi;
  uni[k] = fo + ft - fi;
  tgt[k] = ft;
}

// ------------------------------------------------------------------ view aggregation
struct FlipFlags { int f[16]; };
__global__ void __launch_bounds__(256) views_reduce_kernel(const float* __restrict__ views, int V, long long planes,
                                                           int h, int w, FlipFlags flips, float scale,
                                                           float* __restrict__ out) {
  const long long total = planes * h * w;
  for (long long g = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; g < total;
       g += static_cast<long long>(gridDim.x) * 256) {
    const int x = static_cast<int>(g % w);
    const long long r = g / w;
    const int y = static_cast<int>(r % h);
    const long long plane = r / h;
    float acc = 0.f;
    for (int v = 0; v < V; ++v) {
      const int xs = (flips.f[v] & 1) ? w - 1 - x : x;
      const int ys = (flips.f[v] & 2) ? h - 1 - y : y;
      acc += views[((static_cast<size_t>(v) * planes + plane) * h + ys) * w + xs];
    }
    out[g] = acc * scale;
  }
}

// w % 4 == 0: four consecutive x per thread, 128-bit loads; a horizontally flipped view reads the mirrored float4
// and reverses its components.
__global__ void __launch_bounds__(256) views_reduce_vec4_kernel(const float* __restrict__ views, int V, long long planes,
                                                                int h, int w, FlipFlags flips, float scale,
                                                                float* __restrict__ out) {
  const int w4 = w >> 2;
  const long long total = planes * h * w4;
  for (long long g = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; g < total;
       g += static_cast<long long>(gridDim.x) * 256) {
    const int x = static_cast<int>(g % w4) * 4;
    const long long r = g / w4;
    const int y = static_cast<int>(r % h);
    const long long plane = r / h;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int v = 0; v < V; ++v) {
      const bool fx = flips.f[v] & 1;
      const int xs = fx ? w - 4 - x : x;
      const int ys = (flips.f[v] & 2) ? h - 1 - y : y;
      const float4 t = ld_stream_f4(views + ((static_cast<size_t>(v) * planes + plane) * h + ys) * w + xs);
      acc.x += fx ? t.w : t.x; acc.y += fx ? t.z : t.y; acc.z += fx ? t.y : t.z; acc.w += fx ? t.x : t.w;
    }
    __stcs(reinterpret_cast<float4*>(out + (plane * h + y) * w + x),
           make_float4(acc.x * scale, acc.y * scale, acc.z * scale, acc.w * scale));
  }
}

int launch_upsample_prune(const float* logits_lr, int B, int K, int h, int w, int H, int W, float sy, float sx,
                          const uint8_t* label, int ignore_label, uint8_t* pred, unsigned long long* cm,
                          cudaStream_t st);                      // post_prune.cu
int launch_upsample_regs(const float* logits_lr, int B, int K, int h, int w, int H, int W, float sy, float sx,
                         const uint8_t* label, int ignore_label, uint8_t* pred, float* conf, float* probs,
                         unsigned long long* cm, cudaStream_t st);   // post_regs.cu

static inline int grid_for(long long work_items, int per_sm) {
  long long blocks = (work_items + 255) / 256;
  const long long cap = static_cast<long long>(num_sms()) * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

}  // namespace sl

extern "C" int sl_upsample_argmax(const float* logits_lr, int B, int K, int h, int w, int H, int W,
                                  const uint8_t* label, int ignore_label, uint8_t* pred, float* conf, float* probs,
                                  float* logits_hr, long long* cm, void* stream) {
  SL_CHECK_PTR(logits_lr);
  SL_CHECK_ARG(B >= 1 && K >= 1 && K <= SL_MAX_CLASSES && h >= 1 && w >= 1 && H >= 1 && W >= 1);
  SL_CHECK_ARG(static_cast<long long>(h) * w * K < (1ll << 31));
  if (cm) SL_CHECK_PTR(label);
  SL_CHECK_ARG(pred || conf || probs || logits_hr || cm);
  if (W % 4 == 0) {
    if (label) SL_CHECK_ALIGN(label, 4);
    if (pred) SL_CHECK_ALIGN(pred, 4);
    if (conf) SL_CHECK_ALIGN(conf, 16);
    if (probs) SL_CHECK_ALIGN(probs, 16);
    if (logits_hr) SL_CHECK_ALIGN(logits_hr, 16);
  }
  const float sy = sl::ac_scale(h, H), sx = sl::ac_scale(w, W);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  auto* cmu = reinterpret_cast<unsigned long long*>(cm);
  if (W % 4 == 0 && B <= 65535) {
    // Row-cached fast path.  Pick the CTA shape so that small problems still fill 148 SMs:
    // shared memory per CTA = 2*K*THREADS*16 B (K=12, 256 threads: 96 KB -> 2 CTAs/SM).
    const long long px = static_cast<long long>(B) * H * W;
    const bool big = px >= (8ll << 20) && W >= 1024 && K <= 12;
    int rows = big ? 32 : 16;
    if (big && (conf || probs || logits_hr)) {
      // Soft outputs: 128 registers and 96 KB of row cache allow two 256-thread CTAs per SM, so 16 tiles x 32 bands =
      // 512 CTAs run as 1.7 waves of 296.  Pick the band height that fills whole waves (28 rows: 37 bands x 16 tiles =
      // 592 CTAs = 2.0 waves); cost = waves x (rows + the band's source-row refills at ~1.5 rows each).
      const long long slots = 2ll * sl::num_sms();
      const long long gx = (W / 4 + 255) / 256;
      double best = 1e30;
      for (int r = 16; r <= 64; r += 2) {
        const long long ctas = gx * ((H + r - 1) / r) * B;
        const double cost = static_cast<double>((ctas + slots - 1) / slots) * (r + 1.5 * (r * static_cast<double>(sy) + 2.0));
        if (cost < best) { best = cost; rows = r; }
      }
    }
    // When the prediction map is written anyway, the confusion matrix is accumulated by a second launch over
    // (label, pred) -- the map is still in L2 -- instead of inside the interpolation kernel: that kernel is bound by
    // instruction issue, and on noisy predictions (runs of equal (label, pred) pairs broken every few pixels) the
    // in-kernel histogram cost 0.073 ms per 32 tiles against 0.046 ms for the separate pass.  Without a pred buffer
    // the counting stays fused.
    const int fused_env = sl::env().post_fused_cm;                     // 1 / 0: force the counting inside / outside the kernel
    const int prune_env = sl::env().post_prune;                        // 0: always the row-cached kernel (A/B runs)
    if (prune_env != 0 && pred != nullptr && !conf && !probs && !logits_hr) {
      // prediction-only path: per-cell class pruning (post_prune.cu), confusion counted in the same pass
      const bool fused = cm != nullptr && fused_env != 0;
      const int rc = sl::launch_upsample_prune(logits_lr, B, K, h, w, H, W, sy, sx, fused ? label : nullptr, ignore_label,
                                               pred, fused ? cmu : nullptr, st);
      if (rc != -100) {
        if (rc != 0 || cm == nullptr || fused) return rc;
        sl::confusion_kernel<<<sl::grid_for((px + 15) / 16, 8), 256, 0, st>>>(label, pred, px, K, ignore_label, cmu, nullptr);
        return SL_LAUNCH_RESULT();
      }
    }
    if (sl::env().post_regs != 0 && !logits_hr) {
      // K = 8 / 12, up-sampling by >= 2x, any batch size: source-row intervals in registers (post_regs.cu) -- the
      // prediction-only path (one tile: 5 us against 15 us for the row-cached kernel's small-problem shape) and, at
      // K = 12, the soft outputs (probability maps at 76 % of the HBM copy peak against 71-73 %)
      const bool fused = cm != nullptr && (pred == nullptr || fused_env != 0);
      const int rc = sl::launch_upsample_regs(logits_lr, B, K, h, w, H, W, sy, sx, fused ? label : nullptr, ignore_label,
                                              pred, conf, probs, fused ? cmu : nullptr, st);
      if (rc != -100) {
        if (rc != 0 || cm == nullptr || fused) return rc;
        sl::confusion_kernel<<<sl::grid_for((px + 15) / 16, 8), 256, 0, st>>>(label, pred, px, K, ignore_label, cmu, nullptr);
        return SL_LAUNCH_RESULT();
      }
    }
    // Row-cached kernel: counting inside the interpolation kernel (per-thread vertical run lengths) is now the default:
    // re-measured in round 2 on both coherent and random predictions it is 3-5 % cheaper than a second launch over
    // (label, pred) (profiles/r2_post_probe.txt); SL_POST_FUSED_CM=0 restores the second launch.
    const bool split_cm = cm != nullptr && pred != nullptr && fused_env == 0;
    const uint8_t* k_label = split_cm ? nullptr : label;
    unsigned long long* k_cm = split_cm ? nullptr : cmu;
    const int rc = big ? sl::launch_rows<256>(logits_lr, B, K, h, w, H, W, rows, sy, sx, k_label, ignore_label, pred, conf,
                                              probs, logits_hr, k_cm, st)
                       : sl::launch_rows<64>(logits_lr, B, K, h, w, H, W, rows, sy, sx, k_label, ignore_label, pred, conf,
                                             probs, logits_hr, k_cm, st);
    if (rc != 0 || !split_cm) return rc;
    sl::confusion_kernel<<<sl::grid_for((px + 15) / 16, 8), 256, 0, st>>>(label, pred, px, K, ignore_label, cmu, nullptr);
    return SL_LAUNCH_RESULT();
  }
  const long long items = static_cast<long long>(B) * H * ((W + sl::UP_PX - 1) / sl::UP_PX);
  const int grid = sl::grid_for(items, 8);
#define SL_UP_LAUNCH(KP) sl::upsample_argmax_kernel<KP><<<grid, 256, 0, st>>>( \
      logits_lr, B, K, h, w, H, W, sy, sx, label, ignore_label, pred, conf, probs, logits_hr, cmu)
  if (!(conf || probs)) SL_UP_LAUNCH(0);
  else if (K <= 8) SL_UP_LAUNCH(8);
  else if (K <= 12) SL_UP_LAUNCH(12);
  else if (K <= 16) SL_UP_LAUNCH(16);
  else SL_UP_LAUNCH(32);
#undef SL_UP_LAUNCH
  return SL_LAUNCH_RESULT();
}

extern "C" int sl_pseudo_label(const float* preds2, int B, int K2, int h, int w, int H, int W, int n_base,
                               long long* mask, void* stream) {
  SL_CHECK_PTR(preds2); SL_CHECK_PTR(mask);
  SL_CHECK_ARG(B >= 1 && K2 >= 1 && K2 <= SL_MAX_CLASSES && h >= 1 && w >= 1 && H >= 1 && W >= 1 && n_base >= 0);
  const float sy = sl::ac_scale(h, H), sx = sl::ac_scale(w, W);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (K2 <= 8 && B <= 65535 && (H + 7) / 8 <= 65535) {            // 1 + Kn channels: OEM has 5
    const dim3 grid((w + 31) / 32, (H + 7) / 8, B);
    if (K2 <= 5) sl::pseudo_label_cells_kernel<5><<<grid, 256, 0, st>>>(preds2, K2, h, w, H, W, sy, sx, n_base, mask);
    else sl::pseudo_label_cells_kernel<8><<<grid, 256, 0, st>>>(preds2, K2, h, w, H, W, sy, sx, n_base, mask);
    return SL_LAUNCH_RESULT();
  }
  const long long items = static_cast<long long>(B) * H * W;
  sl::pseudo_label_kernel<<<sl::grid_for(items, 16), 256, 0, st>>>(preds2, B, K2, h, w, H, W, sy, sx, n_base, mask);
  return SL_LAUNCH_RESULT();
}

extern "C" int sl_confusion(const uint8_t* gt, const uint8_t* pred, long long n, int K, int ignore_label,
                            long long* cm, long long* n_bad, void* stream) {
  SL_CHECK_ARG(n >= 0 && K >= 1 && K <= SL_MAX_CLASSES);
  SL_CHECK_PTR(cm);
  if (n == 0) return SL_OK;   // empty maps are legal (and may come with NULL data pointers)
  SL_CHECK_PTR(gt); SL_CHECK_PTR(pred);
  sl::confusion_kernel<<<sl::grid_for((n + 15) / 16, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      gt, pred, n, K, ignore_label, reinterpret_cast<unsigned long long*>(cm),
      reinterpret_cast<unsigned long long*>(n_bad));
  return SL_LAUNCH_RESULT();
}

extern "C" int sl_inter_union(long long* output, const long long* target, long long n, int K, int ignore_label,
                              float* inter, float* uni, float* tgt, long long* ws, void* stream) {
  SL_CHECK_PTR(inter); SL_CHECK_PTR(uni); SL_CHECK_PTR(tgt); SL_CHECK_PTR(ws);
  SL_CHECK_ARG(n >= 0 && K >= 1 && K <= SL_MAX_CLASSES);
  if (n > 0) { SL_CHECK_PTR(output); SL_CHECK_PTR(target); }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(ws, 0, sizeof(long long) * 3 * K, st);
  if (e != cudaSuccess) return static_cast<int>(e);
  if (n > 0)
    sl::inter_union_kernel<<<sl::grid_for(n, 8), 256, 0, st>>>(output, target, n, K, ignore_label,
                                                              reinterpret_cast<unsigned long long*>(ws));
  sl::inter_union_finish_kernel<<<1, 32, 0, st>>>(reinterpret_cast<unsigned long long*>(ws), K, inter, uni, tgt);
  return SL_LAUNCH_RESULT();
}

extern "C" int sl_views_reduce(const float* views, int V, int B, int K, int h, int w, const int* flip_host,
                               float scale, float* out, void* stream) {
  SL_CHECK_PTR(views); SL_CHECK_PTR(out); SL_CHECK_PTR(flip_host);
  SL_CHECK_ARG(V >= 1 && V <= 16 && B >= 1 && K >= 1 && h >= 1 && w >= 1);
  sl::FlipFlags ff;
  for (int v = 0; v < 16; ++v) ff.f[v] = v < V ? flip_host[v] : 0;
  const long long planes = static_cast<long long>(B) * K;
  if (w % 4 == 0 && reinterpret_cast<uintptr_t>(views) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0)
    sl::views_reduce_vec4_kernel<<<sl::grid_for(planes * h * (w / 4), 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        views, V, planes, h, w, ff, scale, out);
  else
    sl::views_reduce_kernel<<<sl::grid_for(planes * h * w, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        views, V, planes, h, w, ff, scale, out);
  return SL_LAUNCH_RESULT();
}
