//  The segmentation cross-entropy tail of OrthLoss.forward / CELoss.forward,
//   loss/criterion.py:17-19 and :51-52:
//     scale_pred = F.interpolate(pred, target.shape[1:], mode='bilinear', align_corners=True)
//     loss = nn.CrossEntropyLoss(ignore_index=255, reduction='mean')(scale_pred, target)
// fused so that the [B,K,H,W] up-sampled logits (50 MB per 1024^2 tile at K = 12) are never written:
// the forward keeps one fp32 log-sum-exp per pixel (4 MB per tile) for the backward, which gathers,
// per low-resolution logit, the bilinear-weighted (softmax - onehot) of the output pixels it feeds.
// Both directions are deterministic (fixed-order reductions, no floating-point atomics).
#include "common.cuh"

namespace sl {

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;

// Output pixels x whose left source column is j ("cell" j): src_coord(scale, x, in).i0 == j.  i0 is non-decreasing
// in x, so the cell is a contiguous range [lo, hi]; empty when hi < lo (down-sampling).
__device__ __forceinline__ int cell_begin(float scale, int j, int in_size, int out_size) {
  if (j <= 0) return 0;
  if (j >= in_size) return out_size;
  int x = scale > 0.f ? static_cast<int>(ceilf(static_cast<float>(j) / scale)) : out_size;
  x = max(0, min(out_size, x));
  while (x > 0 && src_coord(scale, x - 1, in_size).i0 >= j) --x;
  while (x < out_size && src_coord(scale, x, in_size).i0 < j) ++x;
  return x;
}

// ---- forward: one (image, output row, source cell) per thread.  The 2 x 2 x K low-resolution values of the cell
// are blended vertically once (adjacent threads read adjacent columns: coalesced), then every output pixel of the
// cell costs one horizontal lerp + exp per class.  Per-CTA partial loss sums in double, fixed order.
// EXACT: K == KP, so the class loops carry no predicates.
template <int KP, bool EXACT>
__global__ void __launch_bounds__(256) ce_fwd_kernel(const float* __restrict__ logits_lr, int B, int K, int h, int w,
                                                     int H, int W, float sy, float sx,
                                                     const long long* __restrict__ target, long long ignore,
                                                     float* __restrict__ lse_out, double* __restrict__ part_sum,
                                                     unsigned long long* __restrict__ part_cnt) {
  __shared__ double red_s[8];
  __shared__ unsigned long long red_c[8];
  // work item = (image, group of 8 output rows, block of 32 source cells); thread = (cell, row) inside it
  const int jb_n = (w + 31) / 32, yg_n = (H + 7) / 8;
  const int items = B * yg_n * jb_n;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int hw = h * w;
  double acc = 0.0;
  unsigned int cnt = 0, bad = 0;                              // bad: targets outside [0,K) that are not ignore_index
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int jb = item % jb_n, t2 = item / jb_n;
    const int yg = t2 % yg_n, b = t2 / yg_n;
    const int j = jb * 32 + tx, y = yg * 8 + ty;
    if (j >= w || y >= H) continue;
    const int x_lo = cell_begin(sx, j, w, W), x_hi = cell_begin(sx, j + 1, w, W) - 1;
    if (x_hi < x_lo) continue;
    const SrcCoord cy = src_coord(sy, y, h);
    const int j1 = min(j + 1, w - 1);
    const float* r0 = logits_lr + static_cast<size_t>(b) * K * hw + cy.i0 * w;
    const float* r1 = r0 + cy.step * w;
    float va[KP], vb[KP];                                   // the two source columns, vertically blended
#pragma unroll
    for (int k = 0; k < KP; ++k)
      if (EXACT || k < K) {
        va[k] = cy.l0 * __ldg(r0 + j) + cy.l1 * __ldg(r1 + j);
        vb[k] = cy.l0 * __ldg(r0 + j1) + cy.l1 * __ldg(r1 + j1);
        r0 += hw; r1 += hw;
      }
    const size_t row_px = (static_cast<size_t>(b) * H + y) * W;
    for (int x = x_lo; x <= x_hi; ++x) {
      const SrcCoord cx = src_coord(sx, x, w);
      const long long t = target[row_px + x];
      const int ti = (t != ignore && t >= 0 && t < K) ? static_cast<int>(t) : -1;
      // pass 1: maximum and the target's logit; pass 2 re-blends (2 instructions) instead of keeping K values live
      float m = -INFINITY, vt = 0.f;
#pragma unroll
      for (int k = 0; k < KP; ++k)
        if (EXACT || k < K) {
          const float v = cx.l0 * va[k] + cx.l1 * vb[k];
          m = fmaxf(m, v);
          vt = (k == ti) ? v : vt;
        }
      const float m2 = m * kLog2e;
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < KP; ++k)
        if (EXACT || k < K) s += ex2_approx(fmaf(cx.l0 * va[k] + cx.l1 * vb[k], kLog2e, -m2));
      const float lse = m + lg2_approx(s) * kLn2;
      lse_out[row_px + x] = lse;
      if (ti >= 0) {
        acc += static_cast<double>(lse - vt);
        ++cnt;
      } else if (t != ignore) {
        ++bad;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    acc += __shfl_xor_sync(0xffffffffu, acc, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    bad += __shfl_xor_sync(0xffffffffu, bad, o);
  }
  // valid count in the low 40 bits, out-of-range count above (a label map has far fewer than 2^40 pixels)
  if ((threadIdx.x & 31) == 0) {
    red_s[threadIdx.x >> 5] = acc;
    red_c[threadIdx.x >> 5] = static_cast<unsigned long long>(cnt) + (static_cast<unsigned long long>(bad) << 40);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    unsigned long long c = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { t += red_s[i]; c += red_c[i]; }
    part_sum[blockIdx.x] = t;
    part_cnt[blockIdx.x] = c;
  }
}

__global__ void ce_finish_kernel(const double* __restrict__ part_sum, const unsigned long long* __restrict__ part_cnt,
                                 int n, float* __restrict__ loss, long long* __restrict__ n_valid) {
  // one warp: lane l adds partials l, l+32, ... in order, then a fixed shuffle tree -> same result every run
  double t = 0.0;
  unsigned long long c = 0;
  for (int i = threadIdx.x; i < n; i += 32) { t += part_sum[i]; c += part_cnt[i]; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    t += __shfl_xor_sync(0xffffffffu, t, o);
    c += __shfl_xor_sync(0xffffffffu, c, o);
  }
  if (threadIdx.x == 0) {
    const unsigned long long n_ok = c & ((1ull << 40) - 1ull), n_bad = c >> 40;
    // nn.CrossEntropyLoss raises a device assert for a target outside [0,K) that is not ignore_index; here the loss
    // becomes NaN and n_valid = -(number of such targets), so wrong labels cannot train silently.
    loss[0] = (n_ok && !n_bad) ? static_cast<float>(t / static_cast<double>(n_ok)) : nanf("");   // torch: mean over zero elements = nan
    n_valid[0] = n_bad ? -static_cast<long long>(n_bad) : static_cast<long long>(n_ok);
  }
}

// ---- backward, separable and deterministic.
//   d[k](y, x) = softmax_k - onehot_k at output pixel (y, x) (0 where the target is ignored)
//   pass 1 (rows):  r[k][y][j] = sum_x wx(j, x) d[k](y, x)      one (image, output row, source column) per thread;
//                   column j is fed by cell j (weight l0) and cell j-1 (weight l1): ~2/sx pixels, each re-evaluated
//                   from the three vertically blended source columns j-1, j, j+1 kept in registers
//   pass 2 (cols):  grad[k][i][j] = scale * sum_y wy(i, y) r[k][y][j]
// r is [B][K][H][w] fp32 in the workspace (6 MB per 1024^2 tile at K = 12, not the 50 MB of the up-sampled logits).
template <int KP, bool EXACT>
__global__ void __launch_bounds__(256) ce_bwd_rows_kernel(const float* __restrict__ logits_lr, int B, int K, int h, int w,
                                                          int H, int W, float sy, float sx,
                                                          const long long* __restrict__ target, long long ignore,
                                                          const float* __restrict__ lse, float* __restrict__ r_out) {
  // grid (ceil(w/32), ceil(H/8), B), block 256 = 32 source columns x 8 output rows
  const int j = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5), b = blockIdx.z;
  if (j >= w || y >= H) return;
  const int hw = h * w;
  const SrcCoord cy = src_coord(sy, y, h);
  const float* r0 = logits_lr + static_cast<size_t>(b) * K * hw + cy.i0 * w;
  const float* r1 = r0 + cy.step * w;
  const int jm = max(j - 1, 0), jp = min(j + 1, w - 1);
  float vm[KP], vc[KP], vp[KP], acc[KP];
#pragma unroll
  for (int k = 0; k < KP; ++k) {
    acc[k] = 0.f;
    if (EXACT || k < K) {
      vm[k] = cy.l0 * __ldg(r0 + jm) + cy.l1 * __ldg(r1 + jm);
      vc[k] = cy.l0 * __ldg(r0 + j) + cy.l1 * __ldg(r1 + j);
      vp[k] = cy.l0 * __ldg(r0 + jp) + cy.l1 * __ldg(r1 + jp);
      r0 += hw; r1 += hw;
    }
  }
  const size_t row_px = (static_cast<size_t>(b) * H + y) * W;
  const int x_mid = cell_begin(sx, j, w, W);
  const int x_lo = j > 0 ? cell_begin(sx, j - 1, w, W) : x_mid, x_hi = cell_begin(sx, j + 1, w, W) - 1;
  for (int x = x_lo; x <= x_hi; ++x) {
    const long long t = target[row_px + x];
    if (t == ignore || t < 0 || t >= K) continue;
    const SrcCoord cx = src_coord(sx, x, w);
    const bool left = x < x_mid;                             // pixel of cell j-1: column j is its right neighbour
    const float wgt = left ? cx.l1 : cx.l0;
    if (wgt == 0.f) continue;
    const float l2 = lse[row_px + x] * kLog2e;
    const int ti = static_cast<int>(t);
    // the pixel's two source columns: (j-1, j) for a pixel of cell j-1, (j, j+1) for a pixel of cell j
#pragma unroll
    for (int k = 0; k < KP; ++k)
      if (EXACT || k < K) {
        const float v = cx.l0 * (left ? vm[k] : vc[k]) + cx.l1 * (left ? vc[k] : vp[k]);
        const float pr = ex2_approx(fmaf(v, kLog2e, -l2));
        acc[k] = fmaf(wgt, pr - (k == ti ? 1.f : 0.f), acc[k]);
      }
  }
#pragma unroll
  for (int k = 0; k < KP; ++k)
    if (EXACT || k < K) r_out[((static_cast<size_t>(b) * K + k) * H + y) * w + j] = acc[k];
}

__global__ void __launch_bounds__(256) ce_bwd_cols_kernel(const float* __restrict__ r_in, int B, int K, int h, int w, int H,
                                                          float sy, const long long* __restrict__ n_valid,
                                                          const float* __restrict__ grad_out, float* __restrict__ grad_lr) {
  const long long total = static_cast<long long>(B) * K * h * w;
  const long long g = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (g >= total) return;
  const int j = static_cast<int>(g % w);
  const long long t1 = g / w;
  const int i = static_cast<int>(t1 % h);
  const long long bk = t1 / h;
  const long long nv = n_valid[0];
  const float scale = nv > 0 ? grad_out[0] / static_cast<float>(nv) : nv < 0 ? nanf("") : 0.f;   // nv < 0: bad targets
  const float* col = r_in + static_cast<size_t>(bk) * H * w + j;
  const int y_mid = cell_begin(sy, i, h, H);
  const int y_lo = i > 0 ? cell_begin(sy, i - 1, h, H) : y_mid, y_hi = cell_begin(sy, i + 1, h, H) - 1;
  float acc = 0.f;
  for (int y = y_lo; y <= y_hi; ++y) {
    const SrcCoord cy = src_coord(sy, y, h);
    acc = fmaf(y < y_mid ? cy.l1 : cy.l0, col[static_cast<size_t>(y) * w], acc);
  }
  grad_lr[g] = acc * scale;
}

}  // namespace sl

extern "C" size_t sl_upsample_ce_ws_bytes(int B, int K, int w, int H, int W) {
  if (B < 1 || K < 1 || w < 1 || H < 1 || W < 1) return 0;
  // per-pixel log-sum-exp (fp32) + per-CTA partial sums (double) and counts (u64) + the backward's row-reduced
  // gradients [B][K][H][w] (fp32); 16-byte aligned sections
  const size_t lse = (static_cast<size_t>(B) * H * W * sizeof(float) + 15) / 16 * 16;
  return lse + static_cast<size_t>(sl::num_sms()) * 8 * (sizeof(double) + sizeof(unsigned long long)) +
         static_cast<size_t>(B) * K * H * w * sizeof(float);
}

extern "C" int sl_upsample_ce_fwd(const float* logits_lr, int B, int K, int h, int w, int H, int W,
                                  const long long* target, int ignore_label, void* ws, float* loss,
                                  long long* n_valid, void* stream) {
  SL_CHECK_PTR(logits_lr); SL_CHECK_PTR(target); SL_CHECK_PTR(ws); SL_CHECK_PTR(loss); SL_CHECK_PTR(n_valid);
  SL_CHECK_ARG(B >= 1 && K >= 1 && K <= SL_MAX_CLASSES && h >= 1 && w >= 1 && H >= 1 && W >= 1);
  SL_CHECK_ALIGN(ws, 16);
  const size_t lse_bytes = (static_cast<size_t>(B) * H * W * sizeof(float) + 15) / 16 * 16;
  float* lse = static_cast<float*>(ws);
  double* part_sum = reinterpret_cast<double*>(static_cast<char*>(ws) + lse_bytes);
  unsigned long long* part_cnt = reinterpret_cast<unsigned long long*>(part_sum + sl::num_sms() * 8);
  // one thread per (row, source cell), CTAs stride over (image, 8-row group, 32-cell block) items
  long long blocks = static_cast<long long>(B) * ((H + 7) / 8) * ((w + 31) / 32);
  if (blocks > sl::num_sms() * 8) blocks = sl::num_sms() * 8;
  const int grid = static_cast<int>(blocks);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float sy = sl::ac_scale(h, H), sx = sl::ac_scale(w, W);
#define SL_CE_FWD(KP, EX) sl::ce_fwd_kernel<KP, EX><<<grid, 256, 0, st>>>(logits_lr, B, K, h, w, H, W, sy, sx, target, \
                                                                           ignore_label, lse, part_sum, part_cnt)
  if (K == 8) SL_CE_FWD(8, true); else if (K == 12) SL_CE_FWD(12, true);      // OEM: 8 (base) / 12 (ft) classes
  else if (K <= 8) SL_CE_FWD(8, false); else if (K <= 12) SL_CE_FWD(12, false);
  else if (K <= 16) SL_CE_FWD(16, false); else SL_CE_FWD(32, false);
#undef SL_CE_FWD
  sl::ce_finish_kernel<<<1, 32, 0, st>>>(part_sum, part_cnt, grid, loss, n_valid);
  return SL_LAUNCH_RESULT();
}

extern "C" int sl_upsample_ce_bwd(const float* logits_lr, int B, int K, int h, int w, int H, int W,
                                  const long long* target, int ignore_label, void* ws,
                                  const long long* n_valid, const float* grad_out, float* grad_logits_lr,
                                  void* stream) {
  SL_CHECK_PTR(logits_lr); SL_CHECK_PTR(target); SL_CHECK_PTR(ws); SL_CHECK_PTR(n_valid); SL_CHECK_PTR(grad_out);
  SL_CHECK_PTR(grad_logits_lr);
  SL_CHECK_ARG(B >= 1 && K >= 1 && K <= SL_MAX_CLASSES && h >= 1 && w >= 1 && H >= 1 && W >= 1);
  const float* lse = static_cast<const float*>(ws);
  const size_t lse_bytes = (static_cast<size_t>(B) * H * W * sizeof(float) + 15) / 16 * 16;
  float* r = reinterpret_cast<float*>(static_cast<char*>(ws) + lse_bytes +
                                      static_cast<size_t>(sl::num_sms()) * 8 * (sizeof(double) + sizeof(unsigned long long)));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float sy = sl::ac_scale(h, H), sx = sl::ac_scale(w, W);
  SL_CHECK_ARG(B <= 65535 && (H + 7) / 8 <= 65535);
  const dim3 grid1((w + 31) / 32, (H + 7) / 8, B);
#define SL_CE_BWD(KP, EX) sl::ce_bwd_rows_kernel<KP, EX><<<grid1, 256, 0, st>>>(logits_lr, B, K, h, w, H, W, sy, sx, \
                                                                                 target, ignore_label, lse, r)
  if (K == 8) SL_CE_BWD(8, true); else if (K == 12) SL_CE_BWD(12, true);
  else if (K <= 8) SL_CE_BWD(8, false); else if (K <= 12) SL_CE_BWD(12, false);
  else if (K <= 16) SL_CE_BWD(16, false); else SL_CE_BWD(32, false);
#undef SL_CE_BWD
  const long long total = static_cast<long long>(B) * K * h * w;
  sl::ce_bwd_cols_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, st>>>(r, B, K, h, w, H, sy, n_valid, grad_out,
                                                                               grad_logits_lr);
  return SL_LAUNCH_RESULT();
}
