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

__device__ __forceinline__ float bilerp_at(const float* __restrict__ src, int w, const SrcCoord& cy, const SrcCoord& cx) {
  const float* r0 = src + cy.i0 * w + cx.i0;
  const float* r1 = r0 + cy.step * w;
  const float a = __ldg(r0), b = __ldg(r0 + cx.step), c = __ldg(r1), d = __ldg(r1 + cx.step);
  return cy.l0 * (cx.l0 * a + cx.l1 * b) + cy.l1 * (cx.l0 * c + cx.l1 * d);
}

// ---- forward: one output pixel per thread (grid-stride); per-CTA partial sums in double
template <int KP>
__global__ void __launch_bounds__(256) ce_fwd_kernel(const float* __restrict__ logits_lr, int B, int K, int h, int w,
                                                     int H, int W, float sy, float sx,
                                                     const long long* __restrict__ target, long long ignore,
                                                     float* __restrict__ lse_out, double* __restrict__ part_sum,
                                                     unsigned long long* __restrict__ part_cnt) {
  __shared__ double red_s[8];
  __shared__ unsigned int red_c[8];
  const long long total = static_cast<long long>(B) * H * W;
  double acc = 0.0;
  unsigned int cnt = 0;
  for (long long g = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; g < total;
       g += static_cast<long long>(gridDim.x) * 256) {
    const int x = static_cast<int>(g % W);
    const long long rowid = g / W;
    const int y = static_cast<int>(rowid % H);
    const int b = static_cast<int>(rowid / H);
    const SrcCoord cy = src_coord(sy, y, h), cx = src_coord(sx, x, w);
    const float* plane = logits_lr + static_cast<size_t>(b) * K * h * w;
    float v[KP];
    float m = -INFINITY;
#pragma unroll
    for (int k = 0; k < KP; ++k)
      if (k < K) { v[k] = bilerp_at(plane + static_cast<size_t>(k) * h * w, w, cy, cx); m = fmaxf(m, v[k]); }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < KP; ++k)
      if (k < K) s += expf(v[k] - m);
    const float lse = m + logf(s);
    lse_out[g] = lse;
    const long long t = target[g];
    if (t != ignore && t >= 0 && t < K) {
      float vt = 0.f;
#pragma unroll
      for (int k = 0; k < KP; ++k) vt = (k == static_cast<int>(t)) ? v[k] : vt;
      acc += static_cast<double>(lse - vt);
      ++cnt;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    acc += __shfl_xor_sync(0xffffffffu, acc, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if ((threadIdx.x & 31) == 0) { red_s[threadIdx.x >> 5] = acc; red_c[threadIdx.x >> 5] = cnt; }
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
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double t = 0.0;
  unsigned long long c = 0;
  for (int i = 0; i < n; ++i) { t += part_sum[i]; c += part_cnt[i]; }
  loss[0] = c ? static_cast<float>(t / static_cast<double>(c)) : nanf("");   // torch: mean over zero elements = nan
  n_valid[0] = static_cast<long long>(c);
}

// ---- backward: one low-resolution pixel (b, i, j) per thread, all K channels in registers
template <int KP>
__global__ void __launch_bounds__(128) ce_bwd_kernel(const float* __restrict__ logits_lr, int B, int K, int h, int w, int H,
                                                     int W, float sy, float sx, const long long* __restrict__ target,
                                                     long long ignore, const float* __restrict__ lse,
                                                     const long long* __restrict__ n_valid,
                                                     const float* __restrict__ grad_out, float* __restrict__ grad_lr) {
  const long long total = static_cast<long long>(B) * h * w;
  const long long g = static_cast<long long>(blockIdx.x) * 128 + threadIdx.x;
  if (g >= total) return;
  const int j = static_cast<int>(g % w);
  const long long r = g / w;
  const int i = static_cast<int>(r % h);
  const int b = static_cast<int>(r / h);
  const long long nv = n_valid[0];
  const float scale = nv > 0 ? grad_out[0] / static_cast<float>(nv) : 0.f;
  const float* plane = logits_lr + static_cast<size_t>(b) * K * h * w;
  float acc[KP];
#pragma unroll
  for (int k = 0; k < KP; ++k) acc[k] = 0.f;
  // output rows / columns whose source interval touches (i, j): src = scale * dst lies in (i-1, i+1)
  const float inv_sy = sy > 0.f ? 1.f / sy : 0.f, inv_sx = sx > 0.f ? 1.f / sx : 0.f;
  int y_lo = sy > 0.f ? max(0, static_cast<int>(floorf((i - 1) * inv_sy)) - 1) : 0;
  int y_hi = sy > 0.f ? min(H - 1, static_cast<int>(ceilf((i + 1) * inv_sy)) + 1) : H - 1;
  int x_lo = sx > 0.f ? max(0, static_cast<int>(floorf((j - 1) * inv_sx)) - 1) : 0;
  int x_hi = sx > 0.f ? min(W - 1, static_cast<int>(ceilf((j + 1) * inv_sx)) + 1) : W - 1;
  for (int y = y_lo; y <= y_hi; ++y) {
    const SrcCoord cy = src_coord(sy, y, h);
    const float wy = (cy.i0 == i ? cy.l0 : 0.f) + (cy.i0 + cy.step == i ? cy.l1 : 0.f);
    if (wy == 0.f) continue;
    for (int x = x_lo; x <= x_hi; ++x) {
      const SrcCoord cx = src_coord(sx, x, w);
      const float wx = (cx.i0 == j ? cx.l0 : 0.f) + (cx.i0 + cx.step == j ? cx.l1 : 0.f);
      if (wx == 0.f) continue;
      const size_t px = (static_cast<size_t>(b) * H + y) * W + x;
      const long long t = target[px];
      if (t == ignore || t < 0 || t >= K) continue;
      const float l = lse[px];
      const float wgt = wy * wx;
#pragma unroll
      for (int k = 0; k < KP; ++k)
        if (k < K) {
          const float p = expf(bilerp_at(plane + static_cast<size_t>(k) * h * w, w, cy, cx) - l);
          acc[k] = fmaf(wgt, p - (k == static_cast<int>(t) ? 1.f : 0.f), acc[k]);
        }
    }
  }
#pragma unroll
  for (int k = 0; k < KP; ++k)
    if (k < K) grad_lr[(static_cast<size_t>(b) * K + k) * h * w + static_cast<size_t>(i) * w + j] = acc[k] * scale;
}

}  // namespace sl

extern "C" size_t sl_upsample_ce_ws_bytes(int B, int H, int W) {
  if (B < 1 || H < 1 || W < 1) return 0;
  // per-pixel log-sum-exp (fp32) + per-CTA partial sums (double) and counts (u64), 16-byte aligned sections
  const size_t lse = (static_cast<size_t>(B) * H * W * sizeof(float) + 15) / 16 * 16;
  return lse + static_cast<size_t>(sl::kNumSMs) * 8 * (sizeof(double) + sizeof(unsigned long long));
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
  unsigned long long* part_cnt = reinterpret_cast<unsigned long long*>(part_sum + sl::kNumSMs * 8);
  const long long total = static_cast<long long>(B) * H * W;
  long long blocks = (total + 255) / 256;
  if (blocks > sl::kNumSMs * 8) blocks = sl::kNumSMs * 8;
  const int grid = static_cast<int>(blocks);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float sy = sl::ac_scale(h, H), sx = sl::ac_scale(w, W);
#define SL_CE_FWD(KP) sl::ce_fwd_kernel<KP><<<grid, 256, 0, st>>>(logits_lr, B, K, h, w, H, W, sy, sx, target, \
                                                                   ignore_label, lse, part_sum, part_cnt)
  if (K <= 8) SL_CE_FWD(8); else if (K <= 12) SL_CE_FWD(12); else if (K <= 16) SL_CE_FWD(16); else SL_CE_FWD(32);
#undef SL_CE_FWD
  sl::ce_finish_kernel<<<1, 32, 0, st>>>(part_sum, part_cnt, grid, loss, n_valid);
  return SL_LAUNCH_RESULT();
}

extern "C" int sl_upsample_ce_bwd(const float* logits_lr, int B, int K, int h, int w, int H, int W,
                                  const long long* target, int ignore_label, const void* ws,
                                  const long long* n_valid, const float* grad_out, float* grad_logits_lr,
                                  void* stream) {
  SL_CHECK_PTR(logits_lr); SL_CHECK_PTR(target); SL_CHECK_PTR(ws); SL_CHECK_PTR(n_valid); SL_CHECK_PTR(grad_out);
  SL_CHECK_PTR(grad_logits_lr);
  SL_CHECK_ARG(B >= 1 && K >= 1 && K <= SL_MAX_CLASSES && h >= 1 && w >= 1 && H >= 1 && W >= 1);
  const float* lse = static_cast<const float*>(ws);
  const long long total = static_cast<long long>(B) * h * w;
  const int grid = static_cast<int>((total + 127) / 128);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float sy = sl::ac_scale(h, H), sx = sl::ac_scale(w, W);
#define SL_CE_BWD(KP) sl::ce_bwd_kernel<KP><<<grid, 128, 0, st>>>(logits_lr, B, K, h, w, H, W, sy, sx, target, \
                                                                   ignore_label, lse, n_valid, grad_out, grad_logits_lr)
  if (K <= 8) SL_CE_BWD(8); else if (K <= 12) SL_CE_BWD(12); else if (K <= 16) SL_CE_BWD(16); else SL_CE_BWD(32);
#undef SL_CE_BWD
  return SL_LAUNCH_RESULT();
}
