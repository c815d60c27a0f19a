// sl_fuse_argmax: fusemat.py:37-48 for one tile.
//   acc = mats[0]; acc += mats[1]; ...   (sequential fp32 adds in list order, fusemat.py:42-47)
//   fused = acc / len(fusion_list)       (IEEE fp32 division, fusemat.py:48)
//   pred = argmax over K (first maximum) -> uint8
// Pure streaming: M*K*HW*4 bytes in, HW bytes out -- the cleanest HBM-roofline kernel on the path.
// A thread owns 4 consecutive pixels (128-bit loads, uchar4 store); all M*K loads of a thread are
// independent, so memory-level parallelism comes from the class loop itself.
#include "common.cuh"

namespace sl {

struct FusePtrs { const float* m[SL_MAX_FUSE]; };

template <int M>  // 0 = runtime count
__global__ void __launch_bounds__(256) fuse_argmax_kernel(FusePtrs ptrs0, int Mrt, int K, long long HW4, float divisor,
                                                          uint8_t* __restrict__ pred0, float* __restrict__ fused0,
                                                          const uint8_t* __restrict__ label0, int ignore_label,
                                                          unsigned long long* __restrict__ cm, int T) {
  __shared__ unsigned int hist[SL_MAX_CLASSES * SL_MAX_CLASSES];
  const bool do_cm = cm != nullptr;
  if (do_cm) {
    for (int i = threadIdx.x; i < K * K; i += 256) hist[i] = 0u;
    __syncthreads();
  }
  const int Mn = M > 0 ? M : Mrt;
  const long long stride = static_cast<long long>(gridDim.x) * 256;
  const long long start = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  const long long iters = (HW4 + stride - 1) / stride;
  const size_t HW = static_cast<size_t>(HW4) * 4;
  // T tiles, tile-major stacks [T][K][HW]: every CTA takes its slice of every tile, so the per-CTA histogram is
  // flushed once per sweep instead of once per tile (the K*K global atomics per CTA were the limit of per-tile launches)
  for (int tile = 0; tile < T; ++tile) {
  FusePtrs ptrs;
#pragma unroll
  for (int m = 0; m < (M > 0 ? M : SL_MAX_FUSE); ++m)
    ptrs.m[m] = (m < Mn) ? ptrs0.m[m] + static_cast<size_t>(tile) * K * HW : nullptr;
  uint8_t* pred = pred0 + static_cast<size_t>(tile) * HW;
  float* fused = fused0 ? fused0 + static_cast<size_t>(tile) * K * HW : nullptr;
  const uint8_t* label = label0 ? label0 + static_cast<size_t>(tile) * HW : nullptr;
  for (long long it = 0; it < iters; ++it) {
    const long long g = start + it * stride;
    const bool active = g < HW4;
    int idx[4] = {0, 0, 0, 0};
    int lab[4] = {-1, -1, -1, -1};
    if (active) {
      float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll 2
      for (int k = 0; k < K; ++k) {
        const size_t off = static_cast<size_t>(k) * HW + static_cast<size_t>(g) * 4;
        float4 a = ld_stream_f4(ptrs.m[0] + off);
#pragma unroll
        for (int m = 1; m < (M > 0 ? M : SL_MAX_FUSE); ++m) {
          if (m >= Mn) break;
          const float4 t = ld_stream_f4(ptrs.m[m] + off);
          a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
        }
        a.x = __fdiv_rn(a.x, divisor); a.y = __fdiv_rn(a.y, divisor);
        a.z = __fdiv_rn(a.z, divisor); a.w = __fdiv_rn(a.w, divisor);
        if (fused) *reinterpret_cast<float4*>(fused + off) = a;
        const float v[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (v[j] > best[j] || (v[j] != v[j] && best[j] == best[j])) { best[j] = v[j]; idx[j] = k; }
      }
      *reinterpret_cast<uchar4*>(pred + static_cast<size_t>(g) * 4) = make_uchar4(idx[0], idx[1], idx[2], idx[3]);
      if (do_cm) {
        const uchar4 l4 = *reinterpret_cast<const uchar4*>(label + static_cast<size_t>(g) * 4);
        lab[0] = l4.x; lab[1] = l4.y; lab[2] = l4.z; lab[3] = l4.w;
      }
    }
    if (do_cm) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool valid = lab[j] >= 0 && lab[j] != ignore_label && lab[j] < K;
        hist_add_warp(hist, valid ? lab[j] * K + idx[j] : 0, valid);
      }
    }
  }
  }
  if (do_cm) {
    __syncthreads();
    for (int i = threadIdx.x; i < K * K; i += 256)
      if (hist[i]) atomicAdd(&cm[i], static_cast<unsigned long long>(hist[i]));
  }
}

}  // namespace sl

static int fuse_launch(const float* const* mats_host, int M, int T, int K, long long HW, int divisor, uint8_t* pred,
                       float* fused, const uint8_t* label, int ignore_label, long long* cm, void* stream) {
  SL_CHECK_PTR(mats_host); SL_CHECK_PTR(pred);
  SL_CHECK_ARG(T >= 1);
  SL_CHECK_ARG(M >= 1 && M <= SL_MAX_FUSE && K >= 1 && K <= SL_MAX_CLASSES && HW >= 4 && HW % 4 == 0 && divisor >= 1);
  if (cm) SL_CHECK_PTR(label);
  sl::FusePtrs p;
  for (int m = 0; m < SL_MAX_FUSE; ++m) p.m[m] = nullptr;
  for (int m = 0; m < M; ++m) {
    SL_CHECK_PTR(mats_host[m]);
    SL_CHECK_ALIGN(mats_host[m], 16);
    p.m[m] = mats_host[m];
  }
  SL_CHECK_ALIGN(pred, 4);
  if (fused) SL_CHECK_ALIGN(fused, 16);
  if (label) SL_CHECK_ALIGN(label, 4);
  const long long HW4 = HW / 4;
  long long blocks = (HW4 + 255) / 256;
  const long long cap = static_cast<long long>(sl::num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  auto* cmu = reinterpret_cast<unsigned long long*>(cm);
  const float d = static_cast<float>(divisor);
#define SL_FUSE_LAUNCH(MM) sl::fuse_argmax_kernel<MM><<<static_cast<int>(blocks), 256, 0, st>>>( \
      p, M, K, HW4, d, pred, fused, label, ignore_label, cmu, T)
  switch (M) {
    case 1: SL_FUSE_LAUNCH(1); break;
    case 2: SL_FUSE_LAUNCH(2); break;
    case 3: SL_FUSE_LAUNCH(3); break;
    case 4: SL_FUSE_LAUNCH(4); break;
    default: SL_FUSE_LAUNCH(0); break;
  }
#undef SL_FUSE_LAUNCH
  return SL_LAUNCH_RESULT();
}

extern "C" int sl_fuse_argmax(const float* const* mats_host, int M, int K, long long HW, int divisor, uint8_t* pred,
                              float* fused, const uint8_t* label, int ignore_label, long long* cm, void* stream) {
  return fuse_launch(mats_host, M, 1, K, HW, divisor, pred, fused, label, ignore_label, cm, stream);
}

// A sweep of T tiles in ONE launch: mats_host[m] is model m's stack for all tiles, [T,K,HW] (tile-major, as an eval sweep
// writes it); every tile is fused exactly as sl_fuse_argmax fuses one tile (same operations in the same order per pixel).
extern "C" int sl_fuse_argmax_tiles(const float* const* mats_host, int M, int T, int K, long long HW, int divisor,
                                    uint8_t* pred, float* fused, const uint8_t* label, int ignore_label, long long* cm,
                                    void* stream) {
  SL_CHECK_ARG(T >= 0);
  if (T == 0) return SL_OK;
  return fuse_launch(mats_host, M, T, K, HW, divisor, pred, fused, label, ignore_label, cm, stream);
}
