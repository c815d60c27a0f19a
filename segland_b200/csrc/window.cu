// sl_window_accumulate: sliding-window / flip aggregation of per-crop logits at feature resolution.
//
// spec: this repo.  BASELINE.json's north_star names "engine.py's sliding-window/flip aggregation", but the reference
// has none: engine.py:23-143 is argparse + DDP helpers and eval_base.py:162-170 / eval_ft.py:162-172 run one whole
// tile per forward.  The operator is therefore defined here, as the composition every sliding-window
// evaluator uses (crop grid with a last window pulled back to the border, overlap-count normalisation, flipped views
// un-flipped before the sum), and tested against that composition written with PyTorch primitives.
//
//   canvas[b,k,y,x] = ( sum over entries e = (gy, gx, v) covering (y, x), in index order,
//                        of unflip_v(crops[e,b,k])[y - oy[gy], x - ox[gx]] ) / count[y,x]
//
// It is a gather: every crop element is read exactly once and every canvas element written exactly once, there are
// no atomics and the summation order is fixed, so the result is bit-reproducible and equal, bit for bit, to
// `canvas[..., oy:oy+hc, ox:ox+wc] += unflip(crop)` applied entry by entry followed by `canvas / count`.
// HBM-bound: algorithmic bytes = E*B*K*hc*wc*4 read + B*K*h*w*4 written.
#include "common.cuh"

namespace sl {

struct WinPlan {
  int ny, nx, V;
  int oy[SL_MAX_WINDOWS_1D];
  int ox[SL_MAX_WINDOWS_1D];
  int flip[SL_MAX_VIEWS];
};

// windows whose origin o satisfies o <= p < o + len: origins ascend, so the range is contiguous
__device__ __forceinline__ void cover_range(const int* o, int n, int len, int p, int& lo, int& hi) {
  lo = n; hi = -1;
  for (int g = 0; g < n; ++g)
    if (o[g] <= p && p < o[g] + len) { if (g < lo) lo = g; hi = g; }
}

// overlap count -> exact division: powers of two multiply by the (exact) reciprocal, bit-identical to the IEEE
// division and ~10 instructions cheaper; other counts divide
__device__ __forceinline__ float div_count(float a, float cnt, float inv, bool pow2) { return pow2 ? a * inv : a / cnt; }

// One thread = 4 consecutive canvas columns of one row, for PC = 4 consecutive class planes of one image: the covering
// window ranges are worked out once, the (gy, gx, view) loops are nested (no divisions, mostly warp-uniform) and every
// entry issues four independent 16-byte loads, one per plane; per plane the entries are added in index order.
// VEC: ox % 4 == 0, wc % 4 == 0, w % 4 == 0 and 16-byte aligned pointers.
template <bool VEC>
__global__ void __launch_bounds__(256) window_accumulate_kernel(
    const float* __restrict__ crops, long long stride_e, long long stride_b, int K, int hc, int wc,
    const __grid_constant__ WinPlan pl, int h, int w, float* __restrict__ canvas, float* __restrict__ count) {
  constexpr int PC = 4;
  const int wq = (w + 3) >> 2;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= wq * h) return;
  const int y = q / wq, x0 = (q - y * wq) * 4;
  const int kchunks = (K + PC - 1) / PC;
  const int b = blockIdx.y / kchunks, k0 = (blockIdx.y - b * kchunks) * PC;
  const int nk = min(PC, K - k0);
  int gy_lo, gy_hi;
  cover_range(pl.oy, pl.ny, hc, y, gy_lo, gy_hi);
  const int chw = hc * wc;
  const float* img = crops + static_cast<size_t>(b) * stride_b + static_cast<size_t>(k0) * chw;
  float* dst = canvas + ((static_cast<size_t>(b) * K + k0) * h + y) * w + x0;
  const size_t plane_out = static_cast<size_t>(h) * w;
  const bool first_plane = blockIdx.y == 0;
  if (VEC) {
    int gx_lo, gx_hi;
    cover_range(pl.ox, pl.nx, wc, x0, gx_lo, gx_hi);             // origins and widths are multiples of 4: one range
    float4 acc[PC];
#pragma unroll
    for (int i = 0; i < PC; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int gy = gy_lo; gy <= gy_hi; ++gy) {
      const int ry = y - pl.oy[gy];
      for (int gx = gx_lo; gx <= gx_hi; ++gx) {
        const int rx = x0 - pl.ox[gx];
        const float* ent = img + static_cast<size_t>((gy * pl.nx + gx) * pl.V) * stride_e;
        for (int v = 0; v < pl.V; ++v, ent += stride_e) {
          const int f = pl.flip[v];
          const int ys = (f & 2) ? hc - 1 - ry : ry, xs = (f & 1) ? wc - 4 - rx : rx;
          const float* src = ent + ys * wc + xs;
          float4 t[PC];
#pragma unroll
          for (int i = 0; i < PC; ++i) t[i] = ld_stream_f4(src + min(i, nk - 1) * chw);
          if (f & 1) {
#pragma unroll
            for (int i = 0; i < PC; ++i) { acc[i].x += t[i].w; acc[i].y += t[i].z; acc[i].z += t[i].y; acc[i].w += t[i].x; }
          } else {
#pragma unroll
            for (int i = 0; i < PC; ++i) { acc[i].x += t[i].x; acc[i].y += t[i].y; acc[i].z += t[i].z; acc[i].w += t[i].w; }
          }
        }
      }
    }
    const int n = max(0, gy_hi - gy_lo + 1) * max(0, gx_hi - gx_lo + 1) * pl.V;
    const float cnt = static_cast<float>(n), inv = 1.f / cnt;
    const bool pow2 = (n & (n - 1)) == 0;
#pragma unroll
    for (int i = 0; i < PC; ++i)
      if (i < nk)
        __stcs(reinterpret_cast<float4*>(dst + i * plane_out),
               make_float4(div_count(acc[i].x, cnt, inv, pow2), div_count(acc[i].y, cnt, inv, pow2),
                           div_count(acc[i].z, cnt, inv, pow2), div_count(acc[i].w, cnt, inv, pow2)));
    if (count != nullptr && first_plane) *reinterpret_cast<float4*>(count + static_cast<size_t>(y) * w + x0) = make_float4(cnt, cnt, cnt, cnt);
  } else {
#pragma unroll
    for (int jx = 0; jx < 4; ++jx) {
      const int x = x0 + jx;
      if (x >= w) break;
      int gx_lo, gx_hi;
      cover_range(pl.ox, pl.nx, wc, x, gx_lo, gx_hi);
      float acc[PC];
#pragma unroll
      for (int i = 0; i < PC; ++i) acc[i] = 0.f;
      for (int gy = gy_lo; gy <= gy_hi; ++gy) {
        const int ry = y - pl.oy[gy];
        for (int gx = gx_lo; gx <= gx_hi; ++gx) {
          const int rx = x - pl.ox[gx];
          const float* ent = img + static_cast<size_t>((gy * pl.nx + gx) * pl.V) * stride_e;
          for (int v = 0; v < pl.V; ++v, ent += stride_e) {
            const int f = pl.flip[v];
            const int ys = (f & 2) ? hc - 1 - ry : ry, xs = (f & 1) ? wc - 1 - rx : rx;
            const float* src = ent + ys * wc + xs;
            float t[PC];
#pragma unroll
            for (int i = 0; i < PC; ++i) t[i] = __ldg(src + min(i, nk - 1) * chw);
#pragma unroll
            for (int i = 0; i < PC; ++i) acc[i] += t[i];
          }
        }
      }
      const int n = max(0, gy_hi - gy_lo + 1) * max(0, gx_hi - gx_lo + 1) * pl.V;
      const float cnt = static_cast<float>(n);
#pragma unroll
      for (int i = 0; i < PC; ++i)
        if (i < nk) dst[i * plane_out + jx] = acc[i] / cnt;
      if (count != nullptr && first_plane) count[static_cast<size_t>(y) * w + x] = cnt;
    }
  }
}

}  // namespace sl

extern "C" int sl_window_accumulate(const float* crops, long long stride_e, long long stride_b, int B, int K, int hc,
                                    int wc, const int* oy_host, int ny, const int* ox_host, int nx,
                                    const int* flip_host, int V, int h, int w, float* canvas, float* count,
                                    void* stream) {
  SL_CHECK_PTR(crops); SL_CHECK_PTR(canvas); SL_CHECK_PTR(oy_host); SL_CHECK_PTR(ox_host); SL_CHECK_PTR(flip_host);
  SL_CHECK_ARG(B >= 1 && K >= 1 && hc >= 1 && wc >= 1 && h >= hc && w >= wc);
  SL_CHECK_ARG(ny >= 1 && ny <= SL_MAX_WINDOWS_1D && nx >= 1 && nx <= SL_MAX_WINDOWS_1D && V >= 1 && V <= SL_MAX_VIEWS);
  SL_CHECK_ARG(stride_e >= 0 && stride_b >= 0 && static_cast<long long>(B) * K < (1ll << 24));
  sl::WinPlan pl;
  pl.ny = ny; pl.nx = nx; pl.V = V;
  bool vec = (w % 4 == 0) && (wc % 4 == 0) && (stride_e % 4 == 0) && (stride_b % 4 == 0) &&
             (static_cast<long long>(hc) * wc % 4 == 0) &&
             reinterpret_cast<uintptr_t>(crops) % 16 == 0 && reinterpret_cast<uintptr_t>(canvas) % 16 == 0 &&
             (count == nullptr || reinterpret_cast<uintptr_t>(count) % 16 == 0);
  // every canvas row / column must be covered: origins ascend from 0, consecutive windows touch or overlap, and the
  // last one ends at the border (the usual "pull the last window back" plan)
  for (int g = 0; g < SL_MAX_WINDOWS_1D; ++g) { pl.oy[g] = g < ny ? oy_host[g] : 0; pl.ox[g] = g < nx ? ox_host[g] : 0; }
  SL_CHECK_ARG(pl.oy[0] == 0 && pl.ox[0] == 0 && pl.oy[ny - 1] + hc == h && pl.ox[nx - 1] + wc == w);
  for (int g = 1; g < ny; ++g) SL_CHECK_ARG(pl.oy[g] >= pl.oy[g - 1] && pl.oy[g] <= pl.oy[g - 1] + hc);
  for (int g = 1; g < nx; ++g) SL_CHECK_ARG(pl.ox[g] >= pl.ox[g - 1] && pl.ox[g] <= pl.ox[g - 1] + wc);
  for (int g = 0; g < nx; ++g) vec = vec && (pl.ox[g] % 4 == 0);
  for (int v = 0; v < SL_MAX_VIEWS; ++v) {
    pl.flip[v] = v < V ? flip_host[v] : 0;
    SL_CHECK_ARG(pl.flip[v] >= 0 && pl.flip[v] <= 3);
  }
  const long long quads = static_cast<long long>(h) * ((w + 3) / 4);
  SL_CHECK_ARG(quads < (1ll << 31));
  const long long groups = static_cast<long long>(B) * ((K + 3) / 4);       // (image, chunk of 4 class planes)
  SL_CHECK_ARG(groups <= 65535 && static_cast<long long>(hc) * wc * K < (1ll << 31));
  const dim3 grid(static_cast<unsigned>((quads + 255) / 256), static_cast<unsigned>(groups));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (vec) sl::window_accumulate_kernel<true><<<grid, 256, 0, st>>>(crops, stride_e, stride_b, K, hc, wc, pl, h, w, canvas, count);
  else sl::window_accumulate_kernel<false><<<grid, 256, 0, st>>>(crops, stride_e, stride_b, K, hc, wc, pl, h, w, canvas, count);
  return SL_LAUNCH_RESULT();
}
