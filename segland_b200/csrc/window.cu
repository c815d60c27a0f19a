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

// One thread = 4 consecutive canvas columns of one row of one (b, k) plane.  The covering entries are walked as one
// flat index and loaded four at a time (independent loads, then added in entry order), so a thread keeps four
// requests in flight at ~40 registers.  VEC: ox % 4 == 0, wc % 4 == 0, w % 4 == 0 and 16-byte aligned pointers.
template <bool VEC>
__global__ void __launch_bounds__(256) window_accumulate_kernel(
    const float* __restrict__ crops, long long stride_e, long long stride_b, int K, int hc, int wc,
    const __grid_constant__ WinPlan pl, int h, int w, float* __restrict__ canvas, float* __restrict__ count) {
  const int wq = (w + 3) >> 2;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= wq * h) return;
  const int y = q / wq, x0 = (q - y * wq) * 4;
  const int p = blockIdx.y, b = p / K, k = p - b * K;
  int gy_lo, gy_hi;
  cover_range(pl.oy, pl.ny, hc, y, gy_lo, gy_hi);
  const int nyc = max(0, gy_hi - gy_lo + 1);
  const float* plane = crops + static_cast<size_t>(b) * stride_b + static_cast<size_t>(k) * hc * wc;
  float* dst = canvas + (static_cast<size_t>(p) * h + y) * w + x0;
  constexpr int NB = 4;                                          // entries loaded per batch
  if (VEC) {
    int gx_lo, gx_hi;
    cover_range(pl.ox, pl.nx, wc, x0, gx_lo, gx_hi);             // origins and widths are multiples of 4: one range
    const int nxc = max(0, gx_hi - gx_lo + 1), nxv = nxc * pl.V, n = nyc * nxv;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i0 = 0; i0 < n; i0 += NB) {
      float4 t[NB];
      bool fl[NB];
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const int i = min(i0 + j, n - 1);
        const int iy = i / nxv, r = i - iy * nxv, ix = r / pl.V, v = r - ix * pl.V;
        const int gy = gy_lo + iy, gx = gx_lo + ix, f = pl.flip[v];
        const int ry = y - pl.oy[gy], rx = x0 - pl.ox[gx];
        const int ys = (f & 2) ? hc - 1 - ry : ry, xs = (f & 1) ? wc - 4 - rx : rx;
        fl[j] = f & 1;
        t[j] = ld_stream_f4(plane + static_cast<size_t>((gy * pl.nx + gx) * pl.V + v) * stride_e + static_cast<size_t>(ys) * wc + xs);
      }
#pragma unroll
      for (int j = 0; j < NB; ++j)
        if (i0 + j < n) {
          acc.x += fl[j] ? t[j].w : t[j].x; acc.y += fl[j] ? t[j].z : t[j].y;
          acc.z += fl[j] ? t[j].y : t[j].z; acc.w += fl[j] ? t[j].x : t[j].w;
        }
    }
    const float cnt = static_cast<float>(n);
    __stcs(reinterpret_cast<float4*>(dst), make_float4(acc.x / cnt, acc.y / cnt, acc.z / cnt, acc.w / cnt));
    if (count != nullptr && p == 0) *reinterpret_cast<float4*>(count + static_cast<size_t>(y) * w + x0) = make_float4(cnt, cnt, cnt, cnt);
  } else {
#pragma unroll
    for (int jx = 0; jx < 4; ++jx) {
      const int x = x0 + jx;
      if (x >= w) break;
      int gx_lo, gx_hi;
      cover_range(pl.ox, pl.nx, wc, x, gx_lo, gx_hi);
      const int nxc = max(0, gx_hi - gx_lo + 1), nxv = nxc * pl.V, n = nyc * nxv;
      float acc = 0.f;
      for (int i0 = 0; i0 < n; i0 += NB) {
        float t[NB];
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          const int i = min(i0 + j, n - 1);
          const int iy = i / nxv, r = i - iy * nxv, ix = r / pl.V, v = r - ix * pl.V;
          const int gy = gy_lo + iy, gx = gx_lo + ix, f = pl.flip[v];
          const int ry = y - pl.oy[gy], rx = x - pl.ox[gx];
          const int ys = (f & 2) ? hc - 1 - ry : ry, xs = (f & 1) ? wc - 1 - rx : rx;
          t[j] = __ldg(plane + static_cast<size_t>((gy * pl.nx + gx) * pl.V + v) * stride_e + static_cast<size_t>(ys) * wc + xs);
        }
#pragma unroll
        for (int j = 0; j < NB; ++j) if (i0 + j < n) acc += t[j];
      }
      const float cnt = static_cast<float>(n);
      dst[jx] = acc / cnt;
      if (count != nullptr && p == 0) count[static_cast<size_t>(y) * w + x] = cnt;
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
  const int planes = B * K;
  SL_CHECK_ARG(planes <= 65535);
  const dim3 grid(static_cast<unsigned>((quads + 255) / 256), static_cast<unsigned>(planes));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (vec) sl::window_accumulate_kernel<true><<<grid, 256, 0, st>>>(crops, stride_e, stride_b, K, hc, wc, pl, h, w, canvas, count);
  else sl::window_accumulate_kernel<false><<<grid, 256, 0, st>>>(crops, stride_e, stride_b, K, hc, wc, pl, h, w, canvas, count);
  return SL_LAUNCH_RESULT();
}
