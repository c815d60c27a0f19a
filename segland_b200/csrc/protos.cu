// Prototype-side operators:
//   sl_map_proto   masked_average_pooling, networks/pspnet.py:7-15
//   sl_orth_loss   proto_sim (networks/pspnet_pop.py:185-186, :234-239) + OrthLoss.get_orth_loss
//                  (loss/criterion.py:37-43), forward and gradient w.r.t. the trainable rows.
#include "common.cuh"

namespace sl {

// ---- MAP step 1: bilinear align_corners=True resample of the mask to feature resolution.  grid
// (chunks of 1024 low-res pixels, B); every CTA also leaves the sum of its chunk, and step 2 adds the
// chunk sums of an image in index order -> fixed summation order, bit-reproducible.
constexpr int MAP_CHUNK = 1024;
__global__ void __launch_bounds__(256) map_mask_kernel(const float* __restrict__ mask, int h, int w, int H, int W,
                                                       float sy, float sx, float* __restrict__ mask_lr,
                                                       float* __restrict__ partial, int n_chunks) {
  __shared__ float red[8];
  const int b = blockIdx.y;
  const float* src = mask + static_cast<size_t>(b) * H * W;
  float* dst = mask_lr + static_cast<size_t>(b) * h * w;
  float acc = 0.f;
  const int i_end = min(h * w, (static_cast<int>(blockIdx.x) + 1) * MAP_CHUNK);
  for (int i = blockIdx.x * MAP_CHUNK + threadIdx.x; i < i_end; i += 256) {
    const int y = i / w, x = i - y * w;
    const SrcCoord cy = src_coord(sy, y, H), cx = src_coord(sx, x, W);
    const float* r0 = src + static_cast<size_t>(cy.i0) * W + cx.i0;
    const float* r1 = r0 + static_cast<size_t>(cy.step) * W;
    const float v = cy.l0 * (cx.l0 * __ldg(r0) + cx.l1 * __ldg(r0 + cx.step)) +
                    cy.l1 * (cx.l0 * __ldg(r1) + cx.l1 * __ldg(r1 + cx.step));
    dst[i] = v;
    acc += v;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i];
    partial[static_cast<size_t>(b) * n_chunks + blockIdx.x] = t;
  }
}

// ---- MAP step 2: per (image, channel) masked sum.  A warp owns MAP_CH channels at a time so each
// mask value loaded from L1 is reused MAP_CH times; features stream once with 128-bit loads.
constexpr int MAP_CH = 4;
// grid (ceil(C / 32), B, splits): the pixel range is cut into `splits` pieces of split_len pixels so that enough
// warps are in flight to cover the HBM latency (B * C / 4 warps alone left the SMs a quarter full); every piece
// writes its unnormalised sums to part[b][split][c] and map_finish_kernel adds them in index order.
constexpr int MAP_MAX_SPLITS = 8;
__global__ void __launch_bounds__(256) map_reduce_kernel(const uint16_t* __restrict__ feat, int C, int N, int split_len,
                                                         const float* __restrict__ mask_lr, float* __restrict__ part) {
  const int b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = (blockIdx.x * 8 + warp) * MAP_CH;
  if (c0 >= C) return;
  const float* m = mask_lr + static_cast<size_t>(b) * N;
  const uint16_t* f = feat + (static_cast<size_t>(b) * C + c0) * N;
  float acc[MAP_CH];
#pragma unroll
  for (int j = 0; j < MAP_CH; ++j) acc[j] = 0.f;
  // two 8-pixel groups per lane per iteration: 2 x MAP_CH independent 128-bit loads in flight
  auto accumulate = [&](const uint4 (&v)[MAP_CH], const float4& ma, const float4& mb) {
#pragma unroll
    for (int j = 0; j < MAP_CH; ++j) {
      float a = acc[j];
      a = fmaf(bf16lo(v[j].x), ma.x, a); a = fmaf(bf16hi(v[j].x), ma.y, a);
      a = fmaf(bf16lo(v[j].y), ma.z, a); a = fmaf(bf16hi(v[j].y), ma.w, a);
      a = fmaf(bf16lo(v[j].z), mb.x, a); a = fmaf(bf16hi(v[j].z), mb.y, a);
      a = fmaf(bf16lo(v[j].w), mb.z, a); a = fmaf(bf16hi(v[j].w), mb.w, a);
      acc[j] = a;
    }
  };
  const int n_end = min(N, (static_cast<int>(blockIdx.z) + 1) * split_len);
  int n = blockIdx.z * split_len + lane * 8;
  for (; n + 256 < n_end; n += 512) {
    uint4 v0[MAP_CH], v1[MAP_CH];
#pragma unroll
    for (int j = 0; j < MAP_CH; ++j) {
      const bool ok = c0 + j < C;
      v0[j] = ok ? ld_stream_u4(f + static_cast<size_t>(j) * N + n) : make_uint4(0, 0, 0, 0);
      v1[j] = ok ? ld_stream_u4(f + static_cast<size_t>(j) * N + n + 256) : make_uint4(0, 0, 0, 0);
    }
    const float4 ma0 = __ldg(reinterpret_cast<const float4*>(m + n)), mb0 = __ldg(reinterpret_cast<const float4*>(m + n + 4));
    const float4 ma1 = __ldg(reinterpret_cast<const float4*>(m + n + 256)), mb1 = __ldg(reinterpret_cast<const float4*>(m + n + 260));
    accumulate(v0, ma0, mb0);
    accumulate(v1, ma1, mb1);
  }
  for (; n < n_end; n += 256) {
    uint4 v[MAP_CH];
#pragma unroll
    for (int j = 0; j < MAP_CH; ++j)
      v[j] = (c0 + j < C) ? ld_stream_u4(f + static_cast<size_t>(j) * N + n) : make_uint4(0, 0, 0, 0);
    const float4 ma = __ldg(reinterpret_cast<const float4*>(m + n)), mb = __ldg(reinterpret_cast<const float4*>(m + n + 4));
    accumulate(v, ma, mb);
  }
#pragma unroll
  for (int j = 0; j < MAP_CH; ++j) {
    const float t = warp_sum(acc[j]);
    if (lane == 0 && c0 + j < C) part[(static_cast<size_t>(b) * gridDim.z + blockIdx.z) * C + c0 + j] = t;
  }
}

// per_image[b][c] = sum_split part / (sum of the image's mask + 1e-5): one thread per (image, channel) -- the first
// version walked all B images in one thread per channel, a serial chain of ~500 dependent loads that took 34 us of the
// 97 us the whole operator needed for 20 support tiles.  Sums in index order: bit-reproducible.
__global__ void map_finish_kernel(const float* __restrict__ part, int splits, const float* __restrict__ partial,
                                  int n_chunks, int C, float* __restrict__ per_image) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (c >= C) return;
  float msum = 0.f;
  for (int i = 0; i < n_chunks; ++i) msum += partial[static_cast<size_t>(b) * n_chunks + i];
  float sv = 0.f;
  for (int z = 0; z < splits; ++z) sv += part[(static_cast<size_t>(b) * splits + z) * C + c];
  per_image[static_cast<size_t>(b) * C + c] = sv / (msum + 1e-5f);
}
// proto[c] = mean over images, added in image order
__global__ void map_mean_kernel(const float* __restrict__ per_image, int B, int C, float* __restrict__ proto) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float t = 0.f;
  for (int b = 0; b < B; ++b) t += per_image[static_cast<size_t>(b) * C + c];
  proto[c] = t / static_cast<float>(B);
}

// ---- orthogonal-prototype loss, one CTA.  e = [normalize(rows); normalize(others)]  [Kr+Ko, C]
//   sim[i][j] = e_i . e_j  (i < Kr),  loss = mean_{j>i} |sim[i][j]|,
//   d loss / d rows[i] = (I - e_i e_i^T) g_i / max(||rows[i]||, eps),
//   g_i = (1/M) ( sum_{j>i} sgn(sim[i][j]) e_j + sum_{j<i} sgn(sim[j][i]) e_j ).
__global__ void __launch_bounds__(256) orth_loss_kernel(const float* __restrict__ rows, int Kr,
                                                        const float* __restrict__ others, int Ko, int C,
                                                        float* __restrict__ proto_sim, float* __restrict__ loss,
                                                        float* __restrict__ grad_rows) {
  extern __shared__ float sm[];
  const int Kt = Kr + Ko;
  float* e = sm;                 // [Kt][C]
  float* nrm = sm + Kt * C;      // [Kt]
  float* sim = nrm + Kt;         // [Kr][Kt]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < Kt; r += 8) {
    const float* src = r < Kr ? rows + static_cast<size_t>(r) * C : others + static_cast<size_t>(r - Kr) * C;
    float ss = 0.f;
    for (int i = lane; i < C; i += 32) { const float v = src[i]; ss = fmaf(v, v, ss); }
    ss = warp_sum(ss);
    const float n = fmaxf(sqrtf(ss), 1e-12f);
    if (lane == 0) nrm[r] = n;
    for (int i = lane; i < C; i += 32) e[r * C + i] = src[i] / n;
  }
  __syncthreads();
  for (int p = warp; p < Kr * Kt; p += 8) {
    const int i = p / Kt, j = p - i * Kt;
    float acc = 0.f;
    for (int c = lane; c < C; c += 32) acc = fmaf(e[i * C + c], e[j * C + c], acc);
    acc = warp_sum(acc);
    if (lane == 0) { sim[p] = acc; if (proto_sim) proto_sim[p] = acc; }
  }
  __syncthreads();
  int M = 0;
  for (int i = 0; i < Kr; ++i) M += Kt - 1 - i;
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < Kr; ++i)
      for (int j = i + 1; j < Kt; ++j) t += fabsf(sim[i * Kt + j]);
    loss[0] = M > 0 ? t / static_cast<float>(M) : nanf("");
  }
  if (grad_rows == nullptr) return;
  const float invM = M > 0 ? 1.f / static_cast<float>(M) : 0.f;
  auto sgn = [](float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); };
  for (int i = warp; i < Kr; i += 8) {
    // g . e_i first (projection term), then the projected gradient
    float dotp = 0.f;
    for (int c = lane; c < C; c += 32) {
      float g = 0.f;
      for (int j = i + 1; j < Kt; ++j) g = fmaf(sgn(sim[i * Kt + j]), e[j * C + c], g);
      for (int j = 0; j < i; ++j) g = fmaf(sgn(sim[j * Kt + i]), e[j * C + c], g);
      dotp = fmaf(g, e[i * C + c], dotp);
    }
    dotp = warp_sum(dotp);
    for (int c = lane; c < C; c += 32) {
      float g = 0.f;
      for (int j = i + 1; j < Kt; ++j) g = fmaf(sgn(sim[i * Kt + j]), e[j * C + c], g);
      for (int j = 0; j < i; ++j) g = fmaf(sgn(sim[j * Kt + i]), e[j * C + c], g);
      grad_rows[static_cast<size_t>(i) * C + c] = invM * (g - dotp * e[i * C + c]) / nrm[i];
    }
  }
}

// OrthLoss.get_orth_loss (loss/criterion.py:37-43) on a proto_sim matrix the model already built:
// mean |sim[i][j]| over j > i (the entries torch.triu(ones_like(sim), diagonal=1) == 1 selects), summed in row-major
// order by one thread (<= 32 x 64 values), plus d loss / d sim = sign(sim) / M on those entries.
__global__ void orth_from_sim_kernel(const float* __restrict__ sim, int Kr, int Kc, float* __restrict__ loss,
                                     float* __restrict__ grad_sim) {
  int M = 0;
  for (int i = 0; i < Kr; ++i) M += max(0, Kc - 1 - i);
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < Kr; ++i)
      for (int j = i + 1; j < Kc; ++j) t += fabsf(sim[i * Kc + j]);
    loss[0] = M > 0 ? t / static_cast<float>(M) : nanf("");   // torch: mean of an empty selection is nan
  }
  if (grad_sim == nullptr) return;
  const float invM = M > 0 ? 1.f / static_cast<float>(M) : 0.f;
  for (int p = threadIdx.x; p < Kr * Kc; p += blockDim.x) {
    const int i = p / Kc, j = p - i * Kc;
    const float v = sim[p];
    grad_sim[p] = j > i ? (v > 0.f ? invM : (v < 0.f ? -invM : 0.f)) : 0.f;
  }
}

}  // namespace sl

extern "C" int sl_orth_from_sim(const float* proto_sim, int Kr, int Kc, float* loss, float* grad_sim, void* stream) {
  SL_CHECK_PTR(proto_sim); SL_CHECK_PTR(loss);
  SL_CHECK_ARG(Kr >= 1 && Kr <= SL_MAX_CLASSES && Kc >= 1 && Kc <= 2 * SL_MAX_CLASSES);
  sl::orth_from_sim_kernel<<<1, 128, 0, static_cast<cudaStream_t>(stream)>>>(proto_sim, Kr, Kc, loss, grad_sim);
  return SL_LAUNCH_RESULT();
}

extern "C" int sl_map_proto(const uint16_t* feat, const float* mask, int B, int C, int h, int w, int H, int W,
                            float* mask_lr_ws, float* per_image, float* proto, void* stream) {
  SL_CHECK_PTR(feat); SL_CHECK_PTR(mask); SL_CHECK_PTR(mask_lr_ws); SL_CHECK_PTR(per_image); SL_CHECK_PTR(proto);
  SL_CHECK_ARG(B >= 1 && B <= 65535 && C >= 1 && h >= 1 && w >= 1 && H >= 1 && W >= 1);
  const long long N = static_cast<long long>(h) * w;
  SL_CHECK_ARG(N % 8 == 0 && N < (1ll << 30));
  SL_CHECK_ALIGN(feat, 16); SL_CHECK_ALIGN(mask_lr_ws, 16);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int n_chunks = static_cast<int>((N + sl::MAP_CHUNK - 1) / sl::MAP_CHUNK);
  float* partial = mask_lr_ws + static_cast<size_t>(B) * N;   // [B][n_chunks] chunk sums follow the [B,N] map
  sl::map_mask_kernel<<<dim3(n_chunks, B), 256, 0, st>>>(mask, h, w, H, W, sl::ac_scale(H, h), sl::ac_scale(W, w),
                                                        mask_lr_ws, partial, n_chunks);
  // enough pixel splits to put ~48 warps on every SM, each at least 2048 pixels long
  const long long warps = static_cast<long long>(B) * ((C + sl::MAP_CH - 1) / sl::MAP_CH);
  long long splits = (static_cast<long long>(sl::num_sms()) * 48 + warps - 1) / warps;
  if (splits > sl::MAP_MAX_SPLITS) splits = sl::MAP_MAX_SPLITS;
  if (splits > (N + 2047) / 2048) splits = (N + 2047) / 2048;
  if (splits < 1) splits = 1;
  const int split_len = static_cast<int>(((N + splits - 1) / splits + 511) / 512 * 512);
  splits = (N + split_len - 1) / split_len;
  float* part = partial + static_cast<size_t>(B) * n_chunks;     // [B][splits][C]
  dim3 grid((C + 8 * sl::MAP_CH - 1) / (8 * sl::MAP_CH), B, static_cast<unsigned>(splits));
  sl::map_reduce_kernel<<<grid, 256, 0, st>>>(feat, C, static_cast<int>(N), split_len, mask_lr_ws, part);
  sl::map_finish_kernel<<<dim3((C + 127) / 128, B), 128, 0, st>>>(part, static_cast<int>(splits), partial, n_chunks, C,
                                                                  per_image);
  sl::map_mean_kernel<<<(C + 127) / 128, 128, 0, st>>>(per_image, B, C, proto);
  return SL_LAUNCH_RESULT();
}

extern "C" int sl_orth_loss(const float* rows, int Kr, const float* others, int Ko, int C, float* proto_sim,
                            float* loss, float* grad_rows, void* stream) {
  SL_CHECK_PTR(rows); SL_CHECK_PTR(loss);
  SL_CHECK_ARG(Kr >= 1 && Ko >= 0 && Kr + Ko <= SL_MAX_CLASSES && C >= 1 && C <= 1024);
  if (Ko > 0) SL_CHECK_PTR(others);
  const int Kt = Kr + Ko;
  const size_t smem = (static_cast<size_t>(Kt) * C + Kt + static_cast<size_t>(Kr) * Kt) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(sl::orth_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  sl::orth_loss_kernel<<<1, 256, smem, static_cast<cudaStream_t>(stream)>>>(rows, Kr, others, Ko, C, proto_sim, loss,
                                                                           grad_rows);
  return SL_LAUNCH_RESULT();
}
