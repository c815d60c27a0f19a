// Backward of the POP head for training: forward_novel / forward_base in train
// mode (networks/pspnet_pop.py:191-219, :161-189) differentiate through orthogonal_decompose and the two
// bias-free MLPs.  With the forward collapse
//     fg:  y_k = p_k >= 0 ? alpha_k p_k : -beta_k p_k,   p_k = s_hat_k . q
//     bg:  y_0 = w3 . relu(W2 relu(W1' q)),              W1' = W1 (I - S_hat^T S_hat)
// the per-pixel part of the backward (everything whose cost grows with the pixel count) is
//     d_alpha_k = sum [p_k>=0] g_k p_k      d_beta_k = -sum [p_k<0] g_k p_k      gp_k = g_k (p_k>=0 ? alpha_k : -beta_k)
//     d_s_hat_k = sum gp_k q                (through p_k = s_hat_k . q only; the W1' term comes back through dW1')
//     h1 = relu(W1' q), z2 = W2 h1, dz2 = g_0 w3 [z2>0], dw3 = sum g_0 relu(z2)
//     dW2 = sum dz2 h1^T,  dz1 = (W2^T dz2) [h1>0],  dW1' = sum dz1 q^T
//     d_q = W1'^T dz1 + sum_k gp_k s_hat_k  (optional: only when the decoder trains)
// and the parameter-side chain (normalisation, alpha/beta = MLP(+-s_hat), the fold of W1) is O(K C^2) work the
// host layer differentiates on [K,C]/[C,C] tensors.  This file is the exact-fp32 CUDA-core implementation:
// register-tiled SGEMMs (128x128x16 tiles, 8x8 per thread) with the activations / masks / reductions fused
// into their epilogues.  Activations are recomputed, not saved by the forward.
#include "common.cuh"

namespace sl {
namespace bwd {

constexpr int BM = 128, BN = 128, BK = 16, THREADS = 256, PAD = 4;

// ------------------------------------------------------------------------------------------ operand accessors
struct RowMajor {                      // element (r, c) of a row-major fp32 matrix
  const float* p; long long ld;
  __device__ __forceinline__ float operator()(int r, int c) const { return p[static_cast<long long>(r) * ld + c]; }
};
struct ColMajor {                      // element (r, c) of the transpose of a row-major matrix
  const float* p; long long ld;
  __device__ __forceinline__ float operator()(int r, int c) const { return p[static_cast<long long>(c) * ld + r]; }
};
struct FeatPxCh {                      // (pixel, channel) of bf16 features [B][C][N]; pixel = b * N + n
  const uint16_t* p; int C, N;
  __device__ __forceinline__ float operator()(int px, int c) const {
    const int b = px / N, n = px - b * N;
    return bf16_bits_to_f32(p[(static_cast<long long>(b) * C + c) * N + n]);
  }
};
struct FeatChPx {                      // (channel, pixel) view of the same tensor
  FeatPxCh f;
  __device__ __forceinline__ float operator()(int c, int px) const { return f(px, c); }
};
struct ProjKPx {                       // (class k, pixel) of a [B][K][N] fp32 tensor
  const float* p; int K, N;
  __device__ __forceinline__ float operator()(int k, int px) const {
    const int b = px / N, n = px - b * N;
    return p[(static_cast<long long>(b) * K + k) * N + n];
  }
};

// ------------------------------------------------------------------------------------------ SGEMM
// out(m, n) = sum_{k in this CTA's k-chunk} A(m, k) * B(k, n), handed element-wise to `epi(m, n, value)`.
// A_MC / B_NC say which index is contiguous in memory (m for A, n for B; otherwise k) so that the global loads
// of a warp coalesce.  grid = (ceil(N/128), ceil(M/128), k-splits).
template <bool A_MC, bool B_NC, class LA, class LB, class Epi>
__global__ void __launch_bounds__(THREADS) sgemm_kernel(int M, int N, int K, int k_chunk, LA la, LB lb, Epi epi) {
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int k_begin = blockIdx.z * k_chunk;
  const int k_end = min(K, k_begin + k_chunk);
  const int tiles = (k_end - k_begin + BK - 1) / BK;

  float ra[8], rb[8];
  auto load_regs = [&](int t) {
    const int kt = k_begin + t * BK;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int e = tid + i * THREADS;
      const int ml = A_MC ? (e % BM) : (e / BK), kl = A_MC ? (e / BM) : (e % BK);
      const int m = m0 + ml, k = kt + kl;
      ra[i] = (m < M && k < k_end) ? la(m, k) : 0.f;
      const int nl = B_NC ? (e % BN) : (e / BK), kb = B_NC ? (e / BN) : (e % BK);
      const int n = n0 + nl, k2 = kt + kb;
      rb[i] = (n < N && k2 < k_end) ? lb(k2, n) : 0.f;
    }
  };
  auto store_smem = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int e = tid + i * THREADS;
      const int ml = A_MC ? (e % BM) : (e / BK), kl = A_MC ? (e / BM) : (e % BK);
      As[buf][kl][ml] = ra[i];
      const int nl = B_NC ? (e % BN) : (e / BK), kb = B_NC ? (e / BN) : (e % BK);
      Bs[buf][kb][nl] = rb[i];
    }
  };

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  if (tiles > 0) {
    load_regs(0);
    store_smem(0);
  }
  __syncthreads();
  int buf = 0;
  for (int t = 0; t < tiles; ++t) {
    if (t + 1 < tiles) load_regs(t + 1);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (t + 1 < tiles) store_smem(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4);
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + j - 4);
      if (n < N) epi(m, n, acc[i][j]);
    }
  }
}

template <bool A_MC, bool B_NC, class LA, class LB, class Epi>
static void sgemm(int M, int N, int K, int splits, LA la, LB lb, Epi epi, cudaStream_t st) {
  int k_chunk = (K + splits - 1) / splits;
  k_chunk = (k_chunk + BK - 1) / BK * BK;
  splits = (K + k_chunk - 1) / k_chunk;
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, splits);
  sgemm_kernel<A_MC, B_NC, LA, LB, Epi><<<grid, THREADS, 0, st>>>(M, N, K, k_chunk, la, lb, epi);
}

// ------------------------------------------------------------------------------------------ epilogues
struct EpiRelu {                       // h1[px][c] = relu(v)
  float* out; long long ld;
  __device__ __forceinline__ void operator()(int m, int n, float v) const { out[static_cast<long long>(m) * ld + n] = fmaxf(v, 0.f); }
};
struct EpiLayer2 {                     // z2 -> dz2 = g0 w3 [z2 > 0]; h2 = relu(z2) kept for dw3
  float* dz2; float* h2; const float* w3; const float* g; int C, N, Ktot, ch;
  __device__ __forceinline__ void operator()(int m, int n, float v) const {
    const int b = m / N, px = m - b * N;
    const float g0 = g[(static_cast<long long>(b) * Ktot + ch) * N + px];
    const long long o = static_cast<long long>(m) * C + n;
    dz2[o] = v > 0.f ? g0 * w3[n] : 0.f;
    h2[o] = fmaxf(v, 0.f);
  }
};
struct EpiMask {                       // dz1 = v [h1 > 0]  (in place over the h2 buffer)
  float* out; const float* h1; long long ld;
  __device__ __forceinline__ void operator()(int m, int n, float v) const {
    const long long o = static_cast<long long>(m) * ld + n;
    out[o] = h1[o] > 0.f ? v : 0.f;
  }
};
struct EpiAtomic {                     // split-k accumulation into a zeroed fp32 matrix
  float* out; long long ld;
  __device__ __forceinline__ void operator()(int m, int n, float v) const { atomicAdd(out + static_cast<long long>(m) * ld + n, v); }
};
struct EpiDFeat {                      // d_q[b][c][n] = v + sum_k gp_k s_hat_k[c]   (m = channel, n = pixel)
  float* out; const float* gp; const float* s_hat; int C, N, K;
  __device__ __forceinline__ void operator()(int m, int n, float v) const {
    const int b = n / N, px = n - b * N;
    const float* g = gp + static_cast<long long>(b) * K * N + px;
    for (int k = 0; k < K; ++k) v = fmaf(g[static_cast<long long>(k) * N], __ldg(s_hat + k * C + m), v);
    out[(static_cast<long long>(b) * C + m) * N + px] = v;
  }
};

// ------------------------------------------------------------------------------------------ small kernels
__global__ void fill_pm_one_kernel(float* one, float* neg, int K) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < K) { one[i] = 1.f; neg[i] = -1.f; }
}

// d_alpha / d_beta reductions and the per-pixel projection gradient gp = g * (p >= 0 ? alpha : -beta).
// grid (chunks, K, B), 256 threads.
struct FgChannels { int ch[SL_MAX_CLASSES]; };
__global__ void __launch_bounds__(256) fg_coef_grad_kernel(const float* __restrict__ p, const float* __restrict__ g,
                                                           const float* __restrict__ alpha, const float* __restrict__ beta,
                                                           int K, int N, int Ktot, FgChannels chs, float* __restrict__ gp,
                                                           float* __restrict__ d_alpha, float* __restrict__ d_beta) {
  const int k = blockIdx.y, b = blockIdx.z;
  const float a = alpha[k], bt = beta[k];
  const float* pk = p + (static_cast<long long>(b) * K + k) * N;
  const float* gk = g + (static_cast<long long>(b) * Ktot + chs.ch[k]) * N;
  float* gpk = gp + (static_cast<long long>(b) * K + k) * N;
  float da = 0.f, db = 0.f;
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
    const float pv = pk[n], gv = gk[n];
    const bool pos = pv >= 0.f;
    gpk[n] = gv * (pos ? a : -bt);
    const float t = gv * pv;
    da += pos ? t : 0.f;
    db -= pos ? 0.f : t;
  }
  __shared__ float sa[8], sb[8];
  da = warp_sum(da); db = warp_sum(db);
  if ((threadIdx.x & 31) == 0) { sa[threadIdx.x >> 5] = da; sb[threadIdx.x >> 5] = db; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float ta = 0.f, tb = 0.f;
    for (int w = 0; w < 8; ++w) { ta += sa[w]; tb += sb[w]; }
    atomicAdd(d_alpha + k, ta);
    atomicAdd(d_beta + k, tb);
  }
}

// dw3[c] = sum_px g0[px] * h2[px][c].  grid (ceil(C/32), chunks), block (32, 8).
__global__ void __launch_bounds__(256) dw3_kernel(const float* __restrict__ h2, const float* __restrict__ g, int BNpx, int C,
                                                  int N, int Ktot, int ch, float* __restrict__ dw3) {
  const int c = blockIdx.x * 32 + threadIdx.x;
  float acc = 0.f;
  if (c < C)
    for (int px = blockIdx.y * 8 + threadIdx.y; px < BNpx; px += gridDim.y * 8) {
      const int b = px / N, n = px - b * N;
      acc = fmaf(g[(static_cast<long long>(b) * Ktot + ch) * N + n], h2[static_cast<long long>(px) * C + c], acc);
    }
  __shared__ float s[8][33];
  s[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float t = 0.f;
    for (int r = 0; r < 8; ++r) t += s[r][threadIdx.x];
    atomicAdd(dw3 + c, t);
  }
}

// d_s_hat[k][c] += sum_px gp[b][k][px] * q[b][c][px]: HBM-bound stream over the features.
// grid (ceil(C/32), px-chunks of 2048, B), 256 threads: warp w owns channels c0 + 4w .. + 3, a lane owns 8 consecutive
// pixels per iteration; the eight warps read the same gp values (L1 hits), so gp costs C/32 passes over 4KN bytes
// instead of C/8.  KG <= 12 classes of accumulators per pass over the chunk (it stays in L1/L2 for further passes).
constexpr int PG_CH = 4, PG_KG = 12, PG_CHUNK = 2048;
__global__ void __launch_bounds__(256) proto_grad_kernel(const uint16_t* __restrict__ feat, const float* __restrict__ gp,
                                                         int C, int N, int K, float* __restrict__ d_s_hat) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = blockIdx.x * 32 + warp * PG_CH, b = blockIdx.z;
  const int n_begin = blockIdx.y * PG_CHUNK, n_end = min(N, n_begin + PG_CHUNK);
  if (c0 >= C) return;
  for (int k0 = 0; k0 < K; k0 += PG_KG) {
    const int kn = min(PG_KG, K - k0);
    float acc[PG_KG][PG_CH];
#pragma unroll
    for (int k = 0; k < PG_KG; ++k)
#pragma unroll
      for (int c = 0; c < PG_CH; ++c) acc[k][c] = 0.f;
    for (int n = n_begin + lane * 8; n < n_end; n += 256) {
      float q[PG_CH][8];
#pragma unroll
      for (int c = 0; c < PG_CH; ++c) {
        uint4 u = make_uint4(0, 0, 0, 0);
        if (c0 + c < C) u = ld_stream_u4(feat + (static_cast<long long>(b) * C + c0 + c) * N + n);
        q[c][0] = bf16lo(u.x); q[c][1] = bf16hi(u.x); q[c][2] = bf16lo(u.y); q[c][3] = bf16hi(u.y);
        q[c][4] = bf16lo(u.z); q[c][5] = bf16hi(u.z); q[c][6] = bf16lo(u.w); q[c][7] = bf16hi(u.w);
      }
#pragma unroll
      for (int k = 0; k < PG_KG; ++k) {
        if (k < kn) {
          const float4* g4 = reinterpret_cast<const float4*>(gp + (static_cast<long long>(b) * K + k0 + k) * N + n);
          const float4 ga = __ldg(g4), gb = __ldg(g4 + 1);
          const float gv[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
#pragma unroll
          for (int c = 0; c < PG_CH; ++c)
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[k][c] = fmaf(gv[e], q[c][e], acc[k][c]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < PG_KG; ++k)
#pragma unroll
      for (int c = 0; c < PG_CH; ++c) {
        const float v = warp_sum(acc[k][c]);
        if (lane == 0 && k < kn && c0 + c < C) atomicAdd(d_s_hat + (k0 + k) * C + c0 + c, v);
      }
  }
}

// d_feat[b][c][n] += sum_k gp[b][k][n] * s_hat[k][c]  (the projection term of d_q; the GEMM wrote W1'^T dz1).
// grid (px-chunks of 1024, ceil(C/64), B), 256 threads; a thread keeps gp of its 4 pixels in registers.
__global__ void __launch_bounds__(256) dfeat_proj_kernel(const float* __restrict__ gp, const float* __restrict__ s_hat, int C,
                                                         int N, int K, float* __restrict__ d_feat) {
  const int n = blockIdx.x * 1024 + threadIdx.x * 4, b = blockIdx.z;
  const int c_begin = blockIdx.y * 64, c_end = min(C, c_begin + 64);
  if (n >= N) return;
  for (int k0 = 0; k0 < K; k0 += PG_KG) {
    const int kn = min(PG_KG, K - k0);
    float4 g[PG_KG];
#pragma unroll
    for (int k = 0; k < PG_KG; ++k)
      g[k] = k < kn ? __ldg(reinterpret_cast<const float4*>(gp + (static_cast<long long>(b) * K + k0 + k) * N + n))
                    : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int cb = c_begin; cb < c_end; cb += 4) {          // four channels per step: four 16-byte loads in flight
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        v[u] = cb + u < c_end ? *reinterpret_cast<const float4*>(d_feat + (static_cast<long long>(b) * C + cb + u) * N + n)
                              : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < PG_KG; ++k) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float s = (k < kn && cb + u < c_end) ? __ldg(s_hat + (k0 + k) * C + cb + u) : 0.f;
          v[u].x = fmaf(g[k].x, s, v[u].x); v[u].y = fmaf(g[k].y, s, v[u].y);
          v[u].z = fmaf(g[k].z, s, v[u].z); v[u].w = fmaf(g[k].w, s, v[u].w);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (cb + u < c_end) *reinterpret_cast<float4*>(d_feat + (static_cast<long long>(b) * C + cb + u) * N + n) = v[u];
    }
  }
}

}  // namespace bwd
}  // namespace sl

// tensor-core GEMM chain (pop_bwd_tc.cu)
int sl_pop_bwd_flag_cap(long long px, int C);
int sl_pop_bwd_tc_run(const uint16_t* feat, int B, int C, int N, const float* s_hat, int K, const float* W1p, const float* W2,
                      const float* w3, const float* g_logits, int Ktot, int bg_ch, const float* gp, float* dW1p,
                      float* dW2, float* dw3, float* d_feat, uint16_t* act, uint16_t* wsplit, cudaStream_t st);

extern "C" size_t sl_pop_head_bwd_ws_bytes(int B, int C, int N, int K) {
  if (B < 1 || C < 1 || N < 1 || K < 1) return 0;
  const size_t px = static_cast<size_t>(B) * N;
  // p, gp [B,K,N] fp32; h1, dz2, dz1 [B*N, C] (fp32 each, or bf16 hi/lo pairs on the tensor-core path);
  // +-1 coefficient vectors; bf16 hi/lo of W1', W2 and their transposes
  // tensor-core path: 7 bf16 activation planes (h1 x3, dz2 x2, dz1 x2), 9 bf16 weight planes, the sign-fix-up queue
  const size_t act = (3 * px * C * 4 > 7 * px * C * 2 ? 3 * px * C * 4 : 7 * px * C * 2);
  return (2 * px * K + 2 * SL_MAX_CLASSES) * sizeof(float) + act + 9 * static_cast<size_t>(C) * C * 2 +
         (static_cast<size_t>(sl_pop_bwd_flag_cap(static_cast<long long>(px), C)) + 4) * sizeof(int) + 256;
}

extern "C" int sl_pop_head_bwd(const uint16_t* feat, int B, int C, int N, const float* s_hat, const float* alpha,
                               const float* beta, int K, const int* fg_ch_host, const float* W1p, const float* W2,
                               const float* w3, const float* g_logits, int Ktot, int bg_ch, float* d_s_hat,
                               float* d_alpha, float* d_beta, float* dW1p, float* dW2, float* dw3, float* d_feat,
                               int mode, void* ws, void* stream) {
  using namespace sl::bwd;
  SL_CHECK_PTR(feat); SL_CHECK_PTR(s_hat); SL_CHECK_PTR(alpha); SL_CHECK_PTR(beta); SL_CHECK_PTR(fg_ch_host);
  SL_CHECK_PTR(W1p); SL_CHECK_PTR(W2); SL_CHECK_PTR(w3); SL_CHECK_PTR(g_logits); SL_CHECK_PTR(d_s_hat);
  SL_CHECK_PTR(d_alpha); SL_CHECK_PTR(d_beta); SL_CHECK_PTR(dW1p); SL_CHECK_PTR(dW2); SL_CHECK_PTR(dw3); SL_CHECK_PTR(ws);
  SL_CHECK_ARG(B >= 1 && K >= 1 && K < SL_MAX_CLASSES && Ktot > K && Ktot <= SL_MAX_CLASSES);
  SL_CHECK_ARG(C >= 8 && C <= 512 && C % 8 == 0 && N >= 8 && N % 8 == 0);
  SL_CHECK_ARG(bg_ch >= 0 && bg_ch < Ktot);
  SL_CHECK_ARG(mode == SL_BWD_AUTO || mode == SL_BWD_SIMT);
  SL_CHECK_ARG(static_cast<long long>(B) * N < (1ll << 31));
  SL_CHECK_ALIGN(feat, 16); SL_CHECK_ALIGN(ws, 128);
  FgChannels chs;
  for (int k = 0; k < SL_MAX_CLASSES; ++k) chs.ch[k] = 0;
  for (int k = 0; k < K; ++k) {
    SL_CHECK_ARG(fg_ch_host[k] >= 0 && fg_ch_host[k] < Ktot && fg_ch_host[k] != bg_ch);
    chs.ch[k] = fg_ch_host[k];
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int BNpx = B * N;
  const size_t px = static_cast<size_t>(BNpx);
  float* p = static_cast<float*>(ws);
  float* gp = p + px * K;
  float* h1 = gp + px * K;
  float* dz2 = h1 + px * C;
  float* dz1 = dz2 + px * C;          // holds h2 = relu(z2) until dw3 has been reduced
  // +-1 vectors and the bf16 weight planes sit behind the larger (tensor-core) activation region
  float* one = h1 + (7 * px * C * 2 + 3) / 4;
  float* neg = one + SL_MAX_CLASSES;
  uint16_t* wsplit = reinterpret_cast<uint16_t*>(neg + SL_MAX_CLASSES);

  cudaMemsetAsync(d_s_hat, 0, sizeof(float) * K * C, st);
  cudaMemsetAsync(d_alpha, 0, sizeof(float) * K, st);
  cudaMemsetAsync(d_beta, 0, sizeof(float) * K, st);
  cudaMemsetAsync(dW1p, 0, sizeof(float) * C * C, st);
  cudaMemsetAsync(dW2, 0, sizeof(float) * C * C, st);
  cudaMemsetAsync(dw3, 0, sizeof(float) * C, st);

  // projections p_k = s_hat_k . q: the foreground kernel with alpha = 1, beta = -1 returns them unscaled
  fill_pm_one_kernel<<<1, 32, 0, st>>>(one, neg, K);
  {
    int ident[SL_MAX_CLASSES];
    for (int k = 0; k < K; ++k) ident[k] = k;
    const int rc = sl_pop_fg_lowres(feat, B, C, N, s_hat, one, neg, K, p, K, ident, stream);
    if (rc != 0) return rc;
  }
  {
    const int chunks = max(1, min((N + 1023) / 1024, 64));
    fg_coef_grad_kernel<<<dim3(chunks, K, B), 256, 0, st>>>(p, g_logits, alpha, beta, K, N, Ktot, chs, gp, d_alpha, d_beta);
  }
  const FeatPxCh fq{feat, C, N};
  // d_s_hat[k][c] = sum_px gp[k][px] q[px][c]
  proto_grad_kernel<<<dim3((C + 31) / 32, (N + PG_CHUNK - 1) / PG_CHUNK, B), 256, 0, st>>>(feat, gp, C, N, K, d_s_hat);
  if (mode == SL_BWD_AUTO && C >= 32) {
    const int rc = sl_pop_bwd_tc_run(feat, B, C, N, s_hat, K, W1p, W2, w3, g_logits, Ktot, bg_ch, gp, dW1p, dW2, dw3, d_feat,
                                     reinterpret_cast<uint16_t*>(h1), wsplit, st);
    if (rc != 0) return rc;
    if (d_feat != nullptr)
      dfeat_proj_kernel<<<dim3((N + 1023) / 1024, (C + 63) / 64, B), 256, 0, st>>>(gp, s_hat, C, N, K, d_feat);
    return SL_LAUNCH_RESULT();
  }
  const int px_splits = max(1, min(BNpx / 512, 2 * sl::num_sms() / (((C + BM - 1) / BM) * ((C + BN - 1) / BN))));
  // h1 = relu(W1' q)
  sgemm<true, false>(BNpx, C, C, 1, fq, ColMajor{W1p, C}, EpiRelu{h1, C}, st);
  // z2 = W2 h1 -> dz2, h2
  sgemm<false, false>(BNpx, C, C, 1, RowMajor{h1, C}, ColMajor{W2, C}, EpiLayer2{dz2, dz1, w3, g_logits, C, N, Ktot, bg_ch}, st);
  dw3_kernel<<<dim3((C + 31) / 32, max(1, min(BNpx / 64, 256))), dim3(32, 8), 0, st>>>(dz1, g_logits, BNpx, C, N, Ktot, bg_ch, dw3);
  // dW2[i][j] = sum_px dz2[px][i] h1[px][j]
  sgemm<true, true>(C, C, BNpx, px_splits, ColMajor{dz2, C}, RowMajor{h1, C}, EpiAtomic{dW2, C}, st);
  // dz1 = (dz2 W2) [h1 > 0]
  sgemm<false, true>(BNpx, C, C, 1, RowMajor{dz2, C}, RowMajor{W2, C}, EpiMask{dz1, h1, C}, st);
  // dW1'[i][j] = sum_px dz1[px][i] q[px][j]
  sgemm<true, false>(C, C, BNpx, px_splits, ColMajor{dz1, C}, fq, EpiAtomic{dW1p, C}, st);
  if (d_feat != nullptr)
    sgemm<true, false>(C, BNpx, C, 1, ColMajor{W1p, C}, ColMajor{dz1, C}, EpiDFeat{d_feat, gp, s_hat, C, N, K}, st);
  return SL_LAUNCH_RESULT();
}
