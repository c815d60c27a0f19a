// sl_pop_prepare: everything in the POP head that depends only on weights and prototypes.
//   reference: networks/pspnet_pop.py:106,113 (F.normalize of the prototypes),
//              :46-52,57-63 (classifier / classifier_n), :112,118 (bg = q - sum_k p_k s_k).
// Three tiny kernels (K <= 31, C <= 1024): total work ~ K*C^2 FMA, negligible next to one tile.
#include <cuda_fp16.h>
#include <cstdlib>
#include "common.cuh"

namespace sl {

// s_hat[k] = protos[k] / max(||protos[k]||_2, 1e-12)
__global__ void __launch_bounds__(128) normalize_protos_kernel(const float* __restrict__ protos, int C,
                                                               float* __restrict__ s_hat) {
  __shared__ float red[4];
  const float* row = protos + static_cast<size_t>(blockIdx.x) * C;
  float ss = 0.f;
  for (int i = threadIdx.x; i < C; i += 128) { const float v = row[i]; ss = fmaf(v, v, ss); }
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  const float nrm = fmaxf(sqrtf(red[0] + red[1] + red[2] + red[3]), 1e-12f);
  for (int i = threadIdx.x; i < C; i += 128) s_hat[static_cast<size_t>(blockIdx.x) * C + i] = row[i] / nrm;
}

// alpha_k = MLP(+s_hat_k), beta_k = MLP(-s_hat_k) for all K classes: three small launches, one per
// MLP layer, so that every weight row is read once by one warp (C/8 CTAs) instead of K*2 CTAs each
// streaming both C x C matrices.  Vector v in [0, 2K): class v>>1, sign = v&1 ? -1 : +1.
// Classes [0,Kb) use the fg weights, [Kb,K) the bg weights.

// One warp per output row o keeps that weight row in registers and sweeps the input vectors, which
// the CTA stages in shared memory VCHUNK at a time (so they are read from L2 once per CTA, not once
// per warp, and every dot product streams from conflict-free shared memory).
constexpr int VCHUNK = 8;          // input vectors staged per pass: 8 * C * 4 B <= 32 KB at C = 1024
constexpr int MAXC_REG = 32;       // C <= 1024 -> at most 32 weights per lane

// h1[v][o] = relu(+-W1x[o] . s_hat_k)        C/8 CTAs of 256 threads (bx = CTA index among them)
// nrm == nullptr: vecs holds s_hat; otherwise vecs holds the raw prototypes and s_hat_k[i] = vecs[k][i] / nrm[k] is
// formed while staging (the same division normalize_protos_kernel performs, so the values are identical).
__device__ __forceinline__ void mlp1_body(int bx, const float* __restrict__ vecs, const float* nrm, int K, int Kb, int C,
                                          const float* __restrict__ W1f, const float* __restrict__ W1g,
                                          float* __restrict__ h1, float* xs) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int o = bx * 8 + warp;
  const int per_lane = (C + 31) / 32;
  float wf[MAXC_REG], wg[MAXC_REG];
#pragma unroll
  for (int j = 0; j < MAXC_REG; ++j) {
    const int i = lane + 32 * j;
    const bool ok = j < per_lane && i < C && o < C;
    wf[j] = ok ? __ldg(W1f + static_cast<size_t>(o) * C + i) : 0.f;
    wg[j] = W1g == W1f ? wf[j] : (ok ? __ldg(W1g + static_cast<size_t>(o) * C + i) : 0.f);      // base mode: one MLP
  }
  for (int k0 = 0; k0 < K; k0 += VCHUNK) {
    const int nk = min(VCHUNK, K - k0);
    __syncthreads();
    if (nrm == nullptr) {
      for (int idx = threadIdx.x; idx < nk * C; idx += 256) xs[idx] = vecs[static_cast<size_t>(k0) * C + idx];
    } else {
      for (int kk = 0; kk < nk; ++kk) {
        const float nk_ = nrm[k0 + kk];
        for (int i = threadIdx.x; i < C; i += 256) xs[kk * C + i] = vecs[static_cast<size_t>(k0 + kk) * C + i] / nk_;
      }
    }
    __syncthreads();
    // all staged vectors at once: VCHUNK independent accumulation chains instead of one (the loop was latency-bound)
    float acc[VCHUNK];
#pragma unroll
    for (int kk = 0; kk < VCHUNK; ++kk) acc[kk] = 0.f;
#pragma unroll
    for (int j = 0; j < MAXC_REG; ++j)
      if (j < per_lane) {
        const int i = lane + 32 * j;
        if (i < C) {
#pragma unroll
          for (int kk = 0; kk < VCHUNK; ++kk)
            if (kk < nk) acc[kk] = fmaf((k0 + kk) < Kb ? wf[j] : wg[j], xs[kk * C + i], acc[kk]);
        }
      }
#pragma unroll
    for (int kk = 0; kk < VCHUNK; ++kk) acc[kk] = warp_sum(acc[kk]);
    if (lane == 0 && o < C) {
#pragma unroll
      for (int kk = 0; kk < VCHUNK; ++kk)
        if (kk < nk) {
          h1[static_cast<size_t>(2 * (k0 + kk)) * C + o] = fmaxf(acc[kk], 0.f);
          h1[static_cast<size_t>(2 * (k0 + kk) + 1) * C + o] = fmaxf(-acc[kk], 0.f);
        }
    }
  }
}

__global__ void __launch_bounds__(256) mlp1_kernel(const float* __restrict__ s_hat, int K, int Kb, int C,
                                                   const float* __restrict__ W1f, const float* __restrict__ W1g,
                                                   float* __restrict__ h1) {
  extern __shared__ float xs[];     // [VCHUNK][C]
  mlp1_body(blockIdx.x, s_hat, nullptr, K, Kb, C, W1f, W1g, h1, xs);
}

// h2[v][o] = relu(W2x[o] . h1[v])
__global__ void __launch_bounds__(256) mlp2_kernel(const float* __restrict__ h1, int K, int Kb, int C,
                                                   const float* __restrict__ W2f, const float* __restrict__ W2g,
                                                   float* __restrict__ h2) {
  extern __shared__ float xs[];     // [VCHUNK][C]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int o = blockIdx.x * 8 + warp;
  const int per_lane = (C + 31) / 32;
  float wf[MAXC_REG], wg[MAXC_REG];
#pragma unroll
  for (int j = 0; j < MAXC_REG; ++j) {
    const int i = lane + 32 * j;
    const bool ok = j < per_lane && i < C && o < C;
    wf[j] = ok ? __ldg(W2f + static_cast<size_t>(o) * C + i) : 0.f;
    wg[j] = W2g == W2f ? wf[j] : (ok ? __ldg(W2g + static_cast<size_t>(o) * C + i) : 0.f);      // base mode: one MLP
  }
  const int V = 2 * K;
  for (int v0 = 0; v0 < V; v0 += VCHUNK) {
    const int nv = min(VCHUNK, V - v0);
    __syncthreads();
    for (int idx = threadIdx.x; idx < nv * C; idx += 256) xs[idx] = h1[static_cast<size_t>(v0) * C + idx];
    __syncthreads();
    float acc[VCHUNK];
#pragma unroll
    for (int vv = 0; vv < VCHUNK; ++vv) acc[vv] = 0.f;
#pragma unroll
    for (int j = 0; j < MAXC_REG; ++j)
      if (j < per_lane) {
        const int i = lane + 32 * j;
        if (i < C) {
#pragma unroll
          for (int vv = 0; vv < VCHUNK; ++vv)
            if (vv < nv) acc[vv] = fmaf(((v0 + vv) >> 1) < Kb ? wf[j] : wg[j], xs[vv * C + i], acc[vv]);
        }
      }
#pragma unroll
    for (int vv = 0; vv < VCHUNK; ++vv) acc[vv] = warp_sum(acc[vv]);
    if (lane == 0 && o < C) {
#pragma unroll
      for (int vv = 0; vv < VCHUNK; ++vv)
        if (vv < nv) h2[static_cast<size_t>(v0 + vv) * C + o] = fmaxf(acc[vv], 0.f);
    }
  }
}

// alpha/beta = w3x . h2[v]                   grid 2K, 128 threads
__global__ void __launch_bounds__(128) mlp3_kernel(const float* __restrict__ h2, int Kb, int C,
                                                   const float* __restrict__ w3f, const float* __restrict__ w3g,
                                                   float* __restrict__ alpha, float* __restrict__ beta) {
  __shared__ float red[4];
  const int v = blockIdx.x, k = v >> 1;
  const float* w3 = k < Kb ? w3f : w3g;
  const float* x = h2 + static_cast<size_t>(v) * C;
  float acc = 0.f;
  for (int i = threadIdx.x; i < C; i += 128) acc = fmaf(__ldg(w3 + i), x[i], acc);
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) ((v & 1) ? beta : alpha)[k] = red[0] + red[1] + red[2] + red[3];
}

// W1' = W1 (I - S^T S):  row o of W1' = W1[o] - sum_k (W1[o] . s_k) s_k.   grid C, 128 threads.
// Also emits the layouts both background paths want: transposed fp32 and split bf16.
// THREADS threads fold row o; nrm as in mlp1_body (nullptr: vecs = s_hat, else vecs = raw prototypes, divided on the fly).
// Per-element arithmetic (lane-strided dot products, class-ordered correction) does not depend on THREADS.
template <int THREADS>
__device__ __forceinline__ void fold_body(int o, const float* __restrict__ vecs, const float* nrm, int K, int C,
                                          const float* __restrict__ W1, const float* __restrict__ W2,
                                          float* __restrict__ W1p_t, float* __restrict__ W2_t,
                                          uint16_t* __restrict__ W1p_hi, uint16_t* __restrict__ W1p_lo,
                                          uint16_t* __restrict__ W2_hi, uint16_t* __restrict__ W2_lo,
                                          uint16_t* __restrict__ W1p_f16, uint16_t* __restrict__ W2_f16, float* u) {
  const float* w1 = W1 + static_cast<size_t>(o) * C;
  const float* w2 = W2 + static_cast<size_t>(o) * C;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto sv = [&](int k, int i) {
    const float v = vecs[static_cast<size_t>(k) * C + i];
    return nrm == nullptr ? v : v / nrm[k];
  };
  for (int k = warp; k < K; k += THREADS / 32) {
    float acc = 0.f;
    for (int i = lane; i < C; i += 32) acc = fmaf(w1[i], sv(k, i), acc);
    acc = warp_sum(acc);
    if (lane == 0) u[k] = acc;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += THREADS) {
    float corr = 0.f;
    for (int k = 0; k < K; ++k) corr = fmaf(u[k], sv(k, i), corr);
    const float a = w1[i] - corr;
    const float b = w2[i];
    if (W1p_t) W1p_t[static_cast<size_t>(i) * C + o] = a;
    if (W2_t) W2_t[static_cast<size_t>(i) * C + o] = b;
    if (W1p_hi) {
      const uint16_t ah = f32_to_bf16_rn(a), bh = f32_to_bf16_rn(b);
      const size_t idx = static_cast<size_t>(o) * C + i;
      W1p_hi[idx] = ah;
      W1p_lo[idx] = f32_to_bf16_rn(a - bf16_bits_to_f32(ah));
      W2_hi[idx] = bh;
      W2_lo[idx] = f32_to_bf16_rn(b - bf16_bits_to_f32(bh));
    }
    if (W1p_f16) {
      const size_t idx = static_cast<size_t>(o) * C + i;
      W1p_f16[idx] = __half_as_ushort(__float2half_rn(a));
      W2_f16[idx] = __half_as_ushort(__float2half_rn(b));
    }
  }
}

__global__ void __launch_bounds__(128) fold_weights_kernel(const float* __restrict__ s_hat, int K, int C,
                                                           const float* __restrict__ W1, const float* __restrict__ W2,
                                                           float* __restrict__ W1p_t, float* __restrict__ W2_t,
                                                           uint16_t* __restrict__ W1p_hi, uint16_t* __restrict__ W1p_lo,
                                                           uint16_t* __restrict__ W2_hi, uint16_t* __restrict__ W2_lo,
                                                           uint16_t* __restrict__ W1p_f16, uint16_t* __restrict__ W2_f16) {
  __shared__ float u[SL_MAX_CLASSES];
  fold_body<128>(blockIdx.x, s_hat, nullptr, K, C, W1, W2, W1p_t, W2_t, W1p_hi, W1p_lo, W2_hi, W2_lo, W1p_f16, W2_f16, u);
}

// First launch of sl_pop_prepare: prototype normalisation, MLP layer 1 and the weight fold in ONE grid.  The three were
// separate launches (normalise -> layer 1 -> ... and, at the end of the chain, the fold, which only needs s_hat); each
// tiny kernel costs a launch + load-latency chain of 8-13 us, so the chain was 52 us for ~K C^2 FMAs.  Here every CTA
// derives the K norms itself (K rows of C floats from L2) and divides while it reads, CTAs [0, nb1) run layer 1,
// CTAs [nb1, nb1 + C) fold one weight row each, and CTA 0 also writes s_hat for the later stages.
// The norms are evaluated exactly as normalize_protos_kernel does (128 threads: thread t sums i = t, t+128, ...; a
// butterfly per warp; red[0] + red[1] + red[2] + red[3]): one warp per row, lane l playing threads l, l+32, l+64, l+96.
__global__ void __launch_bounds__(256) prep_stage1_kernel(const float* __restrict__ protos, int K, int Kb, int C, int nb1,
                                                          const float* __restrict__ W1f, const float* __restrict__ W1g,
                                                          const float* __restrict__ W2g, float* __restrict__ s_hat,
                                                          float* __restrict__ h1, float* __restrict__ W1p_t,
                                                          float* __restrict__ W2_t, uint16_t* __restrict__ W1p_hi,
                                                          uint16_t* __restrict__ W1p_lo, uint16_t* __restrict__ W2_hi,
                                                          uint16_t* __restrict__ W2_lo, uint16_t* __restrict__ W1p_f16,
                                                          uint16_t* __restrict__ W2_f16) {
  extern __shared__ float xs[];     // [VCHUNK][C] (layer-1 CTAs)
  __shared__ float nrm_s[SL_MAX_CLASSES];
  __shared__ float u[SL_MAX_CLASSES];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int k = warp; k < K; k += 8) {
    const float* row = protos + static_cast<size_t>(k) * C;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int i0 = 0; i0 < C; i0 += 128) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int i = i0 + 32 * q + lane;
        if (i < C) { const float v = row[i]; acc[q] = fmaf(v, v, acc[q]); }
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[q] = warp_sum(acc[q]);
    if (lane == 0) nrm_s[k] = fmaxf(sqrtf(acc[0] + acc[1] + acc[2] + acc[3]), 1e-12f);
  }
  __syncthreads();
  if (blockIdx.x == 0)
    for (int k = 0; k < K; ++k)
      for (int i = threadIdx.x; i < C; i += 256)
        s_hat[static_cast<size_t>(k) * C + i] = protos[static_cast<size_t>(k) * C + i] / nrm_s[k];
  if (static_cast<int>(blockIdx.x) < nb1)
    mlp1_body(blockIdx.x, protos, nrm_s, K, Kb, C, W1f, W1g, h1, xs);
  else
    fold_body<256>(static_cast<int>(blockIdx.x) - nb1, protos, nrm_s, K, C, W1g, W2g, W1p_t, W2_t, W1p_hi, W1p_lo, W2_hi,
                   W2_lo, W1p_f16, W2_f16, u);
}

// ------------------------------------------------------------------------------------------------------------
// Backward of sl_pop_prepare (training): the O(K C^2) parameter-side chain
//   s_hat = normalize(protos);  alpha/beta = MLP(+-s_hat);  W1' = W1_bg (I - S_hat^T S_hat)
// given dL/d{s_hat (through the projections), alpha, beta, W1', W2_bg, w3_bg} from sl_pop_head_bwd.
// All kernels are tiny (K <= 31 vectors of C <= 1024); clarity over speed.

// da2[v] = dy_v w3 [h2[v] > 0];  dw3_target += dy_v h2[v].       grid 2K, 128 threads
__global__ void __launch_bounds__(128) pb_dlayer3_kernel(const float* __restrict__ h2, const float* __restrict__ d_alpha,
                                                         const float* __restrict__ d_beta, int Kb, int C,
                                                         const float* __restrict__ w3f, const float* __restrict__ w3g,
                                                         float* __restrict__ da2, float* dw3f, float* dw3g) {
  const int v = blockIdx.x, k = v >> 1;
  const float dy = (v & 1) ? d_beta[k] : d_alpha[k];
  const float* w3 = k < Kb ? w3f : w3g;
  float* dw3 = k < Kb ? dw3f : dw3g;
  for (int i = threadIdx.x; i < C; i += 128) {
    const float h = h2[static_cast<size_t>(v) * C + i];
    da2[static_cast<size_t>(v) * C + i] = h > 0.f ? dy * w3[i] : 0.f;
    if (dw3 != nullptr) atomicAdd(dw3 + i, dy * h);
  }
}

// out[v][i] = [mask[v][i] > 0] * sum_o W_sel[o][i] d[v][o]   (transposed mat-vec).   grid (ceil(C/32), V), 128 threads:
// lane = column i, the four warps split the rows o (coalesced 128-byte row segments), partial sums meet in shared memory.
__global__ void __launch_bounds__(128) pb_matvec_t_kernel(const float* __restrict__ d, int Kb, int C,
                                                          const float* __restrict__ Wf, const float* __restrict__ Wg,
                                                          const float* __restrict__ mask, float* __restrict__ out) {
  extern __shared__ float dv[];                       // [C] + [4][32]
  float* part = dv + C;
  const int v = blockIdx.y, k = v >> 1;
  const float* W = k < Kb ? Wf : Wg;
  for (int o = threadIdx.x; o < C; o += 128) dv[o] = d[static_cast<size_t>(v) * C + o];
  __syncthreads();
  const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;
  float acc = 0.f;
  if (i < C)
    for (int o = g; o < C; o += 4) acc = fmaf(__ldg(W + static_cast<size_t>(o) * C + i), dv[o], acc);
  part[g * 32 + lane] = acc;
  __syncthreads();
  if (g == 0 && i < C) {
    acc = (part[lane] + part[32 + lane]) + (part[64 + lane] + part[96 + lane]);
    if (mask != nullptr && !(mask[static_cast<size_t>(v) * C + i] > 0.f)) acc = 0.f;
    out[static_cast<size_t>(v) * C + i] = acc;
  }
}

// dW_target[o][i] += sum_{v of that target} a[v][o] * b[v][i];  b = h1, or +-s_hat when b_is_shat.   grid C, 128 threads
__global__ void __launch_bounds__(128) pb_outer_kernel(const float* __restrict__ a, const float* __restrict__ b, int b_is_shat,
                                                       int K, int Kb, int C, float* dWf, float* dWg) {
  const int o = blockIdx.x;
  for (int i = threadIdx.x; i < C; i += 128) {
    float sf = 0.f, sg = 0.f;
    for (int v = 0; v < 2 * K; ++v) {
      const int k = v >> 1;
      const float bv = b_is_shat ? ((v & 1) ? -b[static_cast<size_t>(k) * C + i] : b[static_cast<size_t>(k) * C + i])
                                 : b[static_cast<size_t>(v) * C + i];
      const float t = a[static_cast<size_t>(v) * C + o] * bv;
      if (k < Kb) sf += t; else sg += t;
    }
    if (dWf == dWg) { if (dWf != nullptr) dWf[static_cast<size_t>(o) * C + i] += sf + sg; }
    else {
      if (dWf != nullptr) dWf[static_cast<size_t>(o) * C + i] += sf;
      if (dWg != nullptr) dWg[static_cast<size_t>(o) * C + i] += sg;
    }
  }
}

// out[k][o] = M[o] . s_hat[k]                                                     grid C, 128 threads
__global__ void __launch_bounds__(128) pb_rowdot_kernel(const float* __restrict__ M, const float* __restrict__ s_hat, int K,
                                                        int C, float* __restrict__ out) {
  const int o = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* row = M + static_cast<size_t>(o) * C;
  for (int k = warp; k < K; k += 4) {
    float acc = 0.f;
    for (int i = lane; i < C; i += 32) acc = fmaf(row[i], s_hat[static_cast<size_t>(k) * C + i], acc);
    acc = warp_sum(acc);
    if (lane == 0) out[static_cast<size_t>(k) * C + o] = acc;
  }
}

// dW1_bg[o][i] += D[o][i] - sum_k V[k][o] s_hat[k][i]                            grid C, 128 threads
__global__ void __launch_bounds__(128) pb_fold_dw_kernel(const float* __restrict__ D, const float* __restrict__ V,
                                                         const float* __restrict__ s_hat, int K, int C, float* dW1) {
  const int o = blockIdx.x;
  for (int i = threadIdx.x; i < C; i += 128) {
    float corr = 0.f;
    for (int k = 0; k < K; ++k) corr = fmaf(V[static_cast<size_t>(k) * C + o], s_hat[static_cast<size_t>(k) * C + i], corr);
    dW1[static_cast<size_t>(o) * C + i] += D[static_cast<size_t>(o) * C + i] - corr;
  }
}

// ds_fold[k][i] = - sum_o (U[k][o] D[o][i] + V[k][o] W1[o][i])                  grid (ceil(C/32), K), 128 threads
__global__ void __launch_bounds__(128) pb_fold_ds_kernel(const float* __restrict__ U, const float* __restrict__ V,
                                                         const float* __restrict__ D, const float* __restrict__ W1, int C,
                                                         float* __restrict__ out) {
  extern __shared__ float uv[];                       // U[k][:], V[k][:], then [4][32] partial sums
  float* part = uv + 2 * C;
  const int k = blockIdx.y;
  for (int o = threadIdx.x; o < C; o += 128) {
    uv[o] = U[static_cast<size_t>(k) * C + o];
    uv[C + o] = V[static_cast<size_t>(k) * C + o];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;
  float acc = 0.f;
  if (i < C)
    for (int o = g; o < C; o += 4)
      acc = fmaf(uv[o], __ldg(D + static_cast<size_t>(o) * C + i), fmaf(uv[C + o], __ldg(W1 + static_cast<size_t>(o) * C + i), acc));
  part[g * 32 + lane] = acc;
  __syncthreads();
  if (g == 0 && i < C)
    out[static_cast<size_t>(k) * C + i] = -((part[lane] + part[32 + lane]) + (part[64 + lane] + part[96 + lane]));
}

// d_protos[k] = (g - s_hat_k (s_hat_k . g)) / max(||protos_k||, 1e-12),  g = d_s_in + dx[+] - dx[-] + ds_fold
__global__ void __launch_bounds__(128) pb_finalize_kernel(const float* __restrict__ protos, const float* __restrict__ s_hat,
                                                          const float* __restrict__ d_s_in, const float* __restrict__ dx,
                                                          const float* __restrict__ ds_fold, int C, float* __restrict__ d_protos) {
  __shared__ float red[2][4];
  const int k = blockIdx.x;
  const size_t r = static_cast<size_t>(k) * C;
  float ss = 0.f, dot = 0.f;
  for (int i = threadIdx.x; i < C; i += 128) {
    const float g = d_s_in[r + i] + dx[2 * r + i] - dx[2 * r + C + i] + ds_fold[r + i];
    const float pv = protos[r + i];
    ss = fmaf(pv, pv, ss);
    dot = fmaf(s_hat[r + i], g, dot);
  }
  ss = warp_sum(ss); dot = warp_sum(dot);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = ss; red[1][threadIdx.x >> 5] = dot; }
  __syncthreads();
  const float nrm = fmaxf(sqrtf(red[0][0] + red[0][1] + red[0][2] + red[0][3]), 1e-12f);
  const float d = red[1][0] + red[1][1] + red[1][2] + red[1][3];
  for (int i = threadIdx.x; i < C; i += 128) {
    const float g = d_s_in[r + i] + dx[2 * r + i] - dx[2 * r + C + i] + ds_fold[r + i];
    d_protos[r + i] = (g - s_hat[r + i] * d) / nrm;
  }
}

}  // namespace sl

extern "C" size_t sl_pop_prepare_ws_bytes(int K, int C) {
  if (K < 1 || C < 1) return 0;
  return static_cast<size_t>(4) * K * C * sizeof(float);
}

extern "C" int sl_pop_prepare(const float* protos, int K, int Kb, int C, const float* W1_fg, const float* W2_fg,
                              const float* w3_fg, const float* W1_bg, const float* W2_bg, const float* w3_bg,
                              float* s_hat, float* alpha, float* beta, float* W1p_t, float* W2_t, uint16_t* W1p_hi,
                              uint16_t* W1p_lo, uint16_t* W2_hi, uint16_t* W2_lo, uint16_t* W1p_f16,
                              uint16_t* W2_f16, float* ws, void* stream) {
  SL_CHECK_ARG(K >= 1 && K < SL_MAX_CLASSES && Kb >= 0 && Kb <= K);
  SL_CHECK_ARG(C >= 8 && C <= 1024 && C % 8 == 0);
  SL_CHECK_PTR(protos); SL_CHECK_PTR(W1_fg); SL_CHECK_PTR(W2_fg); SL_CHECK_PTR(w3_fg);
  SL_CHECK_PTR(W1_bg); SL_CHECK_PTR(W2_bg); SL_CHECK_PTR(w3_bg);
  SL_CHECK_PTR(s_hat); SL_CHECK_PTR(alpha); SL_CHECK_PTR(beta); SL_CHECK_PTR(ws);
  const int n_split = (W1p_hi != nullptr) + (W1p_lo != nullptr) + (W2_hi != nullptr) + (W2_lo != nullptr);
  SL_CHECK_ARG(n_split == 0 || n_split == 4);
  SL_CHECK_ARG((W1p_f16 != nullptr) == (W2_f16 != nullptr));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* h1 = ws;                                        // [2K][C]
  float* h2 = ws + static_cast<size_t>(2) * K * C;       // [2K][C]
  const size_t xs_bytes = static_cast<size_t>(sl::VCHUNK) * C * sizeof(float);
  const int nb1 = (C + 7) / 8;
  const bool fold = W1p_t || W2_t || n_split || W1p_f16;
  // three launches: {normalise, layer 1, fold} -> layer 2 -> layer 3 (SL_PREP_SPLIT=1: the five separate launches)
  if (sl::env().prep_split == 1) {
    sl::normalize_protos_kernel<<<K, 128, 0, st>>>(protos, C, s_hat);
    sl::mlp1_kernel<<<nb1, 256, xs_bytes, st>>>(s_hat, K, Kb, C, W1_fg, W1_bg, h1);
  } else {
    sl::prep_stage1_kernel<<<nb1 + (fold ? C : 0), 256, xs_bytes, st>>>(protos, K, Kb, C, nb1, W1_fg, W1_bg, W2_bg, s_hat, h1,
                                                                         W1p_t, W2_t, W1p_hi, W1p_lo, W2_hi, W2_lo, W1p_f16,
                                                                         W2_f16);
  }
  sl::mlp2_kernel<<<nb1, 256, xs_bytes, st>>>(h1, K, Kb, C, W2_fg, W2_bg, h2);
  sl::mlp3_kernel<<<2 * K, 128, 0, st>>>(h2, Kb, C, w3_fg, w3_bg, alpha, beta);
  if (sl::env().prep_split == 1 && fold)
    sl::fold_weights_kernel<<<C, 128, 0, st>>>(s_hat, K, C, W1_bg, W2_bg, W1p_t, W2_t, W1p_hi, W1p_lo, W2_hi, W2_lo,
                                               W1p_f16, W2_f16);
  return SL_LAUNCH_RESULT();
}

extern "C" size_t sl_pop_prepare_bwd_ws_bytes(int K, int C) {
  if (K < 1 || C < 1) return 0;
  return static_cast<size_t>(14) * K * C * sizeof(float);
}

extern "C" int sl_pop_prepare_bwd(const float* protos, int K, int Kb, int C, const float* W1_fg, const float* W2_fg,
                                  const float* w3_fg, const float* W1_bg, const float* W2_bg, const float* w3_bg,
                                  const float* d_s_hat, const float* d_alpha, const float* d_beta, const float* dW1p,
                                  const float* dW2_direct, const float* dw3_direct, float* d_protos, float* dW1_fg,
                                  float* dW2_fg, float* dw3_fg, float* dW1_bg, float* dW2_bg, float* dw3_bg, float* ws,
                                  void* stream) {
  SL_CHECK_ARG(K >= 1 && K < SL_MAX_CLASSES && Kb >= 0 && Kb <= K);
  SL_CHECK_ARG(C >= 8 && C <= 1024 && C % 8 == 0);
  SL_CHECK_PTR(protos); SL_CHECK_PTR(W1_fg); SL_CHECK_PTR(W2_fg); SL_CHECK_PTR(w3_fg);
  SL_CHECK_PTR(W1_bg); SL_CHECK_PTR(W2_bg); SL_CHECK_PTR(w3_bg);
  SL_CHECK_PTR(d_s_hat); SL_CHECK_PTR(d_alpha); SL_CHECK_PTR(d_beta); SL_CHECK_PTR(dW1p); SL_CHECK_PTR(dW2_direct);
  SL_CHECK_PTR(dw3_direct); SL_CHECK_PTR(d_protos); SL_CHECK_PTR(dW1_bg); SL_CHECK_PTR(dW2_bg); SL_CHECK_PTR(dw3_bg);
  SL_CHECK_PTR(ws);
  const bool shared = W1_fg == W1_bg && W2_fg == W2_bg && w3_fg == w3_bg;      // base mode: one MLP for every class
  const int n_fg = (dW1_fg != nullptr) + (dW2_fg != nullptr) + (dw3_fg != nullptr);
  SL_CHECK_ARG(n_fg == 0 || n_fg == 3);
  SL_CHECK_ARG(!(shared && n_fg != 0));                                        // shared weights accumulate into the *_bg outputs
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t KC = static_cast<size_t>(K) * C;
  float* s_hat = ws;                 // [K][C]
  float* h1 = s_hat + KC;            // [2K][C]
  float* h2 = h1 + 2 * KC;
  float* da2 = h2 + 2 * KC;
  float* da1 = da2 + 2 * KC;
  float* dx = da1 + 2 * KC;
  float* U = dx + 2 * KC;            // [K][C]
  float* V = U + KC;
  float* dsf = V + KC;
  // outputs: background MLP starts from its direct gradients, everything else from zero
  cudaMemcpyAsync(dW2_bg, dW2_direct, sizeof(float) * C * C, cudaMemcpyDeviceToDevice, st);
  cudaMemcpyAsync(dw3_bg, dw3_direct, sizeof(float) * C, cudaMemcpyDeviceToDevice, st);
  cudaMemsetAsync(dW1_bg, 0, sizeof(float) * C * C, st);
  if (n_fg) {
    cudaMemsetAsync(dW1_fg, 0, sizeof(float) * C * C, st);
    cudaMemsetAsync(dW2_fg, 0, sizeof(float) * C * C, st);
    cudaMemsetAsync(dw3_fg, 0, sizeof(float) * C, st);
  }
  float* t1 = shared ? dW1_bg : dW1_fg;   // targets of the classes [0, Kb)
  float* t2 = shared ? dW2_bg : dW2_fg;
  float* t3 = shared ? dw3_bg : dw3_fg;
  // recompute the forward activations of the 2K coefficient vectors
  sl::normalize_protos_kernel<<<K, 128, 0, st>>>(protos, C, s_hat);
  const size_t xs_bytes = static_cast<size_t>(sl::VCHUNK) * C * sizeof(float);
  sl::mlp1_kernel<<<(C + 7) / 8, 256, xs_bytes, st>>>(s_hat, K, Kb, C, W1_fg, W1_bg, h1);
  sl::mlp2_kernel<<<(C + 7) / 8, 256, xs_bytes, st>>>(h1, K, Kb, C, W2_fg, W2_bg, h2);
  // alpha / beta -> MLP weights and s_hat
  const dim3 gv((C + 31) / 32, 2 * K);
  sl::pb_dlayer3_kernel<<<2 * K, 128, 0, st>>>(h2, d_alpha, d_beta, Kb, C, w3_fg, w3_bg, da2, t3, dw3_bg);
  sl::pb_outer_kernel<<<C, 128, 0, st>>>(da2, h1, 0, K, Kb, C, t2, dW2_bg);
  sl::pb_matvec_t_kernel<<<gv, 128, (C + 128) * sizeof(float), st>>>(da2, Kb, C, W2_fg, W2_bg, h1, da1);
  sl::pb_outer_kernel<<<C, 128, 0, st>>>(da1, s_hat, 1, K, Kb, C, t1, dW1_bg);
  sl::pb_matvec_t_kernel<<<gv, 128, (C + 128) * sizeof(float), st>>>(da1, Kb, C, W1_fg, W1_bg, nullptr, dx);
  // the fold W1' = W1_bg (I - S^T S)
  sl::pb_rowdot_kernel<<<C, 128, 0, st>>>(W1_bg, s_hat, K, C, U);
  sl::pb_rowdot_kernel<<<C, 128, 0, st>>>(dW1p, s_hat, K, C, V);
  sl::pb_fold_dw_kernel<<<C, 128, 0, st>>>(dW1p, V, s_hat, K, C, dW1_bg);
  sl::pb_fold_ds_kernel<<<dim3((C + 31) / 32, K), 128, (2 * C + 128) * sizeof(float), st>>>(U, V, dW1p, W1_bg, C, dsf);
  // through the normalisation
  sl::pb_finalize_kernel<<<K, 128, 0, st>>>(protos, s_hat, d_s_hat, dx, dsf, C, d_protos);
  return SL_LAUNCH_RESULT();
}
