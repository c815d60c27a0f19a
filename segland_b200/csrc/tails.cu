// Decoder tails that emit the POP head's bf16 NCHW features directly: the last
// operators of the reference's decoders, fused with the fp32 -> bf16 conversion the head wants, so the fp32
// feature tensor never makes a round trip through HBM between decoder and head.
//   sl_tail_layernorm      FPN_Seg_OCR_Decoder.norm over channels, networks/convnext_pop.py:13,27
//   sl_tail_bn_relu_conv   PSPModule.bottleneck[1:4] (BN -> ReLU -> 1x1 conv + bias), networks/pspnet_pop.py:19-22
//                          (same tail in PSP_Plus_Decoder.fc, networks/pspplus_pop.py:44-47)
//   sl_tail_sum            torch.stack(fpn_outs, -1).sum(-1), networks/swin_pop.py:169-172, lsk_pop.py:163-165
//   sl_tail_bn_relu        _ConvBnReLU's bn + relu after _ASPP.fc's convolution, networks/deeplab_pop.py:12-29,61,66;
//                          DoubleConv's last BN + ReLU in VGGUNet.up4, networks/vggunet_pop.py:19-20,79
//   sl_tail_concat         torch.cat([x0, x1, x2, x3], 1) of HRFPN_Seg_Decoder, networks/seghr_pop.py:23-24
// The element-wise kernels are HBM-bound (4 B read + 2 B written per element).  The 1x1 convolution runs on tcgen05:
// by default as the single fused kernel of tail_conv.cu; this file keeps the two-kernel form (BN/ReLU/split into bf16
// planes here + the generic split-bf16 GEMM of pop_bwd_tc.cu with its EPI_TAIL epilogue) for shapes outside the fused
// kernel's range and for A/B runs (SL_TAIL_FUSED=0).
#include <stdlib.h>
#include "common.cuh"

int sl_tail_conv_fused_run(const float* x, int B, int Cin, int N, const float* bn_w, const float* bn_b, const float* bn_m,
                           const float* bn_v, float eps, int relu, const uint16_t* W_hi, const uint16_t* W_lo,
                           const float* bias, int Cout, uint16_t* out, cudaStream_t st);
int sl_tail_gemm_run(const uint16_t* act_hi, const uint16_t* act_lo, int B, int Cin, int N, const uint16_t* W_hi,
                     const uint16_t* W_lo, const float* bias, int Cout, uint16_t* feat_out, cudaStream_t st);

namespace sl {
namespace tails {

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// ------------------------------------------------------------------------------------------ LayerNorm over C
// One CTA = P consecutive pixels of one image, all C channels, staged once in shared memory ([C][P] fp32, filled
// by 16-byte cp.async so the whole tile is in flight at once); exact two-pass moments from the staged tile
// (thread = pixel x channel slice, conflict-free), then every thread normalises 2 x 4 pixels of a channel row.
constexpr int LN_THREADS = 256;

template <int P>
__global__ void __launch_bounds__(LN_THREADS) layernorm_tail_kernel(const float* __restrict__ x, int C, int N,
                                                                    const float* __restrict__ gamma,
                                                                    const float* __restrict__ beta, float eps,
                                                                    uint16_t* __restrict__ out) {
  extern __shared__ __align__(16) float tile[];                 // [C][P] + red[PARTS][P] + mean[P] + rstd[P]
  constexpr int PARTS = LN_THREADS / P;
  float* red = tile + static_cast<size_t>(C) * P;
  float* mean_s = red + PARTS * P;
  float* rstd_s = mean_s + P;
  const int tiles_per_img = (N + P - 1) / P;
  const int img = blockIdx.x / tiles_per_img;
  const int n0 = (blockIdx.x - img * tiles_per_img) * P;
  const int t = threadIdx.x;
  const float* xin = x + static_cast<size_t>(img) * C * N + n0;
  const int valid = min(P, N - n0);                             // multiple of 8

  constexpr int CHUNKS = P / 4;                                 // 16-byte chunks per channel row
  const uint32_t tile_s = static_cast<uint32_t>(__cvta_generic_to_shared(tile));
  {
    const int ck = t % CHUNKS;
    if (ck * 4 < valid)
      for (int c = t / CHUNKS; c < C; c += LN_THREADS / CHUNKS)
        cp_async16(tile_s + (c * P + ck * 4) * 4, xin + static_cast<size_t>(c) * N + ck * 4);
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
  }
  __syncthreads();

  const int px = t % P, part = t / P;
  const bool px_ok = px < valid;
  const float inv_c = 1.f / static_cast<float>(C);
  {
    float s = 0.f;
    if (px_ok)
      for (int c = part; c < C; c += PARTS) s += tile[c * P + px];
    red[part * P + px] = s;
  }
  __syncthreads();
  if (t < P) {
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < PARTS; ++q) s += red[q * P + t];
    mean_s[t] = s * inv_c;
  }
  __syncthreads();
  {
    const float mu = mean_s[px];
    float s = 0.f;
    if (px_ok)
      for (int c = part; c < C; c += PARTS) {
        const float d = tile[c * P + px] - mu;
        s = fmaf(d, d, s);
      }
    red[part * P + px] = s;
  }
  __syncthreads();
  if (t < P) {
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < PARTS; ++q) s += red[q * P + t];
    rstd_s[t] = 1.f / sqrtf(s * inv_c + eps);                    // F.layer_norm: biased variance
  }
  __syncthreads();

  // output: a thread owns pixels [4*oc, 4*oc+4) and [P/2 + 4*oc, ...) of a channel row, so the two float4 reads of
  // a quarter-warp are contiguous (no bank conflicts) and each half leaves with one 8-byte bf16x4 store
  constexpr int OCT = P / 8, HALF = P / 2;
  const int oc = t % OCT;
  const bool ok0 = oc * 4 < valid, ok1 = HALF + oc * 4 < valid;
  if (ok0) {
    float mu[8], rs[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      mu[j] = mean_s[oc * 4 + j]; rs[j] = rstd_s[oc * 4 + j];
      mu[4 + j] = mean_s[HALF + oc * 4 + j]; rs[4 + j] = rstd_s[HALF + oc * 4 + j];
    }
    uint16_t* o = out + static_cast<size_t>(img) * C * N + n0 + oc * 4;
    for (int c = t / OCT; c < C; c += LN_THREADS / OCT) {
      const float4 a = *reinterpret_cast<const float4*>(tile + c * P + oc * 4);
      const float4 b = *reinterpret_cast<const float4*>(tile + c * P + HALF + oc * 4);
      const float g = __ldg(gamma + c), be = __ldg(beta + c);
      const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
      float y[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) y[j] = fmaf((v[j] - mu[j]) * rs[j], g, be);
      uint16_t* oc_row = o + static_cast<size_t>(c) * N;
      *reinterpret_cast<uint2*>(oc_row) = make_uint2(pack_bf16(y[0], y[1]), pack_bf16(y[2], y[3]));   // the head reads it next: no streaming hint
      if (ok1) *reinterpret_cast<uint2*>(oc_row + HALF) = make_uint2(pack_bf16(y[4], y[5]), pack_bf16(y[6], y[7]));
    }
  }
}

// ------------------------------------------------------------------------------------------ sum of M maps
constexpr int SUM_MAX = 8;
struct SumPtrs { const float* p[SUM_MAX]; };

__global__ void __launch_bounds__(256) sum_tail_kernel(SumPtrs maps, int M, long long n8, uint16_t* __restrict__ out) {
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n8; i += gridDim.x * 256ll) {
    float4 a = ld_stream_f4(maps.p[0] + i * 8), b = ld_stream_f4(maps.p[0] + i * 8 + 4);
#pragma unroll
    for (int m = 1; m < SUM_MAX; ++m) {                         // sequential fp32 adds, map order
      if (m >= M) break;
      const float4 c = ld_stream_f4(maps.p[m] + i * 8), d = ld_stream_f4(maps.p[m] + i * 8 + 4);
      a.x += c.x; a.y += c.y; a.z += c.z; a.w += c.w;
      b.x += d.x; b.y += d.y; b.z += d.z; b.w += d.w;
    }
    uint4 w;
    w.x = pack_bf16(a.x, a.y); w.y = pack_bf16(a.z, a.w); w.z = pack_bf16(b.x, b.y); w.w = pack_bf16(b.z, b.w);
    reinterpret_cast<uint4*>(out)[i] = w;
  }
}

// ------------------------------------------------------------------------------------------ BN -> ReLU -> split
// r = relu(bn_eval(x)) as bf16 hi/lo planes (hi = bf16(r), lo = bf16(r - hi): 16 mantissa bits), the A operand
// of the tail GEMM.  Batch-norm in inference form, as ATen evaluates it: alpha = w / sqrt(var + eps),
// y = x * alpha + (b - mean * alpha).  bn_weight == NULL: no normalisation (plain ReLU when relu != 0).
template <bool SPLIT>
__global__ void __launch_bounds__(256) bn_relu_split_kernel(const float* __restrict__ x, int C, int n8_per_row,
                                                            long long total8, const float* __restrict__ bw,
                                                            const float* __restrict__ bb, const float* __restrict__ bm,
                                                            const float* __restrict__ bv, float eps, int relu,
                                                            uint16_t* __restrict__ hi, uint16_t* __restrict__ lo) {
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total8; i += gridDim.x * 256ll) {
    const int c = static_cast<int>((i / n8_per_row) % C);
    float alpha = 1.f, shift = 0.f;
    if (bw != nullptr) {
      alpha = __ldg(bw + c) * (1.f / sqrtf(__ldg(bv + c) + eps));
      shift = __ldg(bb + c) - __ldg(bm + c) * alpha;
    }
    const float4 a = ld_stream_f4(x + i * 8), b = ld_stream_f4(x + i * 8 + 4);
    float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      v[j] = fmaf(v[j], alpha, shift);
      if (relu) v[j] = fmaxf(v[j], 0.f);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      h[j] = pack_bf16(v[2 * j], v[2 * j + 1]);
      l[j] = pack_bf16(v[2 * j] - bf16lo(h[j]), v[2 * j + 1] - bf16hi(h[j]));
    }
    reinterpret_cast<uint4*>(hi)[i] = make_uint4(h[0], h[1], h[2], h[3]);    // re-read by the GEMM / the head: keep in L2
    if (SPLIT) reinterpret_cast<uint4*>(lo)[i] = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

// fp32 [B,Cs,N] -> bf16 channel slice [c_off, c_off + Cs) of [B,Ctot,N] (torch.cat along channels + conversion)
__global__ void __launch_bounds__(256) concat_cast_kernel(const float* __restrict__ src, long long img8, long long total8,
                                                          long long out_img8, long long out_off8,
                                                          uint16_t* __restrict__ out) {
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total8; i += gridDim.x * 256ll) {
    const long long b = i / img8, r = i - b * img8;
    const float4 a = ld_stream_f4(src + i * 8), c = ld_stream_f4(src + i * 8 + 4);
    reinterpret_cast<uint4*>(out)[b * out_img8 + out_off8 + r] =
        make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(c.x, c.y), pack_bf16(c.z, c.w));
  }
}

__global__ void split_rect_kernel(const float* __restrict__ W, long long n, uint16_t* hi, uint16_t* lo) {
  const long long i = blockIdx.x * 256ll + threadIdx.x;
  if (i >= n) return;
  const float v = W[i];
  const uint16_t h = f32_to_bf16_rn(v);
  hi[i] = h;
  lo[i] = f32_to_bf16_rn(v - bf16_bits_to_f32(h));
}

template <int P>
static int launch_ln(const float* x, int B, int C, int N, const float* gamma, const float* beta, float eps, uint16_t* out,
                     cudaStream_t st) {
  const size_t smem = (static_cast<size_t>(C) * P + (LN_THREADS / P) * P + 2 * P) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(layernorm_tail_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  const long long ctas = static_cast<long long>(B) * ((N + P - 1) / P);
  layernorm_tail_kernel<P><<<static_cast<unsigned>(ctas), LN_THREADS, smem, st>>>(x, C, N, gamma, beta, eps, out);
  return SL_LAUNCH_RESULT();
}

}  // namespace tails
}  // namespace sl

extern "C" int sl_tail_layernorm(const float* x, int B, int C, int N, const float* gamma, const float* beta, float eps,
                                 uint16_t* feat_out, void* stream) {
  SL_CHECK_PTR(x); SL_CHECK_PTR(gamma); SL_CHECK_PTR(beta); SL_CHECK_PTR(feat_out);
  SL_CHECK_ARG(B >= 1 && C >= 1 && C <= 1536 && N >= 8 && N % 8 == 0);
  SL_CHECK_ARG(static_cast<long long>(B) * ((N + 31) / 32) < (1ll << 31));
  SL_CHECK_ALIGN(x, 16); SL_CHECK_ALIGN(feat_out, 16);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // 64-pixel tiles while four CTAs still fit on an SM; narrower tiles for wide heads
  // (rows of >= 128 bytes keep the DRAM bursts whole: two CTAs of <= 106 KB per SM for the wide heads)
  if (C <= 208) return sl::tails::launch_ln<64>(x, B, C, N, gamma, beta, eps, feat_out, st);
  if (C <= 832) return sl::tails::launch_ln<32>(x, B, C, N, gamma, beta, eps, feat_out, st);
  return sl::tails::launch_ln<16>(x, B, C, N, gamma, beta, eps, feat_out, st);
}

extern "C" int sl_tail_sum(const float* const* maps_host, int M, long long n, uint16_t* feat_out, void* stream) {
  SL_CHECK_PTR(maps_host); SL_CHECK_PTR(feat_out);
  SL_CHECK_ARG(M >= 1 && M <= sl::tails::SUM_MAX && n >= 8 && n % 8 == 0);
  SL_CHECK_ALIGN(feat_out, 16);
  sl::tails::SumPtrs ptrs{};
  for (int m = 0; m < M; ++m) {
    SL_CHECK_PTR(maps_host[m]);
    SL_CHECK_ALIGN(maps_host[m], 16);
    ptrs.p[m] = maps_host[m];
  }
  const long long n8 = n / 8;
  const long long want = (n8 + 255) / 256;
  const int grid = static_cast<int>(want < 8ll * sl::num_sms() ? want : 8ll * sl::num_sms());
  sl::tails::sum_tail_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(ptrs, M, n8, feat_out);
  return SL_LAUNCH_RESULT();
}

extern "C" int sl_tail_conv_prepare(const float* W, int Cout, int Cin, uint16_t* W_hi, uint16_t* W_lo, void* stream) {
  SL_CHECK_PTR(W); SL_CHECK_PTR(W_hi); SL_CHECK_PTR(W_lo);
  SL_CHECK_ARG(Cout >= 1 && Cin >= 1);
  const long long n = static_cast<long long>(Cout) * Cin;
  sl::tails::split_rect_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(W, n, W_hi,
                                                                                                                  W_lo);
  return SL_LAUNCH_RESULT();
}

extern "C" size_t sl_tail_bn_relu_conv_ws_bytes(int B, int Cin, int N) {
  return 2 * static_cast<size_t>(B) * Cin * N * sizeof(uint16_t) + 256;
}

extern "C" int sl_tail_bn_relu_conv(const float* x, int B, int Cin, int N, const float* bn_weight, const float* bn_bias,
                                    const float* bn_mean, const float* bn_var, float bn_eps, int relu,
                                    const uint16_t* W_hi, const uint16_t* W_lo, const float* bias, int Cout, void* ws,
                                    uint16_t* feat_out, void* stream) {
  SL_CHECK_PTR(x); SL_CHECK_PTR(W_hi); SL_CHECK_PTR(W_lo); SL_CHECK_PTR(ws); SL_CHECK_PTR(feat_out);
  if (bn_weight != nullptr) { SL_CHECK_PTR(bn_bias); SL_CHECK_PTR(bn_mean); SL_CHECK_PTR(bn_var); }
  SL_CHECK_ARG(B >= 1 && N >= 8 && N % 8 == 0 && Cin >= 8 && Cin % 8 == 0 && Cin <= 4096 && Cout >= 1 && Cout <= 4096);
  SL_CHECK_ARG(static_cast<long long>(B) * N < (1ll << 31));
  SL_CHECK_ALIGN(x, 16); SL_CHECK_ALIGN(ws, 128); SL_CHECK_ALIGN(W_hi, 16); SL_CHECK_ALIGN(W_lo, 16);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // Default: the fused kernel of tail_conv.cu (BN/ReLU/split applied to the A operand on the SM, no intermediate
  // planes).  SL_TAIL_FUSED=0, or a shape outside its range (Cin > 1024), takes the two-kernel path below.
  if (sl::env().tail_fused != 0) {
    const int frc = sl_tail_conv_fused_run(x, B, Cin, N, bn_weight, bn_bias, bn_mean, bn_var, bn_eps, relu, W_hi, W_lo, bias,
                                           Cout, feat_out, st);
    if (frc != SL_EINVAL) return frc;
  }
  // One split launch and one GEMM launch for the whole batch.  (Working through the batch in L2-sized groups of
  // images, so the hi/lo planes never leave the chip, was measured slower: 1 / 2 / 4 images per group 49.0 / 42.4 /
  // 36.1 us per PSPNet tile against 34.9 for the whole batch -- the persistent GEMM's partial last wave costs more
  // than the HBM round trip saves.)
  const long long plane = static_cast<long long>(B) * Cin * N;             // elements; * 2 bytes is a multiple of 128
  uint16_t* hi = static_cast<uint16_t*>(ws);
  uint16_t* lo = hi + plane;
  const long long total8 = plane / 8;
  const long long want = (total8 + 255) / 256;
  const int grid = static_cast<int>(want < 16ll * sl::num_sms() ? want : 16ll * sl::num_sms());
  sl::tails::bn_relu_split_kernel<true><<<grid, 256, 0, st>>>(x, Cin, N / 8, total8, bn_weight, bn_bias, bn_mean, bn_var, bn_eps,
                                                        relu, hi, lo);
  const int rc = SL_LAUNCH_RESULT();
  if (rc != 0) return rc;
  return sl_tail_gemm_run(hi, lo, B, Cin, N, W_hi, W_lo, bias, Cout, feat_out, st);
}

extern "C" int sl_tail_bn_relu(const float* x, int B, int C, int N, const float* bn_weight, const float* bn_bias,
                               const float* bn_mean, const float* bn_var, float bn_eps, int relu, uint16_t* feat_out,
                               void* stream) {
  SL_CHECK_PTR(x); SL_CHECK_PTR(feat_out);
  if (bn_weight != nullptr) { SL_CHECK_PTR(bn_bias); SL_CHECK_PTR(bn_mean); SL_CHECK_PTR(bn_var); }
  SL_CHECK_ARG(B >= 1 && C >= 1 && N >= 8 && N % 8 == 0);
  SL_CHECK_ALIGN(x, 16); SL_CHECK_ALIGN(feat_out, 16);
  const long long total8 = static_cast<long long>(B) * C * N / 8;
  const long long want = (total8 + 255) / 256;
  const int grid = static_cast<int>(want < 16ll * sl::num_sms() ? want : 16ll * sl::num_sms());
  sl::tails::bn_relu_split_kernel<false><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, C, N / 8, total8, bn_weight, bn_bias, bn_mean, bn_var, bn_eps, relu, feat_out, nullptr);
  return SL_LAUNCH_RESULT();
}

extern "C" int sl_tail_concat(const float* const* maps_host, const int* channels_host, int M, int B, int N,
                              uint16_t* feat_out, void* stream) {
  SL_CHECK_PTR(maps_host); SL_CHECK_PTR(channels_host); SL_CHECK_PTR(feat_out);
  SL_CHECK_ARG(M >= 1 && M <= 16 && B >= 1 && N >= 8 && N % 8 == 0);
  SL_CHECK_ALIGN(feat_out, 16);
  long long c_total = 0;
  for (int m = 0; m < M; ++m) {
    SL_CHECK_PTR(maps_host[m]); SL_CHECK_ALIGN(maps_host[m], 16);
    SL_CHECK_ARG(channels_host[m] >= 1);
    c_total += channels_host[m];
  }
  long long c_off = 0;
  for (int m = 0; m < M; ++m) {
    const long long img8 = static_cast<long long>(channels_host[m]) * N / 8, total8 = img8 * B;
    const long long want = (total8 + 255) / 256;
    const int grid = static_cast<int>(want < 16ll * sl::num_sms() ? want : 16ll * sl::num_sms());
    sl::tails::concat_cast_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(maps_host[m], img8, total8,
                                                                                       c_total * N / 8, c_off * N / 8, feat_out);
    const int rc = SL_LAUNCH_RESULT();
    if (rc != 0) return rc;
    c_off += channels_host[m];
  }
  return 0;
}
