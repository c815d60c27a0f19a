// sl_pop_fg_lite: the K foreground logits of the POP head (networks/pspnet_pop.py:108-109,114-115 + the alpha/beta
// collapse of the classifier, :150-157 / :178-182) by a kernel that can share an SM with the background-MLP pair
// kernel: 128-thread CTAs, <= 128 registers, NO shared memory, CUDA cores only.
//
// Why: sl_pop_fg_lowres is HBM-bound and streams the features once (0.09 ms per 32 PSPNet tiles), but it needs the
// whole SM (16 warps x 128 registers, 192 KB of TMA rings), so a sweep pays it in series with the tensor-bound
// background MLP, which leaves HBM 8 % busy and the FMA pipe idle for a millisecond.  This kernel computes the same
// projections slowly -- one CTA per SM next to the pair kernel's CTA, a few hundred GB/s -- which is all it takes to
// hide the foreground pass underneath sl_pop_bg_tc on a second stream (sweep.PipelinedTileEvaluator).
//
// It is the FFMA2 formulation of pop_fg.cu (K packed FMAs per feature pair, channels in ascending order per pixel, so
// the logits equal that kernel's bit for bit and the default mma.sync kernel's to ~1e-7) with the shared-memory
// machinery removed: a lane owns 8 consecutive pixels, loads them straight from global memory (16 bytes per channel,
// a warp covers 512 contiguous bytes) into a register ring FL_DEPTH channels deep, and reads the prototypes from the
// kernel-parameter constant bank (a [512][8] fp32 table passed by value: 16 KB of the 32 KB CUDA 12 allows), so a
// channel costs one global load, two constant loads, the bf16 unpacking and 4 K packed FMAs.
// An mma.sync variant of this kernel (registers + movmatrix, bit-identical to pop_fg_mma.cu) was built first and
// dropped: legacy HMMA shares the tensor pipe with tcgen05, so underneath the background MLP it ran 7x slower than
// alone and slowed the MLP by 20 % (profiles/r2b_overlap_probe.txt).
#include "common.cuh"

namespace sl {

namespace {

struct ChMap8 { int ch[8]; };
constexpr int FL_CMAX = 512;
struct FgLiteTable { float s[FL_CMAX][8]; };    // s[c][k]: prototype k (of this pass) at channel c

constexpr int FL_THREADS = 128;
constexpr int FL_PX = 256;                      // pixels per item (one warp, 8 per lane)
constexpr int FL_DEPTH = 8;                     // channels in flight per lane

__device__ __forceinline__ void fl_ffma2(float2& d, const float2 a, const float2 b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;"
      : "+l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<const unsigned long long&>(a)), "l"(reinterpret_cast<const unsigned long long&>(b)));
}

template <int KC>
__global__ void __launch_bounds__(FL_THREADS, 4)
pop_fg_lite_kernel(const uint16_t* __restrict__ feat, int B, int C, int N, const __grid_constant__ FgLiteTable tab,
                   const float* __restrict__ alpha, const float* __restrict__ beta, int k_base,
                   float* __restrict__ logits, int Ktot, ChMap8 map) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int items_per_image = (N + FL_PX - 1) / FL_PX;
  const long long n_items = static_cast<long long>(B) * items_per_image;
  const long long stride = static_cast<long long>(gridDim.x) * (FL_THREADS / 32);
  const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);

  for (long long item = static_cast<long long>(blockIdx.x) * (FL_THREADS / 32) + warp; item < n_items; item += stride) {
    const int b = static_cast<int>(item / items_per_image);
    const int n = static_cast<int>(item - static_cast<long long>(b) * items_per_image) * FL_PX + 8 * lane;
    const bool ok = n < N;                                       // N % 8 == 0: all eight pixels or none
    const uint16_t* src = feat + static_cast<size_t>(b) * C * N + n;
    uint4 ring[FL_DEPTH];
#pragma unroll
    for (int d = 0; d < FL_DEPTH; ++d) ring[d] = (ok && d < C) ? ld_stream_u4(src + static_cast<size_t>(d) * N) : zero4;
    float2 acc[KC][4];
#pragma unroll
    for (int k = 0; k < KC; ++k)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[k][j] = make_float2(0.f, 0.f);

    for (int c0 = 0; c0 < C; c0 += FL_DEPTH) {                   // C % 8 == 0 == FL_DEPTH: whole trips
#pragma unroll
      for (int d = 0; d < FL_DEPTH; ++d) {
        const int c = c0 + d;
        const uint4 v = ring[d];
        if (c + FL_DEPTH < C && ok) ring[d] = ld_stream_u4(src + static_cast<size_t>(c + FL_DEPTH) * N);
        const float2 x[4] = {make_float2(bf16lo(v.x), bf16hi(v.x)), make_float2(bf16lo(v.y), bf16hi(v.y)),
                             make_float2(bf16lo(v.z), bf16hi(v.z)), make_float2(bf16lo(v.w), bf16hi(v.w))};
        const float4 s03 = *reinterpret_cast<const float4*>(&tab.s[c][0]);
        const float4 s47 = *reinterpret_cast<const float4*>(&tab.s[c][4]);
        const float sv[8] = {s03.x, s03.y, s03.z, s03.w, s47.x, s47.y, s47.z, s47.w};
#pragma unroll
        for (int k = 0; k < KC; ++k) {
          const float2 s2 = make_float2(sv[k], sv[k]);
#pragma unroll
          for (int j = 0; j < 4; ++j) fl_ffma2(acc[k][j], s2, x[j]);
        }
      }
    }

    if (ok) {
#pragma unroll
      for (int k = 0; k < KC; ++k) {
        const float al = __ldg(alpha + k_base + k), bt = __ldg(beta + k_base + k);
        float o[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float p0 = acc[k][j].x, p1 = acc[k][j].y;
          o[2 * j] = p0 >= 0.f ? p0 * al : -p0 * bt;
          o[2 * j + 1] = p1 >= 0.f ? p1 * al : -p1 * bt;
        }
        float4* d4 = reinterpret_cast<float4*>(logits + (static_cast<size_t>(b) * Ktot + map.ch[k]) * N + n);
        d4[0] = make_float4(o[0], o[1], o[2], o[3]);
        d4[1] = make_float4(o[4], o[5], o[6], o[7]);
      }
    }
  }
}

}  // namespace

}  // namespace sl

// ws: the prototypes transposed to [C][8 * passes] fp32 on the HOST side of the call boundary?  No -- the table travels
// as a kernel parameter, so sl_pop_fg_lite needs it in host memory: sl_pop_fg_lite_prepare copies s_hat to the
// caller-owned pinned/pageable host buffer `ws` ([passes][512][8] fp32, zero-padded) with one stream-ordered D2H copy;
// the caller must have synchronised with `stream` before the first sl_pop_fg_lite that uses it (PopHead.refresh does).
extern "C" size_t sl_pop_fg_lite_ws_bytes(int K, int C) {
  if (K < 1 || C < 1 || C > sl::FL_CMAX) return 0;
  return static_cast<size_t>(K) * C * sizeof(float);           // a host copy of s_hat [K][C]
}

extern "C" int sl_pop_fg_lite_prepare(const float* s_hat, int K, int C, void* ws_host, void* stream) {
  SL_CHECK_PTR(s_hat); SL_CHECK_PTR(ws_host);
  SL_CHECK_ARG(K >= 1 && K < SL_MAX_CLASSES && C >= 8 && C % 8 == 0 && C <= sl::FL_CMAX);
  const cudaError_t e = cudaMemcpyAsync(ws_host, s_hat, static_cast<size_t>(K) * C * sizeof(float), cudaMemcpyDeviceToHost,
                                        static_cast<cudaStream_t>(stream));
  return static_cast<int>(e);
}

extern "C" int sl_pop_fg_lite(const uint16_t* feat, int B, int C, int N, const void* ws_host, const float* alpha,
                              const float* beta, int K, float* logits, int Ktot, const int* ch_map_host, void* stream) {
  SL_CHECK_PTR(feat); SL_CHECK_PTR(ws_host); SL_CHECK_PTR(alpha); SL_CHECK_PTR(beta); SL_CHECK_PTR(logits);
  SL_CHECK_PTR(ch_map_host);
  SL_CHECK_ARG(B >= 1 && K >= 1 && K < SL_MAX_CLASSES && Ktot >= K && Ktot <= SL_MAX_CLASSES);
  SL_CHECK_ARG(C >= 8 && C % 8 == 0 && C <= sl::FL_CMAX && N >= 8 && N % 8 == 0);
  SL_CHECK_ALIGN(feat, 16); SL_CHECK_ALIGN(logits, 16);
  for (int k = 0; k < K; ++k) SL_CHECK_ARG(ch_map_host[k] >= 0 && ch_map_host[k] < Ktot);
  const float* s_host = static_cast<const float*>(ws_host);
  const long long items = static_cast<long long>(B) * ((N + sl::FL_PX - 1) / sl::FL_PX);
  const long long want = (items + 3) / 4, cap = 4ll * sl::num_sms();
  const int grid = static_cast<int>(want < cap ? want : cap);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  static thread_local sl::FgLiteTable tab;                       // 16 KB: built per launch, copied into the parameter buffer
  for (int k_base = 0; k_base < K; k_base += 8) {
    const int kc = K - k_base < 8 ? K - k_base : 8;
    sl::ChMap8 map;
    for (int k = 0; k < 8; ++k) map.ch[k] = k < kc ? ch_map_host[k_base + k] : 0;
    for (int c = 0; c < sl::FL_CMAX; ++c)
      for (int k = 0; k < 8; ++k) tab.s[c][k] = (c < C && k < kc) ? s_host[static_cast<size_t>(k_base + k) * C + c] : 0.f;
#define SL_FL_CASE(KC) case KC: sl::pop_fg_lite_kernel<KC><<<grid, sl::FL_THREADS, 0, st>>>( \
        feat, B, C, N, tab, alpha, beta, k_base, logits, Ktot, map); break
    switch (kc) {
      SL_FL_CASE(1); SL_FL_CASE(2); SL_FL_CASE(3); SL_FL_CASE(4); SL_FL_CASE(5); SL_FL_CASE(6); SL_FL_CASE(7); SL_FL_CASE(8);
    }
#undef SL_FL_CASE
    const int rc = SL_LAUNCH_RESULT();
    if (rc != 0) return rc;
  }
  return SL_OK;
}
