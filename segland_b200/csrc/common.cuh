// Shared device/host helpers for libsegland_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <cstdlib>
#include "../../include/segland_b200.h"

#define SL_CHECK_PTR(p) do { if ((p) == nullptr) return SL_ENULL; } while (0)
#define SL_CHECK_ARG(c) do { if (!(c)) return SL_EINVAL; } while (0)
#define SL_CHECK_ALIGN(p, a) do { if ((reinterpret_cast<uintptr_t>(p) % (a)) != 0) return SL_EALIGN; } while (0)
// cudaPeekAtLastError keeps sticky-error semantics with the caller; launch-config errors surface here.
#define SL_LAUNCH_RESULT() static_cast<int>(cudaGetLastError())

namespace sl {

// SM count of the caller's current device, queried once per device ordinal (B200: 148; a MIG slice or another
// sm_100 SKU has fewer).  Persistent grids and per-CTA workspaces are sized from it, never from a constant.
inline int num_sms() {
  static int cache[64] = {};                   // benign race: every thread stores the same value
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
  int n = cache[dev];
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) n = 148;
    cache[dev] = n;
  }
  return n;
}

// Environment switches (A/B measurements and debugging only; see include/segland_b200.h).  They are read once per
// process -- a launch never calls getenv -- and again only when the caller asks (sl_env_reload).
struct Env {
  int tc_pair;          // SL_TC_PAIR       (-1 = unset)
  int tc_small;         // SL_TC_SMALL      (-1 = unset)
  int tc_debug;         // SL_TC_DEBUG      (0)
  int prep_split;       // SL_PREP_SPLIT    (0)
  int post_fused_cm;    // SL_POST_FUSED_CM (-1 = unset)
  int post_prune;       // SL_POST_PRUNE    (0)
  int post_regs;        // SL_POST_REGS     (1: register-resident source-row intervals for the prediction-only path)
  int tail_fused;       // SL_TAIL_FUSED    (1)
  int fg_mma;           // SL_FG_MMA        (1: mma.sync projections; 0: the FFMA2 kernel)
  long long small_dbg;  // SL_SMALL_DBG     (0)
};
const Env& env();       // api.cu

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// 128-bit streaming load (read-once data: bypass L1 allocation).
__device__ __forceinline__ uint4 ld_stream_u4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float4 ld_stream_f4(const void* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

// bf16 pair (packed in a u32, low half = first element) -> two fp32 (exact).
__device__ __forceinline__ float bf16lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

__device__ __forceinline__ uint16_t f32_to_bf16_rn(float f) {
  return __bfloat16_as_ushort(__float2bfloat16_rn(f));
}
__device__ __forceinline__ float bf16_bits_to_f32(uint16_t b) { return __uint_as_float(static_cast<uint32_t>(b) << 16); }

// align_corners=True source coordinate, exactly as ATen computes it
// (area_pixel_compute_scale / area_pixel_compute_source_index in UpSample.h):
//   scale = (in-1)/(out-1) in fp32 (0 when out == 1), src = scale * dst.
struct SrcCoord { int i0; int step; float l0; float l1; };
__device__ __forceinline__ SrcCoord src_coord(float scale, int dst, int in_size) {
  const float r = scale * static_cast<float>(dst);
  int i0 = static_cast<int>(r);
  if (i0 > in_size - 1) i0 = in_size - 1;
  SrcCoord c;
  c.i0 = i0;
  c.step = (i0 < in_size - 1) ? 1 : 0;
  float l1 = r - static_cast<float>(i0);
  l1 = fminf(fmaxf(l1, 0.f), 1.f);
  c.l1 = l1;
  c.l0 = 1.f - l1;
  return c;
}
__host__ __device__ __forceinline__ float ac_scale(int in_size, int out_size) {
  return out_size > 1 ? static_cast<float>(in_size - 1) / static_cast<float>(out_size - 1) : 0.f;
}

// Block-private confusion histogram update with same-bin aggregation: spatially coherent
// label/pred maps put most of a warp in one bin; lane 0's bin is added once with a popcount,
// stragglers fall back to individual shared-memory atomics.
__device__ __forceinline__ void hist_add_warp(unsigned int* hist, int bin, bool valid) {
  const unsigned full = 0xffffffffu;
  const unsigned vmask = __ballot_sync(full, valid);
  if (vmask == 0) return;
  const int leader = __ffs(vmask) - 1;
  const int lbin = __shfl_sync(full, bin, leader);
  const unsigned same = __ballot_sync(full, valid && bin == lbin);
  if ((threadIdx.x & 31) == leader) atomicAdd(&hist[lbin], __popc(same));
  if (valid && bin != lbin) atomicAdd(&hist[bin], 1u);
}

}  // namespace sl
