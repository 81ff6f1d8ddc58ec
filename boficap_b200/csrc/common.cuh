// Shared helpers for the BoFi sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <type_traits>
#include <stdint.h>
#include <math.h>

namespace bofi {

constexpr int kD = 512;        // d_model of every uic_sd* config (checked in bofi_create)
constexpr int kHeadDim = 64;   // d_model / heads
constexpr int kMaxKeys = 128;  // regions (<= max_boxes = 100) or slots (<= 22) a query can attend to

typedef __nv_bfloat16 bf16;

template <typename T> __device__ __forceinline__ float to_float(T v);
template <> __device__ __forceinline__ float to_float<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_float<bf16>(bf16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_float(float v);
template <> __device__ __forceinline__ float from_float<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_float<bf16>(float v) { return __float2bfloat16_rn(v); }

// 4 consecutive elements <-> float4 (16 B for fp32, 8 B for bf16)
__device__ __forceinline__ float4 load4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 load4(const bf16* p) {
  uint2 raw = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 lo = *reinterpret_cast<__nv_bfloat162*>(&raw.x);
  __nv_bfloat162 hi = *reinterpret_cast<__nv_bfloat162*>(&raw.y);
  float2 a = __bfloat1622float2(lo), b = __bfloat1622float2(hi);
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void store4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void store4(bf16* p, float4 v) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
  uint2 raw;
  raw.x = *reinterpret_cast<uint32_t*>(&lo);
  raw.y = *reinterpret_cast<uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(p) = raw;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Programmatic dependent launch (every kernel of the path is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization): let the next kernel's CTAs be scheduled as soon as
// this grid has started, then block until everything this kernel depends on has completed and is visible.
// EVERY kernel must execute pdl_wait() before its first global access (also before an early return):
// completion of a grid must imply completion of its predecessors.
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() { pdl_launch(); pdl_wait(); }

// Host side: every kernel of the library goes through launch_k, which attaches the programmatic-dependent-launch
// attribute (BOFI_PDL=0 turns it off).  Works in eager streams and under stream capture (programmatic graph edges).
inline bool& pdl_enabled() {
  static bool on = true;
  return on;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// cudaFuncSetAttribute and the SM count are per DEVICE: one-time-setup caches are kept per device ordinal, so a second
// device in the same process gets its own dynamic shared memory opt-in.
template <typename V>
struct PerDevice {
  V v[64] = {};
  V& get() {
    int d = 0;
    cudaGetDevice(&d);
    return v[d & 63];
  }
};

// Dropout (training path only).  Counter-based: the keep/drop decision of element `idx` of dropout site `key` is a pure
// function (murmur3 finaliser), so the backward pass regenerates the mask instead of storing it, and the CPU oracle can
// reproduce it exactly (oracle/bofi_oracle.py:drop_mask).  thresh = p * 2^32 (0 = disabled), scale = 1 / (1 - p).
struct Drop {
  uint32_t key = 0, thresh = 0;
  float scale = 1.f;
};
__host__ __device__ __forceinline__ uint32_t drop_hash(uint32_t key, uint32_t idx) {
  uint32_t h = (idx * 0x9E3779B1u) ^ key;
  h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
  return h;
}
__device__ __forceinline__ float drop_mul(const Drop& d, uint32_t idx) {      // 0 or 1/(1-p); 1 when disabled
  if (d.thresh == 0) return 1.f;
  return drop_hash(d.key, idx) >= d.thresh ? d.scale : 0.f;
}
// element index of an attention probability: (global query row, head, key)
__device__ __forceinline__ uint32_t att_idx(int qrow, int head, int key) { return ((uint32_t)qrow * 8u + (uint32_t)head) * 128u + (uint32_t)key; }

// Multinomial sampling of CaptionModel.sample_next_word (CaptionModel.py:403-431: logits / temperature, NaN -> -10,
// Categorical(...).sample()) as a Gumbel-max: argmax_v (z_v / T + g_v), g_v = -log(-log(u_v)), u_v from the counter-based
// hash of (key, row, v).  Same distribution as the reference, the library's own random stream (seeded by the caller).
struct Sampler {
  uint32_t key = 0;
  int enabled = 0;          // 0 = greedy
  float inv_temp = 1.f;
};
__device__ __forceinline__ float gumbel_score(const Sampler& sp, float z, int row, int v) {
  const float zz = (z != z) ? -10.f : z * sp.inv_temp;
  const float u = ((float)drop_hash(sp.key, (uint32_t)row * 16384u + (uint32_t)v) + 0.5f) * (1.0f / 4294967296.0f);
  return zz - __logf(-__logf(u));
}

// Work of a bounding step is skipped once every row has finished (device-side `break` of
// TransformerModel.py:1869-1870): kernels of the step read the live-row counter first.
__device__ __forceinline__ bool step_is_dead(const int* live_rows) {
  return live_rows != nullptr && *live_rows == 0;
}

// t / d from a per-row reciprocal: q = t * inv, then one FMA residual step (r = t - q d exactly, q += r * inv).  Equals
// the IEEE quotient except in rare double-rounding cases, at 3 instructions instead of the ~10 of a true division (16 of
// them per lane had the LayerNorm kernel issue-bound).
__device__ __forceinline__ float ln_div(float t, float d, float inv) {
  const float q = t * inv;
  const float r = fmaf(-q, d, t);
  return fmaf(r, inv, q);
}


}  // namespace bofi
