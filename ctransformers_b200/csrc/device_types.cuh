// Device-side data layout of the quantized hot path (sm_100a).
//
// Weights are NOT kept in GGUF's array-of-blocks form.  At load time every 2-D weight is repacked,
// byte for byte (same total size, so the HBM roofline denominator is unchanged):
//   K-quants (Q4_K / Q5_K / Q6_K): the STREAM layout of stream.cuh — 16-row tiles, block-major, every (tile, block) a
//   contiguous 16-byte-aligned piece of 16 x {144,176,210} bytes that one cp.async.bulk moves into shared memory and whose
//   interior is ordered for conflict-free 16-byte shared-memory loads of the mma.sync operand fragments.
//   Other types: per-tensor planes so that every lane of a warp issues 16-byte-aligned, fully coalesced loads no
//   matter how odd the source block size is (Q4_0 = 18 B, Q8_0 = 34 B):
//
//   type   plane qs (per row)          plane qh (per row)   plane sc (per row)               plane d (per row)
//   Q4_0   nb x 16 B nibbles           -                    -                                 nb x fp16
//   Q5_0   nb x 16 B nibbles           nb x 4 B fifth bits  -                                 nb x fp16
//   Q8_0   nb x 32 B int8              -                    -                                 nb x fp16
//   F16    K x 2 B                     -                    -                                 -
//   F32    K x 4 B                     -                    -                                 -
//
// Block contents are exactly the reference's (k_quants.h:76-117, ggml.c:888-925); only their placement
// changes.  Activations are quantized on the fly to the reference's Q8_K / Q8_0 (bit-exact) into
// shared memory (struct ActView) and never touch HBM.
#pragma once
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace ctb {

enum : int { GT_F32 = 0, GT_F16 = 1, GT_Q4_0 = 2, GT_Q5_0 = 6, GT_Q8_0 = 8, GT_Q4_K = 12, GT_Q5_K = 13, GT_Q6_K = 14 };

struct DevMat {
  int type = -1;
  int K = 0, M = 0;     // K: contiguous (input) dim, M: rows (output features)
  int nb = 0;           // quant blocks per row
  const uint8_t* qs = nullptr;
  const uint8_t* qh = nullptr;
  const uint8_t* sc = nullptr;
  const uint16_t* d = nullptr;
  const uint8_t* st = nullptr;   // K-quants: the stream layout of stream.cuh (16-row tiles, block-major; qs/qh/sc/d stay null)
  size_t bytes = 0;     // total bytes of all planes (= GGUF tensor bytes)
};

__host__ __device__ inline bool type_is_kquant(int t) { return t == GT_Q4_K || t == GT_Q5_K || t == GT_Q6_K; }
// activation format each weight type is multiplied with (reference: type_traits vec_dot_type, ggml.c:1638-1808)
enum : int { ACT_Q8_K = 0, ACT_Q8_0 = 1, ACT_F16 = 2, ACT_F32 = 3 };
__host__ __device__ inline int act_format_for(int t) {
  if (type_is_kquant(t)) return ACT_Q8_K;
  if (t == GT_Q4_0 || t == GT_Q5_0 || t == GT_Q8_0) return ACT_Q8_0;
  if (t == GT_F16) return ACT_F16;
  return ACT_F32;
}

// fp16 bit pattern <-> float, IEEE RNE (same results as the F16C instructions of the AVX2 reference build)
__device__ __forceinline__ float h2f(uint16_t h) { return __half2float(__ushort_as_half(h)); }
__device__ __forceinline__ uint16_t f2h(float f) { return __half_as_ushort(__float2half_rn(f)); }

__device__ __forceinline__ int4 ldg_stream16(const void* p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
// read-only 16-byte load that keeps its place in program order (block headers: re-used by the 4 lanes of a row, L1 may keep them)
__device__ __forceinline__ int4 ldg_keep16(const void* p) {
  int4 r;
  asm volatile("ld.global.nc.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ int2 ldg_stream8(const void* p) {
  int2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}

// Programmatic dependent launch (sm_90+): a kernel lets its successor's CTAs start early (they prefetch weights and set up
// while this one drains) and the successor blocks in pdl_wait() until the whole predecessor grid has finished and its
// writes are visible.  Both are no-ops for a kernel launched without the programmatic-serialization attribute.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace ctb
