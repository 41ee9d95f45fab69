// GGUF array-of-blocks → device planes (layout table in device_types.cuh).  Shared by the engine's model
// upload and the op-level entry points.
#pragma once
#include "device_types.cuh"

namespace ctb {

struct PlaneSizes { size_t qs, qh, sc, d; };

inline PlaneSizes plane_sizes(int type, int M, int nb, size_t raw_bytes) {
  const size_t nblk = (size_t)M * nb;
  switch (type) {
    case GT_Q4_K: return {nblk * 128, 0, nblk * 16, 0};
    case GT_Q5_K: return {nblk * 128, nblk * 32, nblk * 16, 0};
    case GT_Q6_K: return {nblk * 128, nblk * 64, nblk * 16, nblk * 2};
    case GT_Q4_0: return {nblk * 16, 0, 0, nblk * 2};
    case GT_Q8_0: return {nblk * 32, 0, 0, nblk * 2};
    default: return {raw_bytes, 0, 0, 0};
  }
}

// GGUF array-of-blocks → planes (device_types.cuh), 2 bytes per thread-iteration.
// Lane-major permutation of a 32-byte-group structured plane.  The reference's AVX2 kernels leave, in int32 lane l (0..7)
// of their accumulator, the products of elements 4l..4l+3 of EVERY 32-element group (maddubs_epi16 + madd_epi16), and fold
// lane l into fp32 lane l once per block.  We give each of 8 GPU lanes one AVX lane, so GPU lane l needs 32-bit word l of each
// 32-byte group of a block; storing those words contiguously ("[l][group]") turns its per-block fetch into one 16-byte load.
//   Q4_K / Q5_K qs (128 B = 4 groups): word j*8+l  -> l*4+j
//   Q6_K ql (128 B: half jj, vector v): word jj*16+v*8+l -> l*4+jj*2+v ;  qh (64 B): word jj*8+l -> l*2+jj
static __global__ void k_repack(int type, const uint16_t* __restrict__ raw, size_t n_u16, uint16_t* qs, uint16_t* qh, uint16_t* sc, uint16_t* d) {
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n_u16; idx += (size_t)gridDim.x * blockDim.x) {
    const uint16_t v = raw[idx];
    switch (type) {
      case GT_Q4_K: case GT_Q5_K: {
        const int per = type == GT_Q4_K ? 72 : 88, qoff = type == GT_Q4_K ? 8 : 24;
        const size_t blk = idx / per; const int o = (int)(idx % per);
        if (o < 8) sc[blk * 8 + o] = v;
        else if (o < qoff) qh[blk * 16 + (o - 8)] = v;
        else {
          const int q16 = o - qoff, w = q16 >> 1, half = q16 & 1, j = w >> 3, l = w & 7;
          qs[blk * 64 + (l * 4 + j) * 2 + half] = v;
        }
      } break;
      case GT_Q6_K: {
        const size_t blk = idx / 105; const int o = (int)(idx % 105);
        if (o < 64) {
          const int w = o >> 1, half = o & 1, jj = w >> 4, vv = (w >> 3) & 1, l = w & 7;
          qs[blk * 64 + (l * 4 + jj * 2 + vv) * 2 + half] = v;
        } else if (o < 96) {
          const int q = o - 64, w = q >> 1, half = q & 1, jj = w >> 3, l = w & 7;
          qh[blk * 32 + (l * 2 + jj) * 2 + half] = v;
        } else if (o < 104) sc[blk * 8 + (o - 96)] = v;
        else d[blk] = v;
      } break;
      case GT_Q4_0: {
        const size_t blk = idx / 9; const int o = (int)(idx % 9);
        if (o == 0) d[blk] = v; else qs[blk * 8 + (o - 1)] = v;
      } break;
      case GT_Q8_0: {
        const size_t blk = idx / 17; const int o = (int)(idx % 17);
        if (o == 0) d[blk] = v; else qs[blk * 16 + (o - 1)] = v;
      } break;
      default: qs[idx] = v;
    }
  }
}

}  // namespace ctb
