// Micro-benchmark: which HBM access shapes reach the copy bandwidth on B200?  (Design input for matvec.cuh.)
//   mode 0: linear — every warp instruction reads 512 contiguous bytes, warps stride the buffer
//   mode 1: "row tile" — a warp owns RW rows of ROWB bytes; per step each row's LPR lanes read (LPR*16) contiguous bytes
//           RW=8,LPR=4: the lane-owned mat-vec shape; RW=2,LPR=16; RW=1,LPR=32
// D = loads in flight per lane.  Prints GB/s.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ __forceinline__ int4 ldg16(const void* p) { int4 r; asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p)); return r; }
template <int D>
__global__ void k_linear(const char* buf, size_t bytes, int* sink) {
  const size_t nchunk = bytes / 16, stride = (size_t)gridDim.x * blockDim.x;
  int acc = 0;
  for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < nchunk; c += stride * D) {
    int4 v[D];
#pragma unroll
    for (int i = 0; i < D; i++) { size_t cc = c + i * stride; v[i] = ldg16(buf + (cc < nchunk ? cc : c) * 16); }
#pragma unroll
    for (int i = 0; i < D; i++) acc += v[i].x ^ v[i].y ^ v[i].z ^ v[i].w;
  }
  if (acc == 0x12345678) *sink = acc;
}
template <int RW, int D>
__global__ void k_rows(const char* buf, int nrows, int rowb, int* sink) {
  constexpr int LPR = 32 / RW;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = (gridDim.x * blockDim.x) >> 5, lane = threadIdx.x & 31;
  const int g = lane / LPR, t = lane % LPR;
  const int steps = rowb / (LPR * 16);
  int acc = 0;
  for (int tile = warp; tile * RW < nrows; tile += nwarp) {
    const char* rp = buf + (size_t)(tile * RW + g) * rowb + t * 16;
    for (int s = 0; s < steps; s += D) {
      int4 v[D];
#pragma unroll
      for (int i = 0; i < D; i++) v[i] = ldg16(rp + (size_t)min(s + i, steps - 1) * (LPR * 16));
#pragma unroll
      for (int i = 0; i < D; i++) acc += v[i].x ^ v[i].y ^ v[i].z ^ v[i].w;
    }
  }
  if (acc == 0x12345678) *sink = acc;
}
template <typename F> float timeit(F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(a); for (int i = 0; i < 5; i++) f(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms / 5;
}
int main() {
  const int rowb = 2304, nrows = 4096 * 12 * 8;   // 906 MB of "Q4_K rows" (K = 4096)
  const size_t bytes = (size_t)rowb * nrows;
  char* buf; int* sink; cudaMalloc(&buf, bytes); cudaMalloc(&sink, 4); cudaMemset(buf, 1, bytes);
  int sm = 148;
#define RUN(name, ...) { float ms = timeit([&] { __VA_ARGS__; }); printf("%-44s %8.1f GB/s\n", name, bytes / ms / 1e6); }
  RUN("linear D=4, 148x512", (k_linear<4><<<sm, 512>>>(buf, bytes, sink)));
  RUN("linear D=8, 148x512", (k_linear<8><<<sm, 512>>>(buf, bytes, sink)));
  RUN("linear D=4, 148x1024", (k_linear<4><<<sm, 1024>>>(buf, bytes, sink)));
  RUN("linear D=8, 296x1024", (k_linear<8><<<sm * 2, 1024>>>(buf, bytes, sink)));
  RUN("rows RW=8 (64B/row) D=4, 148x512", (k_rows<8, 4><<<sm, 512>>>(buf, nrows, rowb, sink)));
  RUN("rows RW=8 (64B/row) D=8, 148x512", (k_rows<8, 8><<<sm, 512>>>(buf, nrows, rowb, sink)));
  RUN("rows RW=8 (64B/row) D=8, 148x1024", (k_rows<8, 8><<<sm, 1024>>>(buf, nrows, rowb, sink)));
  RUN("rows RW=4 (128B/row) D=4, 148x512", (k_rows<4, 4><<<sm, 512>>>(buf, nrows, rowb, sink)));
  RUN("rows RW=4 (128B/row) D=8, 148x512", (k_rows<4, 8><<<sm, 512>>>(buf, nrows, rowb, sink)));
  RUN("rows RW=2 (256B/row) D=4, 148x512", (k_rows<2, 4><<<sm, 512>>>(buf, nrows, rowb, sink)));
  RUN("rows RW=1 (512B/row) D=4, 148x512", (k_rows<1, 4><<<sm, 512>>>(buf, nrows, rowb, sink)));
  RUN("rows RW=1 (512B/row) D=8, 148x1024", (k_rows<1, 8><<<sm, 1024>>>(buf, nrows, rowb, sink)));
  RUN("rows RW=8 D=4, 148x256", (k_rows<8, 4><<<sm, 256>>>(buf, nrows, rowb, sink)));
  RUN("rows RW=8 D=2, 148x512", (k_rows<8, 2><<<sm, 512>>>(buf, nrows, rowb, sink)));
  RUN("rows RW=8 D=1, 148x512", (k_rows<8, 1><<<sm, 512>>>(buf, nrows, rowb, sink)));
  return 0;
}
