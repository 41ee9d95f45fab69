// Micro-benchmarks for the round-2 restructuring of the decode step ("Budget for the 70 % target"):
//   1. grid-wide barrier among one persistent CTA per SM (atomic arrive + spin on a generation word in L2)
//   2. kernel boundary inside a CUDA graph, with and without programmatic dependent launch
//   3. thread-block-cluster barrier and a DSMEM broadcast of a 4 KB quantized activation vector to the cluster
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o sync_bench tools/sync_bench.cu ; run on one B200.
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ int ld_acquire(const int* p) { int v; asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }

// ---- 1. software grid barrier: arrive counter + generation, one CTA per SM (all resident)
__global__ void __launch_bounds__(512, 1) k_grid_barrier(int* arrive, int* gen, int rounds, unsigned long long* out) {
  const int n = gridDim.x;
  unsigned long long t0 = 0;
  for (int r = 0; r < rounds; r++) {
    if (r == rounds / 4 && threadIdx.x == 0) t0 = gtime();
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      const int g = ld_acquire(gen);
      if (atomicAdd(arrive, 1) == n - 1) { atomicExch(arrive, 0); __threadfence(); atomicAdd(gen, 1); }
      else while (ld_acquire(gen) == g) { }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = gtime() - t0;
}

// ---- 2. kernel boundary: a chain of tiny dependent kernels in a graph (each reads what the previous wrote)
__global__ void __launch_bounds__(512, 1) k_link(const float* in, float* out, int pdl) {
  if (pdl) asm volatile("griddepcontrol.launch_dependents;");
  if (pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  out[i] = in[i] + 1.f;
}

// ---- 3. cluster: barrier latency and DSMEM broadcast of 4 KB from every CTA's quarter to all 4 CTAs
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(512, 1) k_cluster(int rounds, unsigned long long* out) {
  __shared__ __align__(16) unsigned char stage[4096];
  cg::cluster_group cl = cg::this_cluster();
  const unsigned rank = cl.block_rank();
  unsigned long long t0 = 0, t1 = 0;
  cl.sync();
  if (threadIdx.x == 0) t0 = gtime();
  for (int r = 0; r < rounds; r++) cl.sync();
  if (threadIdx.x == 0) t1 = gtime();
  // broadcast: every CTA writes its 1 KB quarter into all 4 CTAs' stage buffers (uint4 per thread for 64 threads)
  unsigned long long t2 = 0, t3 = 0;
  cl.sync();
  if (threadIdx.x == 0) t2 = gtime();
  for (int r = 0; r < rounds; r++) {
    if (threadIdx.x < 64) {
      const uint4 v = make_uint4(r, rank, threadIdx.x, 7);
      for (unsigned peer = 0; peer < 4; peer++) {
        uint4* dst = (uint4*)cl.map_shared_rank(stage, peer) + rank * 64 + threadIdx.x;
        *dst = v;
      }
    }
    cl.sync();
  }
  if (threadIdx.x == 0) t3 = gtime();
  if (threadIdx.x == 0 && blockIdx.x == 0) { out[1] = t1 - t0; out[2] = t3 - t2; out[3] = stage[5]; }
}

int main() {
  int dev = 0, n_sm = 0;
  CK(cudaSetDevice(dev));
  CK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  int *arrive, *gen; unsigned long long* out;
  CK(cudaMalloc(&arrive, 4)); CK(cudaMalloc(&gen, 4)); CK(cudaMalloc(&out, 64));
  CK(cudaMemset(arrive, 0, 4)); CK(cudaMemset(gen, 0, 4)); CK(cudaMemset(out, 0, 64));
  const int rounds = 4000;
  k_grid_barrier<<<n_sm, 512>>>(arrive, gen, rounds, out);
  CK(cudaDeviceSynchronize());
  unsigned long long h[8];
  CK(cudaMemcpy(h, out, 64, cudaMemcpyDeviceToHost));
  printf("grid barrier (%d CTAs x 512 threads): %.3f us per barrier\n", n_sm, h[0] / 1e3 / (rounds - rounds / 4));

  float *a, *b;
  const int n = n_sm * 512;
  CK(cudaMalloc(&a, n * 4)); CK(cudaMalloc(&b, n * 4)); CK(cudaMemset(a, 0, n * 4));
  cudaStream_t st; CK(cudaStreamCreate(&st));
  for (int pdl = 0; pdl < 2; pdl++) {
    cudaGraph_t g; cudaGraphExec_t ex;
    const int links = 200;
    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    for (int i = 0; i < links; i++) {
      cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(n_sm); cfg.blockDim = dim3(512); cfg.stream = st;
      cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = at; cfg.numAttrs = pdl;
      CK(cudaLaunchKernelEx(&cfg, k_link, (const float*)((i & 1) ? b : a), (i & 1) ? a : b, pdl));
    }
    CK(cudaStreamEndCapture(st, &g));
    CK(cudaGraphInstantiate(&ex, g, 0));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaGraphLaunch(ex, st));
    CK(cudaEventRecord(e0, st));
    for (int r = 0; r < 10; r++) CK(cudaGraphLaunch(ex, st));
    CK(cudaEventRecord(e1, st));
    CK(cudaStreamSynchronize(st));
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    printf("kernel boundary in a graph, %s: %.3f us per dependent launch\n", pdl ? "programmatic dependent launch" : "plain", 1e3 * ms / (10 * links));
    cudaGraphExecDestroy(ex); cudaGraphDestroy(g);
  }

  const int crounds = 2000;
  k_cluster<<<(n_sm / 4) * 4, 512>>>(crounds, out);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(h, out, 64, cudaMemcpyDeviceToHost));
  printf("cluster of 4: barrier %.3f us; 4 KB DSMEM broadcast + barrier %.3f us\n", h[1] / 1e3 / crounds, h[2] / 1e3 / crounds);
  return 0;
}
