// Floor of one host-in-the-loop round on this box: launch a kernel that writes a result and a
// sequence flag to mapped host memory, spin on the flag, repeat. Prints microseconds per round for
// (a) an empty one-thread kernel, (b) a one-block kernel doing a dependent chain of 13 Montgomery
// products and a block reduction (the arithmetic depth of a late sumcheck round).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/round_latency tools/round_latency.cu
#include <chrono>
#include <cstdio>
#include <cuda_runtime.h>

#include "../spartan_parallel_b200/csrc/fq.cuh"
using namespace spg;

__global__ void k_empty(volatile unsigned long long *flag, unsigned long long seq) {
  __threadfence_system();
  *flag = seq;
}
__global__ void k_round(const fq *in, fq *res, volatile unsigned long long *flag, unsigned long long seq) {
  __shared__ fq sm[32];
  fq a = fq_load(in + threadIdx.x), b = fq_load(in + 128 + threadIdx.x);
  for (int k = 0; k < 13; k++) a = fq_mul(a, b);  // dependent chain
  a = fq_warp_sum(a);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x < 32) {
    fq v = threadIdx.x < 4 ? sm[threadIdx.x] : fq_zero();
    v = fq_warp_sum(v);
    if (threadIdx.x == 0) {
      fq_store(res, v);
      __threadfence_system();
      *flag = seq;
    }
  }
}

int main() {
  unsigned long long *h_flag, *d_flag;
  fq *h_res, *d_res, *d_in;
  cudaHostAlloc(&h_flag, 64, cudaHostAllocMapped);
  cudaHostGetDevicePointer(&d_flag, h_flag, 0);
  cudaHostAlloc(&h_res, 64, cudaHostAllocMapped);
  cudaHostGetDevicePointer(&d_res, h_res, 0);
  cudaMalloc(&d_in, 256 * sizeof(fq));
  cudaMemset(d_in, 1, 256 * sizeof(fq));
  cudaStream_t st;
  cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  *h_flag = 0;
  unsigned long long seq = 0;
  for (int mode = 0; mode < 2; mode++) {
    for (int rep = 0; rep < 3; rep++) {
      const int n = 2000;
      auto t0 = std::chrono::steady_clock::now();
      for (int i = 0; i < n; i++) {
        ++seq;
        if (mode == 0) k_empty<<<1, 1, 0, st>>>(h_flag ? d_flag : nullptr, seq);
        else k_round<<<1, 128, 0, st>>>(d_in, d_res, d_flag, seq);
        while (*(volatile unsigned long long *)h_flag != seq) {
        }
      }
      double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / n;
      printf("%s: %.2f us per round\n", mode == 0 ? "empty kernel + mapped flag" : "13 dependent products + block sum + mapped result", us);
    }
  }
  printf("cuda status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
