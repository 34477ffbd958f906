// Integer-pipe roofline microbenchmark for B200:
//   (a) mad.wide.u32 into a 64-bit accumulator, 8 independent chains per thread. ptxas does NOT
//       keep this fused on sm_100a: it emits IMAD.WIDE.U32 (zero addend) + IADD3 + IADD3.X
//       (cuobjdump -sass build/imad_peak), so this is the rate of that three-instruction form,
//   (b) carry-chained IMAD.WIDE.U32.X (mad.lo.cc / madc.hi.cc pairs), which ptxas does fuse:
//       the form fq_mul_lazy is made of, and the multiply rate the roofline is built on,
//   (c) in-register 256-bit Montgomery products/s with the library's fq_mul_lazy / fq_mul.
// Prints one JSON object. Build: make tools ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>
#include "../spartan_parallel_b200/csrc/fq.cuh"
using namespace spg;

template <int ILP>
__global__ void k_imad_wide(unsigned long long *out, unsigned int a, unsigned int b, int iters) {
  unsigned long long acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) acc[i] = threadIdx.x + i;
  unsigned int y = b + threadIdx.x;
  // the multiplicand of chain i is the low word of its own accumulator: every product is distinct
  // (ILP independent chains per thread, so the pipe and not the latency is measured), so
  // ptxas cannot hoist one common x*y out of the loop and leave only 64-bit additions behind
  // (which is what the first version of this kernel measured).
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++)
      asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"((unsigned int)acc[i] | a), "r"(y));
  }
  unsigned long long s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s ^= acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// 8-wide carry chain per step, 4 independent chains
__global__ void k_imad_carry(unsigned int *out, unsigned int a, unsigned int b, int iters) {
  unsigned int t[4][9];
#pragma unroll
  for (int c = 0; c < 4; c++)
#pragma unroll
    for (int i = 0; i < 9; i++) t[c][i] = threadIdx.x + i + c;
  unsigned int x0 = a + threadIdx.x, x1 = a ^ 0x1234567u, x2 = a * 3u, x3 = a + 77u, y = b;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int c = 0; c < 4; c++)
      asm volatile(
          "mad.lo.cc.u32   %0, %9,  %13, %0;\n\t"
          "madc.hi.cc.u32  %1, %9,  %13, %1;\n\t"
          "madc.lo.cc.u32  %2, %10, %13, %2;\n\t"
          "madc.hi.cc.u32  %3, %10, %13, %3;\n\t"
          "madc.lo.cc.u32  %4, %11, %13, %4;\n\t"
          "madc.hi.cc.u32  %5, %11, %13, %5;\n\t"
          "madc.lo.cc.u32  %6, %12, %13, %6;\n\t"
          "madc.hi.cc.u32  %7, %12, %13, %7;\n\t"
          "addc.u32        %8, %8, 0;\n\t"
          : "+r"(t[c][0]), "+r"(t[c][1]), "+r"(t[c][2]), "+r"(t[c][3]), "+r"(t[c][4]), "+r"(t[c][5]),
            "+r"(t[c][6]), "+r"(t[c][7]), "+r"(t[c][8])
          : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(y));
  }
  unsigned int s = 0;
#pragma unroll
  for (int c = 0; c < 4; c++)
#pragma unroll
    for (int i = 0; i < 9; i++) s ^= t[c][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP, bool CANON>
__global__ void k_modmul(fq *out, fq seed, int iters) {
  fq x[ILP], y = seed;
#pragma unroll
  for (int i = 0; i < ILP; i++) {
    x[i] = seed;
    x[i].v[0] ^= (threadIdx.x + 131 * i);
    x[i].v[7] &= 0x0fffffffu;
  }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = CANON ? fq_mul(x[i], y) : fq_mul_lazy(x[i], y);
  }
  fq s = x[0];
#pragma unroll
  for (int i = 1; i < ILP; i++)
#pragma unroll
    for (int k = 0; k < 8; k++) s.v[k] ^= x[i].v[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// 8-limb add chains (IADD3 + 7 IADD3.X), 4 independent chains per thread: the ALU-pipe side of the field code
__global__ void k_add_carry(unsigned int *out, unsigned int a, int iters) {
  unsigned int t[4][8], x[8];
#pragma unroll
  for (int i = 0; i < 8; i++) x[i] = a * (i + 3) + threadIdx.x;
#pragma unroll
  for (int c = 0; c < 4; c++)
#pragma unroll
    for (int i = 0; i < 8; i++) t[c][i] = threadIdx.x + i + c;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int c = 0; c < 4; c++)
      asm volatile(
          "add.cc.u32  %0, %0, %8;\n\t"
          "addc.cc.u32 %1, %1, %9;\n\t"
          "addc.cc.u32 %2, %2, %10;\n\t"
          "addc.cc.u32 %3, %3, %11;\n\t"
          "addc.cc.u32 %4, %4, %12;\n\t"
          "addc.cc.u32 %5, %5, %13;\n\t"
          "addc.cc.u32 %6, %6, %14;\n\t"
          "addc.u32    %7, %7, %15;\n\t"
          : "+r"(t[c][0]), "+r"(t[c][1]), "+r"(t[c][2]), "+r"(t[c][3]), "+r"(t[c][4]), "+r"(t[c][5]), "+r"(t[c][6]), "+r"(t[c][7])
          : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(x[4]), "r"(x[5]), "r"(x[6]), "r"(x[7]));
  }
  unsigned int s = 0;
#pragma unroll
  for (int c = 0; c < 4; c++)
#pragma unroll
    for (int i = 0; i < 8; i++) s ^= t[c][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// The bind of the sumcheck kernels, lo + r * (hi - lo), in registers. V = 0: as the kernels do it today
// (conditional corrections after every add / sub, canonical result); V = 1: the lean form, tables kept in
// [0, 2q): d = hi - lo + 2q without a condition, one conditional subtraction of 2q at the end.
template <int V>
__device__ __forceinline__ fq bind_form(const fq &lo, const fq &hi, const fq &r) {
  if (V == 0) return fq_canon(fq_add_lazy(lo, fq_mul_lazy(r, fq_sub_lazy(hi, lo))));
  fq d;
  asm("{\n\t"
      "sub.cc.u32  %0, %8,  %16;\n\t"
      "subc.cc.u32 %1, %9,  %17;\n\t"
      "subc.cc.u32 %2, %10, %18;\n\t"
      "subc.cc.u32 %3, %11, %19;\n\t"
      "subc.cc.u32 %4, %12, %20;\n\t"
      "subc.cc.u32 %5, %13, %21;\n\t"
      "subc.cc.u32 %6, %14, %22;\n\t"
      "subc.u32    %7, %15, %23;\n\t"
      "}"
      : "=r"(d.v[0]), "=r"(d.v[1]), "=r"(d.v[2]), "=r"(d.v[3]), "=r"(d.v[4]), "=r"(d.v[5]), "=r"(d.v[6]), "=r"(d.v[7])
      : "r"(hi.v[0]), "r"(hi.v[1]), "r"(hi.v[2]), "r"(hi.v[3]), "r"(hi.v[4]), "r"(hi.v[5]), "r"(hi.v[6]), "r"(hi.v[7]),
        "r"(lo.v[0]), "r"(lo.v[1]), "r"(lo.v[2]), "r"(lo.v[3]), "r"(lo.v[4]), "r"(lo.v[5]), "r"(lo.v[6]), "r"(lo.v[7]));
  asm("{\n\t"
      "add.cc.u32  %0, %0, %8;\n\t"
      "addc.cc.u32 %1, %1, %9;\n\t"
      "addc.cc.u32 %2, %2, %10;\n\t"
      "addc.cc.u32 %3, %3, %11;\n\t"
      "addc.cc.u32 %4, %4, 0;\n\t"
      "addc.cc.u32 %5, %5, 0;\n\t"
      "addc.cc.u32 %6, %6, 0;\n\t"
      "addc.u32    %7, %7, %12;\n\t"
      "}"
      : "+r"(d.v[0]), "+r"(d.v[1]), "+r"(d.v[2]), "+r"(d.v[3]), "+r"(d.v[4]), "+r"(d.v[5]), "+r"(d.v[6]), "+r"(d.v[7])
      : "r"(SPG_2Q0), "r"(SPG_2Q1), "r"(SPG_2Q2), "r"(SPG_2Q3), "r"(SPG_2Q7));
  return fq_cond_sub(fq_raw_add(lo, fq_mul_lazy(r, d)), SPG_2Q0, SPG_2Q1, SPG_2Q2, SPG_2Q3, SPG_2Q7);
}

template <int ILP, int V>
__global__ void k_bind(fq *out, fq seed, int iters) {
  fq lo[ILP], hi[ILP], r = seed;
#pragma unroll
  for (int i = 0; i < ILP; i++) {
    lo[i] = hi[i] = seed;
    lo[i].v[0] ^= (threadIdx.x + 131 * i);
    hi[i].v[1] ^= (threadIdx.x * 7 + i);
    lo[i].v[7] &= 0x0fffffffu;
    hi[i].v[7] &= 0x0fffffffu;
  }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) {
      fq v = bind_form<V>(lo[i], hi[i], r);
      hi[i] = lo[i];
      lo[i] = v;
    }
  }
  fq s = lo[0];
#pragma unroll
  for (int i = 1; i < ILP; i++)
#pragma unroll
    for (int k = 0; k < 8; k++) s.v[k] ^= lo[i].v[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float time_ms(F f) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  f();
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int i = 0; i < 5; i++) f();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms / 5;
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  int sms = prop.multiProcessorCount;
  void *buf;
  cudaMalloc(&buf, (size_t)sms * 8 * 1024 * 32);
  const int iters = 2048;
  fq seed;
  for (int i = 0; i < 8; i++) seed.v[i] = 0x9e3779b9u * (i + 1);
  seed.v[7] &= 0x0fffffffu;
  printf("{\"sms\": %d", sms);
  {
    int blocks = sms * 8, threads = 256;
    float ms = time_ms([&] { k_imad_wide<8><<<blocks, threads>>>((unsigned long long *)buf, 3, 5, iters); });
    double ops = (double)blocks * threads * iters * 8;
    printf(", \"mad_wide_acc64_split_per_s\": %.4g, \"mad_wide_acc64_split_per_clk_per_sm_at_1965\": %.2f", ops / (ms * 1e-3),
           ops / (ms * 1e-3) / sms / 1.965e9);
  }
  {
    int blocks = sms * 8, threads = 256;
    float ms = time_ms([&] { k_imad_carry<<<blocks, threads>>>((unsigned int *)buf, 3, 5, iters); });
    double ops = (double)blocks * threads * iters * 4 * 4;  // wide mads (pairs)
    printf(", \"imad_wide_carry_per_s\": %.4g, \"imad_wide_carry_per_clk_per_sm_at_1965\": %.2f", ops / (ms * 1e-3),
           ops / (ms * 1e-3) / sms / 1.965e9);
  }
  for (int threads : {128, 256, 512}) {
    int blocks = sms * (2048 / threads);
    float ms = time_ms([&] { k_modmul<2, false><<<blocks, threads>>>((fq *)buf, seed, iters / 8); });
    double ops = (double)blocks * threads * (iters / 8) * 2;
    printf(", \"modmul_lazy_ilp2_t%d_per_s\": %.4g", threads, ops / (ms * 1e-3));
  }
  {
    int threads = 256, blocks = sms * 8;
    float ms = time_ms([&] { k_modmul<1, false><<<blocks, threads>>>((fq *)buf, seed, iters / 8); });
    double ops = (double)blocks * threads * (iters / 8);
    printf(", \"modmul_lazy_ilp1_per_s\": %.4g", ops / (ms * 1e-3));
    ms = time_ms([&] { k_modmul<2, true><<<blocks, threads>>>((fq *)buf, seed, iters / 8); });
    ops = (double)blocks * threads * (iters / 8) * 2;
    printf(", \"modmul_canon_ilp2_per_s\": %.4g", ops / (ms * 1e-3));
    // low occupancy: 2 warps per SMSP, like a 150-register kernel
    blocks = sms;
    ms = time_ms([&] { k_modmul<2, false><<<blocks, 256>>>((fq *)buf, seed, iters / 8); });
    ops = (double)blocks * 256 * (iters / 8) * 2;
    printf(", \"modmul_lazy_ilp2_8warps_per_sm_per_s\": %.4g", ops / (ms * 1e-3));
    ms = time_ms([&] { k_modmul<4, false><<<blocks, 256>>>((fq *)buf, seed, iters / 8); });
    ops = (double)blocks * 256 * (iters / 8) * 4;
    printf(", \"modmul_lazy_ilp4_8warps_per_sm_per_s\": %.4g", ops / (ms * 1e-3));
  }
  {
    int blocks = sms * 8, threads = 256;
    float ms = time_ms([&] { k_add_carry<<<blocks, threads>>>((unsigned int *)buf, 3, iters); });
    double ops = (double)blocks * threads * iters * 4 * 8;
    printf(", \"iadd3_carry_per_s\": %.4g, \"iadd3_carry_per_clk_per_sm_at_1965\": %.2f", ops / (ms * 1e-3), ops / (ms * 1e-3) / sms / 1.965e9);
    // 4 warps per SMSP, the occupancy of the 128-register row kernels
    blocks = sms * 4;
    threads = 128;
    ms = time_ms([&] { k_bind<2, 0><<<blocks, threads>>>((fq *)buf, seed, iters / 8); });
    ops = (double)blocks * threads * (iters / 8) * 2;
    printf(", \"bind_current_per_s\": %.4g", ops / (ms * 1e-3));
    ms = time_ms([&] { k_bind<2, 1><<<blocks, threads>>>((fq *)buf, seed, iters / 8); });
    printf(", \"bind_lean_per_s\": %.4g", ops / (ms * 1e-3));
    ms = time_ms([&] { k_modmul<2, false><<<blocks, threads>>>((fq *)buf, seed, iters / 8); });
    printf(", \"modmul_lazy_ilp2_16warps_per_sm_per_s\": %.4g", ops / (ms * 1e-3));
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf(", \"cuda_status\": \"%s\"}\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
