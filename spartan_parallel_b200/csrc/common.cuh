// Shared host-side plumbing for libspgpu: context, device vectors, error handling,
// launch accounting and the small reductions every sumcheck round ends with.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/spgpu.h"
#include "fq.cuh"
#include "host_fq.h"

namespace spg {

// ---------------------------------------------------------------- errors
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define SPG_CUDA(expr)                                                       \
  do {                                                                       \
    cudaError_t _e = (expr);                                                 \
    if (_e != cudaSuccess) return spg::cuda_fail(_e, #expr, __FILE__, __LINE__); \
  } while (0)

#define SPG_CHECK(cond, ...)      \
  do {                            \
    if (!(cond)) {                \
      spg::set_error(__VA_ARGS__); \
      return SPG_EINVAL;          \
    }                             \
  } while (0)

#define SPG_TRY(expr)          \
  do {                         \
    int _rc = (expr);          \
    if (_rc != SPG_OK) return _rc; \
  } while (0)

static inline bool is_pow2(size_t n) { return n && !(n & (n - 1)); }
static inline size_t next_pow2(size_t n) {
  size_t p = 1;
  while (p < n) p <<= 1;
  return p;
}
static inline unsigned log2u(size_t n) {
  unsigned l = 0;
  while (((size_t)1 << l) < n) l++;
  return l;
}

}  // namespace spg

// ---------------------------------------------------------------- handles
struct spg_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;  // H2D uploads that overlap with kernels on `stream`
  uint64_t launches = 0;
  // per-block partial sums of the round kernels, and the 3-scalar result slot
  spg::fq *d_partials = nullptr;
  size_t partial_cap = 0;  // in fq
  spg::fq *h_result = nullptr;  // pinned + mapped
  spg::fq *d_result = nullptr;  // device alias of h_result
  // in-kernel final reduction of small grids: ticket counter (zero between launches) and the
  // sequence flag the last block publishes in mapped host memory after the result
  unsigned int *d_counter = nullptr;
  unsigned long long *h_flag = nullptr, *d_flag = nullptr;
  unsigned long long seq = 0;
  // small staging for scalar arguments
  spg::fq *d_scalars = nullptr;  // device scratch, 64 fq
  spg::fq *d_stage = nullptr;    // first-stage output of large reductions
  // optional per-kernel timing with CUDA events on the launching stream
  bool profiling = false;
  struct ProfRec {
    const char *name;
    cudaEvent_t a, b;
    double units;  // caller-defined work units of the launch (bytes, items ...)
  };
  std::vector<ProfRec> prof;
  std::vector<cudaEvent_t> ev_pool;
  double next_units = 0;  // set by the launcher just before SPG_LAUNCH
  // large per-proof buffers (tables, witness sections) recycled by the context itself, see dev_alloc
  struct BigBlock {
    void *p;
    size_t bytes;
    bool busy;
  };
  std::vector<BigBlock> big;
};

struct spg_vec {
  spg_ctx *ctx = nullptr;
  spg::fq *d = nullptr;
  size_t n = 0;      // logical length
  size_t cap = 0;    // allocated length
  bool owned = true;
};

namespace spg {

// Every entry point that takes a context or a handle runs on that context's device, whatever
// device the calling thread had current (another host thread than the creating one -- current
// device is per thread and defaults to 0 --, a library that switched devices in between, one
// process holding several contexts), and leaves the caller's current device as it found it.
// Contexts that exist right now. Handles may outlive their context (a host language's finalisers run
// in no particular order at exit): such a handle's destroy function must not touch the dead context --
// its device memory went with spg_ctx_destroy -- and only frees its own host-side struct.
bool ctx_alive(const spg_ctx *ctx);
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(const spg_ctx *ctx) {
    if (!ctx || !ctx_alive(ctx)) return;
    int cur = -1;
    if (cudaGetDevice(&cur) == cudaSuccess && cur != ctx->device && cudaSetDevice(ctx->device) == cudaSuccess) prev = cur;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard &) = delete;
  DeviceGuard &operator=(const DeviceGuard &) = delete;
};
template <typename H>
static inline const spg_ctx *ctx_of(const H *h) {
  return h ? h->ctx : nullptr;
}
static inline const spg_ctx *ctx_of(const spg_ctx *c) { return c; }

// launch bookkeeping: every kernel goes through this so gpu_launches is exact
#define SPG_LAUNCH(ctx, kernel, grid, block, smem, ...)                       \
  do {                                                                        \
    cudaEvent_t _ea = nullptr, _eb = nullptr;                                 \
    if ((ctx)->profiling) spg::prof_begin((ctx), #kernel, &_ea, &_eb);        \
    kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);          \
    if (_eb) cudaEventRecord(_eb, (ctx)->stream);                             \
    (ctx)->launches++;                                                        \
    (ctx)->next_units = 0;                                                    \
    SPG_CUDA(cudaGetLastError());                                             \
  } while (0)

void prof_begin(spg_ctx *ctx, const char *name, cudaEvent_t *a, cudaEvent_t *b);

// Stream-ordered allocation. Small buffers come from the device's default pool (kept warm: the
// release threshold is raised in spg_ctx_create). Buffers of SPG_BIG_BYTES and more -- the
// per-proof tables and witness sections, GiBs each -- are recycled by the context itself: a
// freed block stays with the context and serves the next request it fits (same stream, so the
// stream order that made cudaFreeAsync -> cudaMallocAsync reuse legal still holds). The pool
// alone is not enough: one small live allocation carved out of a freed 2 GiB block (the tail
// prover of a sharded proof, say) forces the pool to map fresh physical memory for the next
// 2 GiB request, and with eight processes doing that at once the mapping calls serialise in
// the driver -- 12-15 ms per proof at 8 GPUs, as measured.
constexpr size_t SPG_BIG_BYTES = (size_t)16 << 20;
cudaError_t dev_alloc_bytes(spg_ctx *ctx, void **p, size_t bytes);
void dev_free(spg_ctx *ctx, void *p);
template <typename T>
static inline cudaError_t dev_alloc(spg_ctx *ctx, T **p, size_t bytes) {
  return dev_alloc_bytes(ctx, (void **)p, bytes);
}
// Scope guards for the temporaries of an entry point: every early return (SPG_CUDA / SPG_TRY /
// SPG_CHECK) releases them, so a failed call does not pin pool blocks or leak output vectors.
struct DevTmp {
  spg_ctx *ctx;
  void *p = nullptr;
  explicit DevTmp(spg_ctx *c) : ctx(c) {}
  ~DevTmp() {
    if (p) dev_free(ctx, p);
  }
  cudaError_t alloc(size_t bytes) { return dev_alloc_bytes(ctx, &p, bytes); }
  template <typename T>
  T *as() const {
    return static_cast<T *>(p);
  }
  DevTmp(const DevTmp &) = delete;
  DevTmp &operator=(const DevTmp &) = delete;
};
struct CudaTmp {  // plain cudaMalloc'ed scratch
  void *p = nullptr;
  ~CudaTmp() {
    if (p) cudaFree(p);
  }
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes); }
  template <typename T>
  T *as() const {
    return static_cast<T *>(p);
  }
  CudaTmp() = default;
  CudaTmp(const CudaTmp &) = delete;
  CudaTmp &operator=(const CudaTmp &) = delete;
};
struct VecOut {  // an output vector that is freed unless released
  spg_vec *v = nullptr;
  ~VecOut();
  spg_vec *release() {
    spg_vec *r = v;
    v = nullptr;
    return r;
  }
  VecOut() = default;
  VecOut(const VecOut &) = delete;
  VecOut &operator=(const VecOut &) = delete;
};

// one exchange through the host mailbox of a sharded proof (csrc/sc1.cu): publish `mine`, collect all ranks'
int mailbox_exchange(char *base, size_t slot_stride, int rank, int world, uint64_t c, const void *mine, size_t nbytes,
                     void *out);

int ensure_partials(spg_ctx *ctx, size_t n_fq);
int vec_new(spg_ctx *ctx, size_t n, spg_vec **out);

// grid sizing: persistent-style grids in multiples of the SM count
static inline int grid_for(const spg_ctx *ctx, size_t items, int block, int max_waves = 8) {
  size_t blocks = (items + block - 1) / block;
  size_t cap = (size_t)ctx->sm_count * max_waves;
  if (blocks > cap) blocks = cap;
  if (blocks == 0) blocks = 1;
  return (int)blocks;
}

// sum `count` groups of `width` scalars laid out [block][width] into out[width]
int reduce_partials(spg_ctx *ctx, const fq *partials, size_t nblocks, int width, fq *d_out);
// fetch `width` scalars from the mapped result slot after a stream sync
int fetch_result(spg_ctx *ctx, int width, spg_fq *out);
// out[i] = *ptrs[i] for n device scalars, through the mapped result page (one tiny launch per 48)
int gather_heads(spg_ctx *ctx, const fq *const *ptrs, size_t n, spg_fq *out);

// Final reduction inside the round kernel: every block stores its partial sums; the block
// that draws the last ticket adds all of them, writes the result to mapped host memory and then
// publishes `seq` in the flag word, which the host polls (spg::wait_flag) -- one launch and no
// stream synchronisation per round. seq == 0: a separate reduce kernel follows instead.
struct FinishArgs {
  fq *partials;  // [gridDim.x][W]
  unsigned int *counter;
  fq *result;
  unsigned long long *flag;
  unsigned long long seq;
};
// polite busy-wait: the spinning hardware thread yields its pipeline to the sibling hyperthread
static inline void spin_pause() {
#if defined(__x86_64__) || defined(__i386__)
  __builtin_ia32_pause();
#endif
}
FinishArgs finish_args(spg_ctx *ctx, size_t nblocks);  // seq != 0 iff the grid is small enough
int wait_flag(spg_ctx *ctx, unsigned long long seq, int width, spg_fq *out);
// the launch's result: polled when the kernel finished the reduction itself, else reduce + sync
int finish_result(spg_ctx *ctx, const FinishArgs &fa, size_t nblocks, int width, spg_fq *out);

// block-level modular sum of `W` accumulators; thread 0 of the block gets the result
template <int W>
__device__ __forceinline__ void block_sum(fq (&acc)[W], fq *smem /* [W * 32] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < W; k++) acc[k] = fq_warp_sum(acc[k]);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < W; k++) smem[k * 32 + warp] = acc[k];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < W; k++) {
      fq v = lane < nwarps ? smem[k * 32 + lane] : fq_zero();
      acc[k] = fq_warp_sum(v);
    }
  }
}

// `mine` (valid in thread 0) = this block's W sums; see FinishArgs
template <int W>
__device__ __forceinline__ void finish_block(const FinishArgs &fa, const fq (&mine)[W], fq *smem /* [W * 32] */) {
  if (gridDim.x == 1 && fa.seq) {
    // a single block: its sums are the result. The late rounds of every sumcheck run like this, one
    // warp per scheduler with nothing to hide latency behind, so the partials round trip, the ticket
    // and the second block_sum (~1500 dependent instructions) were ~5 us of a ~25 us round.
    if (threadIdx.x == 0) {
#pragma unroll
      for (int k = 0; k < W; k++) fq_store(fa.result + k, mine[k]);
      __threadfence_system();
      *(volatile unsigned long long *)fa.flag = fa.seq;
    }
    return;
  }
  __shared__ int is_last;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < W; k++) fq_store(fa.partials + (size_t)blockIdx.x * W + k, mine[k]);
    int last = 0;
    if (fa.seq) {
      __threadfence();
      last = atomicAdd(fa.counter, 1u) == gridDim.x - 1;
    }
    is_last = last;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  fq acc[W];
#pragma unroll
  for (int k = 0; k < W; k++) acc[k] = fq_zero();
  for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x)
#pragma unroll
    for (int k = 0; k < W; k++) acc[k] = fq_add(acc[k], fq_load_cg(fa.partials + (size_t)b * W + k));
  __syncthreads();  // smem may still hold the caller's own block_sum
  block_sum<W>(acc, smem);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < W; k++) fq_store(fa.result + k, acc[k]);
    *fa.counter = 0;
    __threadfence_system();
    *(volatile unsigned long long *)fa.flag = fa.seq;
  }
}

}  // namespace spg
