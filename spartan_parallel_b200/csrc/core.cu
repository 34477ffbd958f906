// Context, device vectors, error reporting and the shared reduction kernels.
#include <atomic>
#include <chrono>
#include <mutex>
#include <unordered_set>

#include "common.cuh"

namespace spg {

static thread_local char g_err[512] = "";

// live contexts (see DeviceGuard): never destroyed, so handle finalisers that run during process exit
// can still consult it
static std::mutex &g_live_mu = *new std::mutex();
static std::unordered_set<const spg_ctx *> &g_live = *new std::unordered_set<const spg_ctx *>();
bool ctx_alive(const spg_ctx *ctx) {
  std::lock_guard<std::mutex> lk(g_live_mu);
  return g_live.count(ctx) != 0;
}

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
  set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
  return e == cudaErrorMemoryAllocation ? SPG_ENOMEM : SPG_ECUDA;
}

cudaError_t dev_alloc_bytes(spg_ctx *ctx, void **p, size_t bytes) {
  if (bytes < SPG_BIG_BYTES) return cudaMallocAsync(p, bytes ? bytes : 32, ctx->stream);
  // best fit among the context's free blocks, wasting at most half of a block
  spg_ctx::BigBlock *best = nullptr;
  for (auto &b : ctx->big)
    if (!b.busy && b.bytes >= bytes && b.bytes / 2 <= bytes && (!best || b.bytes < best->bytes)) best = &b;
  if (best) {
    best->busy = true;
    *p = best->p;
    return cudaSuccess;
  }
  cudaError_t e = cudaMallocAsync(p, bytes, ctx->stream);
  if (e == cudaErrorMemoryAllocation) {
    // give the idle blocks back and try once more
    cudaGetLastError();
    for (size_t i = ctx->big.size(); i-- > 0;)
      if (!ctx->big[i].busy) {
        cudaFreeAsync(ctx->big[i].p, ctx->stream);
        ctx->big.erase(ctx->big.begin() + i);
      }
    e = cudaMallocAsync(p, bytes, ctx->stream);
  }
  if (e == cudaSuccess) ctx->big.push_back(spg_ctx::BigBlock{*p, bytes, true});
  return e;
}

void dev_free(spg_ctx *ctx, void *p) {
  if (!p) return;
  if (!ctx_alive(ctx)) return;  // the context (and with it this memory) is already gone
  for (auto &b : ctx->big)
    if (b.p == p) {
      b.busy = false;
      return;
    }
  cudaFreeAsync(p, ctx->stream);
}

VecOut::~VecOut() {
  if (v) spg_vec_free(v);
}

void prof_begin(spg_ctx *ctx, const char *name, cudaEvent_t *a, cudaEvent_t *b) {
  auto get = [&]() {
    cudaEvent_t e = nullptr;
    if (!ctx->ev_pool.empty()) {
      e = ctx->ev_pool.back();
      ctx->ev_pool.pop_back();
    } else {
      cudaEventCreate(&e);
    }
    return e;
  };
  *a = get();
  *b = get();
  cudaEventRecord(*a, ctx->stream);
  ctx->prof.push_back(spg_ctx::ProfRec{name, *a, *b, ctx->next_units});
}

int ensure_partials(spg_ctx *ctx, size_t n_fq) {
  if (ctx->partial_cap >= n_fq) return SPG_OK;
  if (ctx->d_partials) SPG_CUDA(cudaFree(ctx->d_partials));
  ctx->d_partials = nullptr;
  ctx->partial_cap = 0;
  SPG_CUDA(cudaMalloc(&ctx->d_partials, n_fq * sizeof(fq)));
  ctx->partial_cap = n_fq;
  return SPG_OK;
}

int vec_new(spg_ctx *ctx, size_t n, spg_vec **out) {
  spg_vec *v = new (std::nothrow) spg_vec();
  if (!v) return SPG_ENOMEM;
  v->ctx = ctx;
  v->n = v->cap = n;
  cudaError_t e = dev_alloc(ctx, &v->d, n * sizeof(fq));
  if (e != cudaSuccess) {
    delete v;
    return cuda_fail(e, "cudaMallocAsync(vec)", __FILE__, __LINE__);
  }
  *out = v;
  return SPG_OK;
}

// one block: sums partials[b*width + k] over b for each k < width (width <= 8)
__global__ void k_reduce_partials(const fq *__restrict__ partials, size_t nblocks, int width,
                                  fq *__restrict__ out) {
  __shared__ fq sm[32];
  for (int k = 0; k < width; k++) {
    fq acc = fq_zero();
    for (size_t b = threadIdx.x; b < nblocks; b += blockDim.x)
      acc = fq_add(acc, partials[b * width + k]);
    acc = fq_warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
      fq v = threadIdx.x < ((blockDim.x + 31) >> 5) ? sm[threadIdx.x] : fq_zero();
      v = fq_warp_sum(v);
      if (threadIdx.x == 0) out[k] = v;
    }
    __syncthreads();
  }
}

// stage 1 of a large reduction: block b sums rows b, b + gridDim.x, ... into out[b*width + k]
__global__ void k_reduce_stage(const fq *__restrict__ partials, size_t nblocks, int width, fq *__restrict__ out) {
  __shared__ fq sm[32];
  for (int k = 0; k < width; k++) {
    fq acc = fq_zero();
    for (size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; b < nblocks; b += (size_t)gridDim.x * blockDim.x)
      acc = fq_add(acc, fq_load(partials + b * width + k));
    acc = fq_warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
      fq v = threadIdx.x < ((blockDim.x + 31) >> 5) ? sm[threadIdx.x] : fq_zero();
      v = fq_warp_sum(v);
      if (threadIdx.x == 0) out[(size_t)blockIdx.x * width + k] = v;
    }
    __syncthreads();
  }
}

int reduce_partials(spg_ctx *ctx, const fq *partials, size_t nblocks, int width, fq *d_out) {
  if (nblocks > 4096) {
    // two stages: 64 blocks fold the partial sums into the reserved tail of d_scalars' sibling buffer
    const int stage_blocks = 64;
    if (!ctx->d_stage) SPG_CUDA(cudaMalloc(&ctx->d_stage, (size_t)stage_blocks * 8 * sizeof(fq)));
    SPG_LAUNCH(ctx, k_reduce_stage, stage_blocks, 256, 0, partials, nblocks, width, ctx->d_stage);
    partials = ctx->d_stage;
    nblocks = stage_blocks;
  }
  int threads = nblocks >= 256 ? 256 : (nblocks > 32 ? 128 : 32);
  SPG_LAUNCH(ctx, k_reduce_partials, 1, threads, 0, partials, nblocks, width, d_out);
  return SPG_OK;
}

FinishArgs finish_args(spg_ctx *ctx, size_t nblocks) {
  FinishArgs fa;
  fa.partials = ctx->d_partials;
  fa.counter = ctx->d_counter;
  fa.result = ctx->d_result;
  fa.flag = ctx->d_flag;
  // one block adding up to 4096 partial rows costs a few microseconds; beyond that the
  // two-stage reduce kernels are faster and the extra launch is noise next to the main kernel
  fa.seq = nblocks <= 4096 ? ++ctx->seq : 0;
  return fa;
}

int wait_flag(spg_ctx *ctx, unsigned long long seq, int width, spg_fq *out) {
  volatile unsigned long long *f = ctx->h_flag;
  unsigned int spins = 0;
  auto t0 = std::chrono::steady_clock::now();
  while (*f != seq) {
    spin_pause();
    if ((++spins & 0x3fff) == 0) {
      cudaError_t e = cudaStreamQuery(ctx->stream);
      if (e != cudaSuccess && e != cudaErrorNotReady) return cuda_fail(e, "round kernel", __FILE__, __LINE__);
      if (e == cudaSuccess && *f != seq) {
        set_error("round kernel finished without publishing its result (flag %llu, expected %llu)", (unsigned long long)*f,
                  seq);
        return SPG_ECUDA;
      }
      if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 30.0) {
        set_error("timed out waiting for a round kernel");
        return SPG_ECUDA;
      }
    }
  }
  std::atomic_thread_fence(std::memory_order_acquire);
  memcpy(out, ctx->h_result, sizeof(spg_fq) * width);
  return SPG_OK;
}

int finish_result(spg_ctx *ctx, const FinishArgs &fa, size_t nblocks, int width, spg_fq *out) {
  if (fa.seq) return wait_flag(ctx, fa.seq, width, out);
  SPG_TRY(reduce_partials(ctx, ctx->d_partials, nblocks, width, ctx->d_result));
  return fetch_result(ctx, width, out);
}

int fetch_result(spg_ctx *ctx, int width, spg_fq *out) {
  SPG_CUDA(cudaStreamSynchronize(ctx->stream));
  memcpy(out, ctx->h_result, sizeof(spg_fq) * width);
  return SPG_OK;
}

// First scalar of up to HEADS_PER_LAUNCH device tables -> the mapped result page, published through the
// flag word like a round result: the final claims of a sumcheck (one scalar per bound table) reach the
// host with one tiny launch instead of one cudaMemcpyAsync per table (25 - 43 copies of 32 bytes at
// ~6 us each per product-circuit layer, 42 layers per sparse proof).
constexpr int HEADS_PER_LAUNCH = 48;  // the page holds 56 scalars before the flag word
struct HeadPtrs {
  const fq *p[HEADS_PER_LAUNCH];
};
__global__ void k_gather_heads(const __grid_constant__ HeadPtrs P, int n, fq *__restrict__ result, unsigned long long *flag,
                               unsigned long long seq) {
  if ((int)threadIdx.x < n) {
    fq_store(result + threadIdx.x, fq_load(P.p[threadIdx.x]));
    __threadfence_system();
  }
  __syncthreads();
  if (threadIdx.x == 0) *(volatile unsigned long long *)flag = seq;
}

int gather_heads(spg_ctx *ctx, const fq *const *ptrs, size_t n, spg_fq *out) {
  for (size_t base = 0; base < n; base += HEADS_PER_LAUNCH) {
    int cnt = (int)(n - base < (size_t)HEADS_PER_LAUNCH ? n - base : (size_t)HEADS_PER_LAUNCH);
    HeadPtrs P;
    memset(&P, 0, sizeof P);
    for (int i = 0; i < cnt; i++) P.p[i] = ptrs[base + i];
    unsigned long long seq = ++ctx->seq;
    SPG_LAUNCH(ctx, k_gather_heads, 1, 64, 0, P, cnt, ctx->d_result, ctx->d_flag, seq);
    SPG_TRY(wait_flag(ctx, seq, cnt, out + base));
  }
  return SPG_OK;
}

}  // namespace spg

using namespace spg;

extern "C" {

const char *spg_last_error(void) { return g_err; }
int spg_version(void) { return 100; }

int spg_ctx_create(int device, spg_ctx **out) {
  SPG_CHECK(out != nullptr, "spg_ctx_create: out is null");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    set_error("spg_ctx_create: no CUDA device available (%s); this backend has no CPU fallback",
              cudaGetErrorString(e));
    return SPG_ECUDA;
  }
  SPG_CHECK(device >= 0 && device < count, "spg_ctx_create: device %d out of range (%d devices)",
            device, count);
  SPG_CUDA(cudaSetDevice(device));
  spg_ctx *ctx = new (std::nothrow) spg_ctx();
  if (!ctx) return SPG_ENOMEM;
  ctx->device = device;
  cudaDeviceProp prop;
  SPG_CUDA(cudaGetDeviceProperties(&prop, device));
  ctx->sm_count = prop.multiProcessorCount;
  SPG_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  SPG_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  {
    cudaMemPool_t pool;
    SPG_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
    uint64_t keep = UINT64_MAX;
    SPG_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
  }
  SPG_CUDA(cudaHostAlloc(&ctx->h_result, 64 * sizeof(fq), cudaHostAllocMapped));
  SPG_CUDA(cudaHostGetDevicePointer(&ctx->d_result, ctx->h_result, 0));
  SPG_CUDA(cudaMalloc(&ctx->d_scalars, 64 * sizeof(fq)));
  SPG_CUDA(cudaMalloc(&ctx->d_counter, sizeof(unsigned int)));
  SPG_CUDA(cudaMemset(ctx->d_counter, 0, sizeof(unsigned int)));
  ctx->h_flag = (unsigned long long *)(ctx->h_result + 56);  // tail of the mapped result page
  ctx->d_flag = (unsigned long long *)(ctx->d_result + 56);
  *ctx->h_flag = 0;
  int rc = ensure_partials(ctx, (size_t)ctx->sm_count * 16 * 8);
  if (rc != SPG_OK) return rc;
  {
    std::lock_guard<std::mutex> lk(g_live_mu);
    g_live.insert(ctx);
  }
  *out = ctx;
  return SPG_OK;
}

void spg_ctx_destroy(spg_ctx *ctx) {
  if (!ctx) return;
  {
    std::lock_guard<std::mutex> lk(g_live_mu);
    if (!g_live.erase(ctx)) return;  // not a live context (destroyed twice)
  }
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (auto &b : ctx->big) cudaFreeAsync(b.p, ctx->stream);
  ctx->big.clear();
  cudaStreamSynchronize(ctx->stream);
  if (ctx->d_partials) cudaFree(ctx->d_partials);
  if (ctx->d_scalars) cudaFree(ctx->d_scalars);
  if (ctx->d_counter) cudaFree(ctx->d_counter);
  if (ctx->d_stage) cudaFree(ctx->d_stage);
  if (ctx->h_result) cudaFreeHost(ctx->h_result);
  if (ctx->copy_stream) {
    cudaStreamSynchronize(ctx->copy_stream);
    cudaStreamDestroy(ctx->copy_stream);
  }
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

int spg_ctx_sync(spg_ctx *ctx) {
  spg::DeviceGuard _dev(spg::ctx_of(ctx));
  SPG_CHECK(ctx, "null ctx");
  SPG_CUDA(cudaStreamSynchronize(ctx->stream));
  return SPG_OK;
}

uint64_t spg_ctx_launch_count(const spg_ctx *ctx) { return ctx ? ctx->launches : 0; }

int spg_ctx_profile_begin(spg_ctx *ctx) {
  spg::DeviceGuard _dev(spg::ctx_of(ctx));
  SPG_CHECK(ctx, "null ctx");
  SPG_CUDA(cudaStreamSynchronize(ctx->stream));
  for (auto &r : ctx->prof) {
    ctx->ev_pool.push_back(r.a);
    ctx->ev_pool.push_back(r.b);
  }
  ctx->prof.clear();
  ctx->profiling = true;
  return SPG_OK;
}

int spg_ctx_profile_end(spg_ctx *ctx, char *out, size_t cap) {
  spg::DeviceGuard _dev(spg::ctx_of(ctx));
  SPG_CHECK(ctx && out && cap > 2, "spg_ctx_profile_end: bad buffer");
  SPG_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->profiling = false;
  struct Agg {
    const char *name;
    double ms, units, max_ms, max_units;
    size_t n;
  };
  std::vector<Agg> aggs;
  for (auto &r : ctx->prof) {
    float ms = 0;
    cudaEventElapsedTime(&ms, r.a, r.b);
    Agg *g = nullptr;
    for (auto &a : aggs)
      if (a.name == r.name || strcmp(a.name, r.name) == 0) g = &a;
    if (!g) {
      aggs.push_back(Agg{r.name, 0, 0, 0, 0, 0});
      g = &aggs.back();
    }
    g->ms += ms;
    g->units += r.units;
    g->n++;
    if (ms > g->max_ms) {
      g->max_ms = ms;
      g->max_units = r.units;
    }
    ctx->ev_pool.push_back(r.a);
    ctx->ev_pool.push_back(r.b);
  }
  ctx->prof.clear();
  std::string js = "[";
  char buf[512];
  for (size_t i = 0; i < aggs.size(); i++) {
    // template instantiations are passed to SPG_LAUNCH in parentheses: drop them from the name
    std::string nm = aggs[i].name;
    if (nm.size() >= 2 && nm.front() == '(' && nm.back() == ')') nm = nm.substr(1, nm.size() - 2);
    snprintf(buf, sizeof buf,
             "%s{\"kernel\": \"%s\", \"launches\": %zu, \"total_ms\": %.6f, \"units\": %.6g, "
             "\"max_ms\": %.6f, \"max_units\": %.6g}",
             i ? ", " : "", nm.c_str(), aggs[i].n, aggs[i].ms, aggs[i].units, aggs[i].max_ms, aggs[i].max_units);
    js += buf;
  }
  js += "]";
  SPG_CHECK(js.size() + 1 <= cap, "spg_ctx_profile_end: buffer too small (%zu needed)", js.size() + 1);
  memcpy(out, js.c_str(), js.size() + 1);
  return SPG_OK;
}
void *spg_ctx_stream(const spg_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int spg_fq_host_sum(const spg_fq *in, size_t count, size_t width, spg_fq *out) {
  SPG_CHECK(in && out, "spg_fq_host_sum: null argument");
  for (size_t k = 0; k < width; k++) {
    hfq acc = hfq_zero();
    for (size_t i = 0; i < count; i++) acc = hfq_add(acc, hfq_from(in[i * width + k]));
    out[k] = hfq_to(acc);
  }
  return SPG_OK;
}

int spg_fq_host_mul(const spg_fq *a, const spg_fq *b, spg_fq *out) {
  SPG_CHECK(a && b && out, "spg_fq_host_mul: null argument");
  *out = hfq_to(hfq_mul(hfq_from(*a), hfq_from(*b)));
  return SPG_OK;
}

int spg_fq_host_eq_weight(const spg_fq *tau, size_t nbits, uint64_t index, spg_fq *out) {
  SPG_CHECK((tau || nbits == 0) && out, "spg_fq_host_eq_weight: null argument");
  hfq acc = hfq_one(), one = hfq_one();
  for (size_t k = 0; k < nbits; k++) {
    hfq t = hfq_from(tau[k]);
    acc = hfq_mul(acc, ((index >> k) & 1) ? t : hfq_sub(one, t));
  }
  *out = hfq_to(acc);
  return SPG_OK;
}

int spg_host_alloc(size_t bytes, void **out) {
  SPG_CHECK(out, "null out");
  SPG_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
  return SPG_OK;
}
void spg_host_free(void *p) {
  if (p) cudaFreeHost(p);
}

int spg_vec_alloc(spg_ctx *ctx, size_t n, spg_vec **out) {
  spg::DeviceGuard _dev(spg::ctx_of(ctx));
  SPG_CHECK(ctx && out, "spg_vec_alloc: null argument");
  return vec_new(ctx, n, out);
}

int spg_vec_zero(spg_ctx *ctx, spg_vec *v) {
  spg::DeviceGuard _dev(spg::ctx_of(ctx));
  SPG_CHECK(ctx && v, "spg_vec_zero: null argument");
  if (v->n) SPG_CUDA(cudaMemsetAsync(v->d, 0, v->n * sizeof(fq), ctx->stream));
  return SPG_OK;
}

int spg_vec_upload(spg_ctx *ctx, const spg_fq *host, size_t n, spg_vec **out) {
  spg::DeviceGuard _dev(spg::ctx_of(ctx));
  SPG_CHECK(ctx && out && (host || n == 0), "spg_vec_upload: null argument");
  spg_vec *v = nullptr;
  SPG_TRY(vec_new(ctx, n, &v));
  if (n) {
    cudaError_t e = cudaMemcpyAsync(v->d, host, n * sizeof(fq), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
      spg_vec_free(v);
      return cuda_fail(e, "upload", __FILE__, __LINE__);
    }
  }
  *out = v;
  return SPG_OK;
}

int spg_vec_wrap(spg_ctx *ctx, void *device_ptr, size_t n, spg_vec **out) {
  spg::DeviceGuard _dev(spg::ctx_of(ctx));
  SPG_CHECK(ctx && out && device_ptr, "spg_vec_wrap: null argument");
  SPG_CHECK(((uintptr_t)device_ptr & 31) == 0, "spg_vec_wrap: pointer must be 32-byte aligned");
  spg_vec *v = new (std::nothrow) spg_vec();
  if (!v) return SPG_ENOMEM;
  v->ctx = ctx;
  v->d = (fq *)device_ptr;
  v->n = v->cap = n;
  v->owned = false;
  *out = v;
  return SPG_OK;
}

int spg_vec_download(spg_ctx *ctx, const spg_vec *v, size_t offset, size_t n, spg_fq *host) {
  spg::DeviceGuard _dev(spg::ctx_of(ctx));
  SPG_CHECK(ctx && v && (host || n == 0), "spg_vec_download: null argument");
  SPG_CHECK(offset + n <= v->n, "spg_vec_download: range [%zu, %zu) exceeds length %zu", offset,
            offset + n, v->n);
  if (n) {
    SPG_CUDA(cudaMemcpyAsync(host, v->d + offset, n * sizeof(fq), cudaMemcpyDeviceToHost, ctx->stream));
    SPG_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return SPG_OK;
}

size_t spg_vec_len(const spg_vec *v) { return v ? v->n : 0; }
void *spg_vec_device_ptr(const spg_vec *v) { return v ? (void *)v->d : nullptr; }

void spg_vec_free(spg_vec *v) {
  spg::DeviceGuard _dev(spg::ctx_of(v));
  if (!v) return;
  if (v->owned && v->d) dev_free(v->ctx, v->d);
  delete v;
}

}  // extern "C"
