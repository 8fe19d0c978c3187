// bik.cu -- sm_100a kernels and the C ABI of libbik (include/bik.h).
//
// Kernels (all persistent: grid = SMs x resident CTAs, CTAs loop over tiles):
//   k1_kernel<G>    FK + task errors/Jacobians + collision rows; G lanes per instance, problem image
//                   staged into shared memory by one bulk async copy (cp.async.bulk + mbarrier),
//                   Jacobian rows staged per warp and flushed as contiguous runs.
//   k2_kernel<T>    QP assembly + active-set solve, one warp per instance, packed H and Cholesky
//                   factor in shared memory (T = double | float).
//   fk_kernel<G>    frame poses / CoM / body-frame Jacobians for the Configuration API.
//   integrate_kernel, check_limits_kernel: one thread per instance.
//
// There is no CPU fallback in this library: without a CUDA device every entry point fails.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include <mutex>
#include <string>
#include <vector>

#include "bik_build.h"
#include "bik_k2lr.h"
#include "bik_k2t.h"
#include "bik_k2x.h"

using namespace bik;

// ------------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D bulk async copy (TMA engine; SASS: UBLKCP / SYNCS)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t phase) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(phase)
      : "memory");
  return ok != 0;
}

// Stage the problem image into shared memory: one elected thread issues a single bulk copy and
// everybody waits on the mbarrier (bounded spin; traps instead of hanging the GPU).
__device__ __forceinline__ void stage_image(uint32_t* smem_image, const uint32_t* gimage, int words, uint64_t* bar, int use_tma) {
  if (use_tma) {
    if (threadIdx.x == 0) {
      mbar_init(bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      mbar_expect_tx(bar, (uint32_t)words * 4u);
      bulk_g2s(smem_image, gimage, (uint32_t)words * 4u, bar);
    }
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, 0)) {
      if (++spins > (1u << 24)) __trap();
    }
  } else {
    const uint4* src = reinterpret_cast<const uint4*>(gimage);
    uint4* dst = reinterpret_cast<uint4*>(smem_image);
    for (int k = threadIdx.x; k < words / 4; k += blockDim.x) dst[k] = src[k];
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
#ifndef BIK_K1_MINBLOCKS
#define BIK_K1_MINBLOCKS 5   // CTAs of 4 warps per SM the register allocation must allow: 5 -> 96 registers, no spills, 20 warps/SM
                             // (G1: 0.101 -> 0.094 ms; 6 -> 80 registers spills: 0.104 ms)
#endif
template <int G>
__global__ void __launch_bounds__(128, BIK_K1_MINBLOCKS) k1_kernel(const uint32_t* __restrict__ gimage, int words, int use_tma, K1Args a) {
  extern __shared__ __align__(16) uint32_t smem[];
  __shared__ __align__(8) uint64_t bar;
  stage_image(smem, gimage, words, &bar, use_tma);
  PView P{smem};
  constexpr int IPW = 32 / G;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  float* wsm = reinterpret_cast<float*>(smem + words) + warp * k1_warp_words(P.h(), IPW);
  const int ntiles = (a.B + IPW - 1) / IPW;
  for (int tile = blockIdx.x * nwarps + warp; tile < ntiles; tile += gridDim.x * nwarps) k1_warp_tile<G, 32>(P, a, tile * IPW, wsm, lane);
}

template <typename T, int SLOTS>
__global__ void __launch_bounds__(256) k2_kernel(const uint32_t* __restrict__ gimage, int words, int use_tma, K2Args a) {
  extern __shared__ __align__(16) uint32_t smem[];
  __shared__ __align__(8) uint64_t bar;
  stage_image(smem, gimage, words, &bar, use_tma);
  PView P{smem};
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  char* wsm = reinterpret_cast<char*>(smem + words) + (size_t)warp * k2_warp_bytes(P.h(), sizeof(T));
  if (a.lockstep && a.dq) {  // block-uniform trip count: every warp reaches every CTA-wide vote
    for (int base = blockIdx.x * nwarps; base < a.B; base += gridDim.x * nwarps) {
      __syncthreads();
      k2_warp<T, 32, SLOTS>(P, a, base + warp, wsm, lane, base + warp < a.B);
    }
    return;
  }
  for (int b = blockIdx.x * nwarps + warp; b < a.B; b += gridDim.x * nwarps) k2_warp<T, 32, SLOTS>(P, a, b, wsm, lane);
}

// Small-group K2 (bik_k2t.h): G lanes per problem, 32/G problems per warp, tables read straight from the (L1-resident)
// global image so that all of shared memory goes to the per-problem triangles; warps are independent.
// Tiles are handed out by an atomic counter (sched[0]) when the launcher has one for this stream: the pivoting count
// differs between instances (G1: 91 % need one iteration, 1 in 10 000 needs six or more), and with a static round-robin
// the warp that meets a slow instance also keeps its whole share of ordinary tiles, which set the kernel time.  The last
// CTA to leave (sched[1] counts them) rewinds the counter for the next launch on the stream.
template <typename T, int G, int MAXT, typename M>
__global__ void __launch_bounds__(MAXT, 512 / MAXT) k2t_kernel(const uint32_t* __restrict__ gimage, int warp_bytes, K2Args a, unsigned int* sched) {
  extern __shared__ __align__(16) uint32_t smem[];
  constexpr int NS = 32 / G;
  PView P{gimage};
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  char* wsm = reinterpret_cast<char*>(smem) + (size_t)warp * warp_bytes;
  const long long ntiles = ((long long)a.B + NS - 1) / NS;
  const int St = k2t_task_tile_words(P), uw = k2t_union_words(P, sizeof(T));
  if (sched) {
    for (;;) {
      unsigned int t = 0;
      if (lane == 0) t = atomicAdd(&sched[0], 1u);
      t = __shfl_sync(0xffffffffu, t, 0);
      if ((long long)t >= ntiles) break;
      k2t_warp_tile<T, G, NS, M>(P, a, (long long)t * NS, wsm, lane, St, uw);
    }
    __syncthreads();
    if (threadIdx.x == 0 && atomicInc(&sched[1], gridDim.x - 1) == gridDim.x - 1) { __threadfence(); sched[0] = 0u; }
    return;
  }
  for (long long tile = (long long)blockIdx.x * nwarps + warp; tile < ntiles; tile += (long long)gridDim.x * nwarps)
    k2t_warp_tile<T, G, NS, M>(P, a, tile * NS, wsm, lane, St, uw);
}

// Fixed-size thread-per-problem K2 (bik_k2x.h): 32 problems per warp, factor in shared memory, H / c / box in a per-warp
// global scratch (L2 resident), tables read straight from the global image; warps are independent.
template <typename TL, typename TH, int N>
__global__ void __launch_bounds__(256) k2x_kernel(const uint32_t* __restrict__ gimage, int warp_smem, unsigned char* scratch, unsigned long long warp_scratch, K2Args a) {
  extern __shared__ __align__(16) uint32_t smem[];
  PView P{gimage};
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  char* wsm = reinterpret_cast<char*>(smem) + (size_t)warp * warp_smem;
  unsigned char* wsc = scratch + ((size_t)blockIdx.x * nwarps + warp) * warp_scratch;
  const long long ntiles = ((long long)a.B + 31) / 32;
  for (long long tile = (long long)blockIdx.x * nwarps + warp; tile < ntiles; tile += (long long)gridDim.x * nwarps)
    k2x_warp_tile<TL, TH, N, 32>(P, a, tile * 32, wsm, wsc, lane);
}

template <int SLOTS>
__global__ void __launch_bounds__(256) k2lr_kernel(const uint32_t* __restrict__ gimage, int words, int use_tma, K2Args a) {
  extern __shared__ __align__(16) uint32_t smem[];
  __shared__ __align__(8) uint64_t bar;
  stage_image(smem, gimage, words, &bar, use_tma);
  PView P{smem};
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  // packed-triangle index -> (row, col) table of the K x K capacitance matrix, shared by the CTA
  const int K = P.h().K, NP = tri(K);
  uint16_t* pairtab = reinterpret_cast<uint16_t*>(smem + words);
  for (int p = threadIdx.x; p < NP; p += blockDim.x) { int r, s; tri_unflatten(p, &r, &s); pairtab[p] = (uint16_t)((r << 8) | s); }
  __syncthreads();
  char* wsm = reinterpret_cast<char*>(smem + words) + ((NP * 2 + 15) & ~15) + (size_t)warp * k2lr_warp_bytes(P.h());
  for (int b = blockIdx.x * nwarps + warp; b < a.B; b += gridDim.x * nwarps) k2lr_warp<32, SLOTS>(P, a, b, wsm, pairtab, lane);
}

template <int G>
__global__ void __launch_bounds__(128) fk_kernel(const uint32_t* __restrict__ gimage, int words, int use_tma, FkArgs a) {
  extern __shared__ __align__(16) uint32_t smem[];
  __shared__ __align__(8) uint64_t bar;
  stage_image(smem, gimage, words, &bar, use_tma);
  PView P{smem};
  constexpr int IPW = 32 / G;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int wwords = (IPW * ((7 * P.h().nnode) | 1) + 3) & ~3;
  float* wsm = reinterpret_cast<float*>(smem + words) + warp * wwords;
  const int ntiles = (a.B + IPW - 1) / IPW;
  for (int tile = blockIdx.x * nwarps + warp; tile < ntiles; tile += gridDim.x * nwarps) fk_warp_tile<G, 32>(P, a, tile * IPW, wsm, lane);
}

__global__ void integrate_kernel(const uint32_t* __restrict__ gimage, int B, float* q, const float* dq) {
  PView P{gimage};
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) integrate_instance(P, q + (size_t)b * P.h().nq, dq + (size_t)b * P.h().nv);
}
// Converge driver: instances whose every frame-task error is under the thresholds stop (done = 1, iters = steps taken so far);
// the others are counted.  e is K1's unweighted error at the CURRENT q.
__global__ void converge_check_kernel(const uint32_t* __restrict__ gimage, int B, const float* __restrict__ e, int it, float pos_thr, float ori_thr,
                                      int32_t* done, int32_t* iters, int* not_done) {
  PView P{gimage};
  const PHeader& h = P.h();
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B || done[b]) return;
  bool ok = true;
  for (int f = 0; f < h.F; ++f) {
    const float* eb = e + (size_t)b * h.K + P.frame(f).row0;
    const float pos = sqrtf(eb[0] * eb[0] + eb[1] * eb[1] + eb[2] * eb[2]), ori = sqrtf(eb[3] * eb[3] + eb[4] * eb[4] + eb[5] * eb[5]);
    ok = ok && pos <= pos_thr && ori <= ori_thr;
  }
  if (ok) { done[b] = 1; iters[b] = it; }
  else atomicAdd(not_done, 1);
}
__global__ void integrate_masked_kernel(const uint32_t* __restrict__ gimage, int B, float* q, const float* dq, const int32_t* __restrict__ done) {
  PView P{gimage};
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B && !done[b]) integrate_instance(P, q + (size_t)b * P.h().nq, dq + (size_t)b * P.h().nv);
}
// Coalesced variants: a CTA of 128 threads moves the q (and dq) rows of its 64 instances through a shared-memory tile with odd
// row stride (flat, contiguous global traffic), then one thread per instance integrates / checks its row in the tile.
enum { TILE_INST = 64, TILE_THREADS = 128 };
__device__ __forceinline__ void tile_rows_in(float* tile, int S, const float* __restrict__ g, int n, int len) {
  const int total = n * len;
#pragma unroll 4
  for (int k = threadIdx.x; k < total; k += TILE_THREADS) { int i = k / len; tile[i * S + (k - i * len)] = g[k]; }
}
__global__ void __launch_bounds__(TILE_THREADS) integrate_tiled_kernel(const uint32_t* __restrict__ gimage, int B, float* q, const float* __restrict__ dq, const int32_t* __restrict__ done) {
  extern __shared__ __align__(16) float tile[];
  PView P{gimage};
  const int nq = P.h().nq, nv = P.h().nv, Sq = nq | 1, Sd = nv | 1;
  const long long b0 = (long long)blockIdx.x * TILE_INST;
  const int n = (B - b0) < TILE_INST ? (int)(B - b0) : TILE_INST;
  float* tq = tile; float* td = tile + TILE_INST * Sq;
  tile_rows_in(tq, Sq, q + b0 * nq, n, nq);
  tile_rows_in(td, Sd, dq + b0 * nv, n, nv);
  __syncthreads();
  if ((int)threadIdx.x < n && !(done && done[b0 + threadIdx.x])) integrate_instance(P, tq + threadIdx.x * Sq, td + threadIdx.x * Sd);
  __syncthreads();
  float* gq = q + b0 * nq;
  const int total = n * nq;
#pragma unroll 4
  for (int k = threadIdx.x; k < total; k += TILE_THREADS) { int i = k / nq; gq[k] = tq[i * Sq + (k - i * nq)]; }
}
__global__ void __launch_bounds__(TILE_THREADS) check_limits_tiled_kernel(const uint32_t* __restrict__ gimage, int B, const float* __restrict__ q, float tol, int32_t* status, int accumulate) {
  extern __shared__ __align__(16) float tile[];
  PView P{gimage};
  const int nq = P.h().nq, Sq = nq | 1;
  const long long b0 = (long long)blockIdx.x * TILE_INST;
  const int n = (B - b0) < TILE_INST ? (int)(B - b0) : TILE_INST;
  tile_rows_in(tile, Sq, q + b0 * nq, n, nq);
  __syncthreads();
  if ((int)threadIdx.x < n) {
    int s = check_limits_instance(P, tile + threadIdx.x * Sq, tol);
    status[b0 + threadIdx.x] = accumulate ? (status[b0 + threadIdx.x] | s) : s;
  }
}
__global__ void converge_finish_kernel(int B, int max_iters, const int32_t* __restrict__ done, int32_t* iters, int32_t* status) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B && !done[b]) { iters[b] = max_iters; if (status) status[b] |= 16; }
}
__global__ void check_limits_kernel(const uint32_t* __restrict__ gimage, int B, const float* q, float tol, int32_t* status, int accumulate) {
  PView P{gimage};
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) {
    int s = check_limits_instance(P, q + (size_t)b * P.h().nq, tol);
    status[b] = accumulate ? (status[b] | s) : s;
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CUDA_OK(expr)                                                                                   \
  do {                                                                                                  \
    cudaError_t _e = (expr);                                                                            \
    if (_e != cudaSuccess) return fail(BIK_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev); else prev = -1;
    (void)cudaGetLastError();  // drop stale errors of earlier, unrelated calls on this thread
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

static int env_int(const char* name, int dflt) { const char* v = getenv(name); return v ? atoi(v) : dflt; }

struct bik_model {
  int device = 0, nsm = 0, max_smem = 0;
  int G = 4, use_tma = 1;
  HostModel hm;
  std::vector<uint32_t> image;  // model-only image (no tasks)
  uint32_t* d_image = nullptr;
};

enum { SCHED_SLOTS = 8 };
struct bik_problem {
  const bik_model* model = nullptr;
  int device = 0;  // copy: the model may be destroyed first
  std::vector<uint32_t> image;
  uint32_t* d_image = nullptr;
  PHeader h;
  int solve_double = 1;
  int k2_path = 0;  // 0 auto, 1 warp-per-instance dense only (BIK_K2_PATH=dense), 2 low-rank whenever valid (=lowrank),
                    // 3 small-group path whenever valid (=group), 4 fixed-size thread-per-problem path whenever valid (=fixed)
  mutable unsigned char* k2x_scratch = nullptr;  // per-warp H / c / box scratch of the fixed-size path
  mutable size_t k2x_scratch_bytes = 0;
  int k2_group = 4; // lanes per problem on the small-group path (BIK_K2_GROUP = 4 | 8)
  int k2_warps = 8; // warps per CTA of the K2 kernels (BIK_K2_WARPS)
  int k2_lockstep = 1; // BIK_K2_LOCKSTEP=0 disables: CTA-wide lock-step pivoting iterations
  // lazily grown scratch between K1 and K2 (one caller at a time per problem)
  std::mutex mu;
  size_t ws_B = 0;
  float *J = nullptr, *e = nullptr, *ep = nullptr, *Gc = nullptr, *hc = nullptr;
  signed char* warm = nullptr;  // [B][nu] active-set guess carried between the steps of one bik_step call
  int32_t* flags = nullptr;     // [B] instances the mixed-precision K2 hands to the fp64 kernel
  int k2_dynamic = 1;           // BIK_K2_DYNAMIC=0: static tile assignment in the small-group K2
  int k2_wide = 0;              // BIK_K2_WIDE=1: box-only problems with 33..64 coupled dofs take the small-group path (64-bit masks)
  unsigned int* d_sched = nullptr;   // SCHED_SLOTS x 32 words: {next tile, finished CTAs} per launching stream (own 128-byte line each)
  cudaStream_t sched_stream[8];
  int n_sched = 0;
  // bik_step_host staging
  size_t host_B = 0;
  float *hq = nullptr, *hft = nullptr, *hpt = nullptr, *hct = nullptr, *hdq = nullptr;
  int32_t* hst = nullptr;
  size_t host_pt_elems = 0;
  cudaStream_t hs[2] = {nullptr, nullptr};
  cudaEvent_t hev = nullptr;
  // bik_converge state
  size_t conv_B = 0;
  int32_t* conv_done = nullptr;
  float* conv_dq = nullptr;
  int* conv_count = nullptr;
  int* conv_host = nullptr;
};

// integrate / check_limits launches: tiled (coalesced) when the tile fits the default 48 KB of dynamic shared memory
static int launch_integrate(const uint32_t* d_image, const PHeader& h, int B, float* q, const float* dq, const int32_t* done, cudaStream_t st) {
  const size_t smem = (size_t)TILE_INST * ((h.nq | 1) + (h.nv | 1)) * 4;
  if (smem <= 48 * 1024) integrate_tiled_kernel<<<(B + TILE_INST - 1) / TILE_INST, TILE_THREADS, smem, st>>>(d_image, B, q, dq, done);
  else if (done) integrate_masked_kernel<<<(B + 127) / 128, 128, 0, st>>>(d_image, B, q, dq, done);
  else integrate_kernel<<<(B + 127) / 128, 128, 0, st>>>(d_image, B, q, dq);
  CUDA_OK(cudaGetLastError());
  return BIK_OK;
}
static int launch_check_limits(const uint32_t* d_image, const PHeader& h, int B, const float* q, float tol, int32_t* status, int accumulate, cudaStream_t st) {
  const size_t smem = (size_t)TILE_INST * (h.nq | 1) * 4;
  if (smem <= 48 * 1024) check_limits_tiled_kernel<<<(B + TILE_INST - 1) / TILE_INST, TILE_THREADS, smem, st>>>(d_image, B, q, tol, status, accumulate);
  else check_limits_kernel<<<(B + 127) / 128, 128, 0, st>>>(d_image, B, q, tol, status, accumulate);
  CUDA_OK(cudaGetLastError());
  return BIK_OK;
}

static int valid_group(int G) { return G == 1 || G == 2 || G == 4 || G == 8 || G == 16 || G == 32; }

extern "C" int bik_version(void) { return BIK_VERSION; }
extern "C" const char* bik_last_error(void) { return g_err.c_str(); }

extern "C" int bik_model_create(const void* blob, size_t nbytes, int device, bik_model** out) {
  if (!blob || !out) return fail(BIK_ERR_INVALID, "null argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(BIK_ERR_CUDA, "no CUDA device: libbik has no CPU path");
  if (device < 0 || device >= ndev) return fail(BIK_ERR_INVALID, "bad device index");
  bik_model* m = new bik_model;
  std::string err;
  if (!parse_model_blob(blob, nbytes, &m->hm, &err)) { delete m; return fail(BIK_ERR_INVALID, err); }
  m->device = device;
  m->G = env_int("BIK_K1_GROUP", 0);  // 0: chosen per problem from the visited tree (bik_build.h)
  if (m->G != 0 && !valid_group(m->G)) { delete m; return fail(BIK_ERR_INVALID, "BIK_K1_GROUP must be 1,2,4,8,16 or 32"); }
  m->use_tma = env_int("BIK_USE_TMA", 1);
  if (!build_image(m->hm, nullptr, 0, nullptr, 0, m->G, &m->image, &err)) { delete m; return fail(BIK_ERR_UNSUPPORTED, err); }
  DeviceGuard g(device);
  cudaDeviceProp prop;
  CUDA_OK(cudaGetDeviceProperties(&prop, device));
  m->nsm = prop.multiProcessorCount;
  m->max_smem = (int)prop.sharedMemPerBlockOptin;
  CUDA_OK(cudaMalloc(&m->d_image, m->image.size() * 4));
  CUDA_OK(cudaMemcpy(m->d_image, m->image.data(), m->image.size() * 4, cudaMemcpyHostToDevice));
  *out = m;
  return BIK_OK;
}
extern "C" void bik_model_destroy(bik_model* m) {
  if (!m) return;
  DeviceGuard g(m->device);
  cudaFree(m->d_image);
  delete m;
}

extern "C" int bik_problem_create(const bik_model* model, const bik_task_desc* tasks, int ntasks, const bik_limit_desc* limits, int nlimits,
                                  bik_problem** out) {
  if (!model || !out || ntasks < 0 || nlimits < 0 || (ntasks && !tasks) || (nlimits && !limits)) return fail(BIK_ERR_INVALID, "null argument");
  bik_problem* p = new bik_problem;
  p->model = model;
  p->device = model->device;
  std::string err;
  if (!build_image(model->hm, tasks, ntasks, limits, nlimits, model->G, &p->image, &err)) { delete p; return fail(BIK_ERR_UNSUPPORTED, err); }
  {  // launch-independent solver knob kept in the image header so that every K2 entry point sees it
    PHeader* hh = reinterpret_cast<PHeader*>(p->image.data());
    hh->k2_sweeps = std::max(0, std::min(16, env_int("BIK_K2_SWEEPS", hh->k2_sweeps)));
    hh->k2_rule = env_int("BIK_K2_RULE", hh->k2_rule) != 0;
  }
  memcpy(&p->h, p->image.data(), sizeof(PHeader));
  const char* prec = getenv("BIK_SOLVE_PRECISION");
  p->solve_double = !(prec && (std::string(prec) == "f32" || std::string(prec) == "float"));
  const char* path = getenv("BIK_K2_PATH");
  p->k2_path = (path && std::string(path) == "dense") ? 1 : ((path && std::string(path) == "lowrank") ? 2 : ((path && std::string(path) == "group") ? 3 : ((path && std::string(path) == "fixed") ? 4 : 0)));
  p->k2_group = env_int("BIK_K2_GROUP", p->h.nu > 8 ? 8 : 4) == 8 ? 8 : 4;
  p->k2_warps = env_int("BIK_K2_WARPS", 8);
  p->k2_lockstep = env_int("BIK_K2_LOCKSTEP", 1);
  p->k2_dynamic = env_int("BIK_K2_DYNAMIC", 1) != 0;
  p->k2_wide = env_int("BIK_K2_WIDE", 0) != 0;
  if (p->k2_warps != 1 && p->k2_warps != 2 && p->k2_warps != 4 && p->k2_warps != 8) p->k2_warps = 8;
  DeviceGuard g(model->device);
  CUDA_OK(cudaMalloc(&p->d_image, p->image.size() * 4));
  CUDA_OK(cudaMemcpy(p->d_image, p->image.data(), p->image.size() * 4, cudaMemcpyHostToDevice));
  *out = p;
  return BIK_OK;
}
extern "C" void bik_problem_destroy(bik_problem* p) {
  if (!p) return;
  DeviceGuard g(p->device);
  cudaFree(p->d_sched); cudaFree(p->d_image); cudaFree(p->J); cudaFree(p->e); cudaFree(p->ep); cudaFree(p->Gc); cudaFree(p->hc); cudaFree(p->warm); cudaFree(p->flags); cudaFree(p->k2x_scratch);
  cudaFree(p->hq); cudaFree(p->hft); cudaFree(p->hpt); cudaFree(p->hct); cudaFree(p->hdq); cudaFree(p->hst);
  cudaFree(p->conv_done); cudaFree(p->conv_dq); cudaFree(p->conv_count);
  if (p->conv_host) cudaFreeHost(p->conv_host);
  if (p->hs[0]) { cudaStreamDestroy(p->hs[0]); cudaStreamDestroy(p->hs[1]); cudaEventDestroy(p->hev); }
  delete p;
}
extern "C" int bik_problem_dims(const bik_problem* p, bik_dims* out) {
  if (!p || !out) return fail(BIK_ERR_INVALID, "null argument");
  out->nq = p->h.nq; out->nv = p->h.nv; out->nnode = p->h.nnode; out->nframe = p->h.F; out->nposture = p->h.P; out->ncom = p->h.C;
  out->nrows = p->h.K; out->npairs = p->h.npairs;
  return BIK_OK;
}
extern "C" size_t bik_workspace_bytes(const bik_problem* p, int B) {
  if (!p || B <= 0) return 0;
  const PHeader& h = p->h;
  return (size_t)B * 4 * ((size_t)h.K * h.nv + h.K + (size_t)h.P * h.nv + (size_t)h.npairs * (h.nv + 1) + 4);
}

// (kernel, smem, threads) -> resident CTAs per SM; the attribute/occupancy queries run once per combination.
struct GeomKey { const void* kern; size_t smem; int threads; int device; };
static std::mutex g_geom_mu;
static std::vector<std::pair<GeomKey, int>> g_geom;

template <typename Kern>
static int launch_geometry(Kern kern, const bik_model* m, size_t smem, int threads, long long work_ctas, int* grid) {
  if ((int)smem > m->max_smem) return fail(BIK_ERR_UNSUPPORTED, "problem does not fit in shared memory (" + std::to_string(smem) + " B)");
  int per_sm = 0;
  {
    std::lock_guard<std::mutex> lock(g_geom_mu);
    for (auto& e : g_geom)
      if (e.first.kern == (const void*)kern && e.first.smem == smem && e.first.threads == threads && e.first.device == m->device) per_sm = e.second;
  }
  if (per_sm == 0) {
    // the opt-in limit is per kernel and must never shrink below what an earlier, cached combination needs
    size_t prev_max = 0;
    {
      std::lock_guard<std::mutex> lock(g_geom_mu);
      for (auto& e : g_geom) if (e.first.kern == (const void*)kern && e.first.device == m->device && e.first.smem > prev_max) prev_max = e.first.smem;
    }
    if (smem > prev_max) CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
    if (per_sm < 1) return fail(BIK_ERR_UNSUPPORTED, "kernel cannot be resident");
    std::lock_guard<std::mutex> lock(g_geom_mu);
    g_geom.push_back({GeomKey{(const void*)kern, smem, threads, m->device}, per_sm});
  }
  long long g = (long long)m->nsm * per_sm;  // persistent: one wave, CTAs loop over tiles
  if (work_ctas < g) g = work_ctas;
  *grid = (int)(g < 1 ? 1 : g);
  return BIK_OK;
}

template <int G>
static int launch_k1(const bik_problem* p, const K1Args& a, cudaStream_t st) {
  const PHeader& h = p->h;
  constexpr int IPW = 32 / G;
  int NW = 4;  // warps per CTA: as many as fit next to the image in shared memory
  auto need = [&](int nw) { return (size_t)h.words * 4 + (size_t)nw * k1_warp_words(h, IPW) * 4; };
  while (NW > 1 && (int)need(NW) > p->model->max_smem) NW >>= 1;
  size_t smem = need(NW);
  int grid = 1;
  long long tiles = ((long long)a.B + IPW - 1) / IPW;
  int rc = launch_geometry(k1_kernel<G>, p->model, smem, 32 * NW, (tiles + NW - 1) / NW, &grid);
  if (rc) return rc;
  k1_kernel<G><<<grid, 32 * NW, smem, st>>>(p->d_image, h.words, p->model->use_tma, a);
  CUDA_OK(cudaGetLastError());
  return BIK_OK;
}
static int dispatch_k1(const bik_problem* p, const K1Args& a, cudaStream_t st) {
  switch (p->h.G) {
    case 1: return launch_k1<1>(p, a, st);
    case 2: return launch_k1<2>(p, a, st);
    case 4: return launch_k1<4>(p, a, st);
    case 8: return launch_k1<8>(p, a, st);
    case 16: return launch_k1<16>(p, a, st);
    default: return launch_k1<32>(p, a, st);
  }
}
template <typename T, int SLOTS>
static int launch_k2(const bik_problem* p, const K2Args& a, cudaStream_t st) {
  const PHeader& h = p->h;
  int NW = p->k2_warps;
  auto need = [&](int nw) { return (size_t)h.words * 4 + (size_t)nw * k2_warp_bytes(h, sizeof(T)); };
  while (NW > 1 && (int)need(NW) > p->model->max_smem) NW >>= 1;
  size_t smem = need(NW);
  int grid = 1;
  int rc = launch_geometry(k2_kernel<T, SLOTS>, p->model, smem, 32 * NW, ((long long)a.B + NW - 1) / NW, &grid);
  if (rc) return rc;
  k2_kernel<T, SLOTS><<<grid, 32 * NW, smem, st>>>(p->d_image, h.words, p->model->use_tma, a);
  CUDA_OK(cudaGetLastError());
  return BIK_OK;
}
template <typename T>
static int dispatch_k2_slots(const bik_problem* p, const K2Args& a, cudaStream_t st) {
  int slots = (p->h.nu + 1 + 31) / 32;  // rows of the augmented factor (coupled dofs + rhs) per lane
  if (slots <= 1) return launch_k2<T, 1>(p, a, st);
  if (slots == 2) return launch_k2<T, 2>(p, a, st);
  return launch_k2<T, 3>(p, a, st);
}
template <int SLOTS>
static int launch_k2lr(const bik_problem* p, const K2Args& a, cudaStream_t st) {
  const PHeader& h = p->h;
  int NW = p->k2_warps;
  auto need = [&](int nw) { return (size_t)h.words * 4 + ((tri(h.K) * 2 + 15) & ~15) + (size_t)nw * k2lr_warp_bytes(h); };
  while (NW > 1 && (int)need(NW) > p->model->max_smem) NW >>= 1;
  size_t smem = need(NW);
  int grid = 1;
  int rc = launch_geometry(k2lr_kernel<SLOTS>, p->model, smem, 32 * NW, ((long long)a.B + NW - 1) / NW, &grid);
  if (rc) return rc;
  k2lr_kernel<SLOTS><<<grid, 32 * NW, smem, st>>>(p->d_image, h.words, p->model->use_tma, a);
  CUDA_OK(cudaGetLastError());
  return BIK_OK;
}
// Low-rank path when the stacked task rows are well below the number of COUPLED dofs (e.g. a CoM task:
// 3 rows coupling every dof), no general rows, D bounded away from 0.  Otherwise the dense path on the
// coupled block wins (G1 config: K = 18 rows, 18 coupled dofs of 43).
static bool use_low_rank(const bik_problem* p, const K2Args& a) {
  const PHeader& h = p->h;
  if (p->k2_path == 1 || !a.dq || a.Hout || a.lo_out) return false;
  if (p->k2_path == 2) return h.npairs == 0 && h.K > 0 && h.K < 63 && a.damping > 0;
  return h.npairs == 0 && h.K > 0 && h.K < 63 && 2 * h.K <= h.nu && a.damping >= 1e-6;
}
static bool use_low_rank_static(const bik_problem* p, double damping) {
  K2Args a;
  memset(&a, 0, sizeof a);
  a.dq = reinterpret_cast<float*>(1); a.damping = damping;
  return use_low_rank(p, a);
}
// One (next tile, finished CTAs) pair per stream that launches the small-group K2 of this problem: launches on one stream
// are ordered, launches on different streams (bik_step_host alternates two) must not share a counter.  More than
// SCHED_SLOTS streams, or BIK_K2_DYNAMIC=0: static round-robin.
static unsigned int* tile_counter(const bik_problem* cp, cudaStream_t st) {
  bik_problem* p = const_cast<bik_problem*>(cp);
  if (!p->k2_dynamic) return nullptr;
  if (!p->d_sched) {
    if (cudaMalloc(&p->d_sched, SCHED_SLOTS * 32 * sizeof(unsigned int)) != cudaSuccess) { cudaGetLastError(); p->k2_dynamic = 0; return nullptr; }
    cudaMemset(p->d_sched, 0, SCHED_SLOTS * 32 * sizeof(unsigned int));
  }
  for (int i = 0; i < p->n_sched; ++i) if (p->sched_stream[i] == st) return p->d_sched + 32 * i;
  if (p->n_sched == SCHED_SLOTS) return nullptr;
  p->sched_stream[p->n_sched] = st;
  return p->d_sched + 32 * p->n_sched++;
}
template <typename T, int G, typename M>
static int launch_k2t(const bik_problem* p, const K2Args& a, cudaStream_t st) {
  constexpr int NS = 32 / G, MAXT = sizeof(T) == 8 ? 256 : 512;
  PView P{p->image.data()};
  const size_t wb = (size_t)k2t_warp_bytes(P, sizeof(T), NS);
  int NW = MAXT / 32;
  while (NW > 1 && NW * wb > (size_t)p->model->max_smem) --NW;
  const size_t smem = NW * wb;
  int grid = 1;
  long long tiles = ((long long)a.B + NS - 1) / NS;
  int rc = launch_geometry(k2t_kernel<T, G, MAXT, M>, p->model, smem, 32 * NW, (tiles + NW - 1) / NW, &grid);
  if (rc) return rc;
  k2t_kernel<T, G, MAXT, M><<<grid, 32 * NW, smem, st>>>(p->d_image, (int)wb, a, tile_counter(p, st));
  CUDA_OK(cudaGetLastError());
  return BIK_OK;
}
template <typename T>
static int dispatch_k2t(const bik_problem* p, const K2Args& a, cudaStream_t st) {
  if (p->h.nu > K2T_NMAX) return launch_k2t<T, 8, uint64_t>(p, a, st);   // wide: 64-bit active-set masks (BIK_K2_WIDE=1)
  return p->k2_group == 8 ? launch_k2t<T, 8, uint32_t>(p, a, st) : launch_k2t<T, 4, uint32_t>(p, a, st);
}
static int k2x_size(int nu) {
  const int sizes[] = {6, 8, 12, 16, 18, 20, 24};
  for (int n : sizes) if (nu <= n) return n;
  return 0;
}
template <typename TL, typename TH, int N>
static int launch_k2x(const bik_problem* p, const K2Args& a, cudaStream_t st) {
  PView P{p->image.data()};
  const size_t wb = (size_t)k2x_warp_smem_bytes(P, sizeof(TL), sizeof(TH), N, 32);
  int NW = 8;
  while (NW > 1 && NW * wb > (size_t)p->model->max_smem) --NW;
  const size_t smem = NW * wb;
  int grid = 1;
  long long tiles = ((long long)a.B + 31) / 32;
  int rc = launch_geometry(k2x_kernel<TL, TH, N>, p->model, smem, 32 * NW, (tiles + NW - 1) / NW, &grid);
  if (rc) return rc;
  const size_t ws = k2x_warp_scratch_bytes(sizeof(TH), N, 32), need = ws * NW * (size_t)grid;
  if (need > p->k2x_scratch_bytes) {
    CUDA_OK(cudaDeviceSynchronize());
    cudaFree(p->k2x_scratch); p->k2x_scratch = nullptr; p->k2x_scratch_bytes = 0;
    CUDA_OK(cudaMalloc(&p->k2x_scratch, need));
    p->k2x_scratch_bytes = need;
  }
  k2x_kernel<TL, TH, N><<<grid, 32 * NW, smem, st>>>(p->d_image, (int)wb, p->k2x_scratch, (unsigned long long)ws, a);
  CUDA_OK(cudaGetLastError());
  return BIK_OK;
}
template <typename TL, typename TH>
static int dispatch_k2x(const bik_problem* p, const K2Args& a, cudaStream_t st) {
  switch (k2x_size(p->h.nu)) {
    case 6: return launch_k2x<TL, TH, 6>(p, a, st);
    case 8: return launch_k2x<TL, TH, 8>(p, a, st);
    case 12: return launch_k2x<TL, TH, 12>(p, a, st);
    case 16: return launch_k2x<TL, TH, 16>(p, a, st);
    case 18: return launch_k2x<TL, TH, 18>(p, a, st);
    case 20: return launch_k2x<TL, TH, 20>(p, a, st);
    default: return launch_k2x<TL, TH, 24>(p, a, st);
  }
}
// Fixed-size path: box-only problems with at most 24 coupled dofs.  Two instantiations: all fp32 (chosen for
// BIK_SOLVE_PRECISION=f32: G1 0.72 ms against 1.2 ms for the small-group path in fp32) and mixed precision (fp32 factorisations
// of the fp64 H, polished and re-judged in fp64, flagged instances handed to the fp64 small-group kernel; fp64-level results
// but 1.94 ms on G1 against 1.29 ms for the small-group fp64 path -- its fp64 vectors spill -- so it only runs when forced
// with BIK_K2_PATH=fixed).  An all-fp64 instantiation was slower still (2.2 ms) and is not built.
static bool use_fixed(const bik_problem* p, const K2Args& a) {
  const PHeader& h = p->h;
  if (p->k2_path == 1 || p->k2_path == 2 || p->k2_path == 3 || !a.dq || a.Hout || a.lo_out) return false;
  if (p->solve_double && p->k2_path != 4) return false;
  if (h.npairs != 0 || h.nu < 1 || k2x_size(h.nu) == 0) return false;
  PView P{p->image.data()};
  return k2x_warp_smem_bytes(P, 4, p->solve_double ? 8 : 4, k2x_size(h.nu), 32) <= p->model->max_smem;
}
// Small-group path: box-only problems whose coupled block is small enough (at least one warp of them must fit in
// an SM's shared memory).
static bool use_thread(const bik_problem* p, const K2Args& a) {
  const PHeader& h = p->h;
  if (p->k2_path == 1 || p->k2_path == 2 || p->k2_path == 4 || !a.dq || a.Hout || a.lo_out) return false;
  if (h.npairs != 0 || h.nu < 1 || h.nu > (p->k2_wide ? K2T_NMAX_WIDE : K2T_NMAX)) return false;
  PView P{p->image.data()};
  return k2t_warp_bytes(P, p->solve_double ? 8 : 4, 32 / (p->k2_group == 8 ? 8 : 4)) <= p->model->max_smem;
}
static int dispatch_k2(const bik_problem* p, const K2Args& a, cudaStream_t st) {
  if (use_fixed(p, a)) {
    if (!p->solve_double) return dispatch_k2x<float, float>(p, a, st);
    // mixed precision, then the fp64 small-group kernel on the instances it flagged (normally none: the launch returns at once)
    K2Args am = a;
    am.flag_out = a.flag_out;   // caller passes the workspace slice (dispatch sites below)
    int rc = dispatch_k2x<float, double>(p, am, st);
    if (rc || !a.flag_out) return rc;
    K2Args af = a;
    af.flag_out = nullptr; af.only = a.flag_out;
    return dispatch_k2t<double>(p, af, st);
  }
  if (use_thread(p, a)) return p->solve_double ? dispatch_k2t<double>(p, a, st) : dispatch_k2t<float>(p, a, st);
  if (use_low_rank(p, a)) return (p->h.K + 1 <= 32) ? launch_k2lr<1>(p, a, st) : launch_k2lr<2>(p, a, st);
  return p->solve_double ? dispatch_k2_slots<double>(p, a, st) : dispatch_k2_slots<float>(p, a, st);
}

template <int G>
static int launch_fk(const bik_model* m, const FkArgs& a, cudaStream_t st) {
  PHeader h;
  memcpy(&h, m->image.data(), sizeof h);
  constexpr int IPW = 32 / G, THREADS = 128, NW = THREADS / 32;
  size_t wwords = (IPW * ((7 * h.nnode) | 1) + 3) & ~3;
  size_t smem = (size_t)h.words * 4 + NW * wwords * 4;
  int grid = 1;
  long long tiles = ((long long)a.B + IPW - 1) / IPW;
  int rc = launch_geometry(fk_kernel<G>, m, smem, THREADS, (tiles + NW - 1) / NW, &grid);
  if (rc) return rc;
  fk_kernel<G><<<grid, THREADS, smem, st>>>(m->d_image, h.words, m->use_tma, a);
  CUDA_OK(cudaGetLastError());
  return BIK_OK;
}
static int dispatch_fk(const bik_model* m, const FkArgs& a, cudaStream_t st) {
  PHeader mh;
  memcpy(&mh, m->image.data(), sizeof mh);
  switch (mh.G) {
    case 1: return launch_fk<1>(m, a, st);
    case 2: return launch_fk<2>(m, a, st);
    case 4: return launch_fk<4>(m, a, st);
    case 8: return launch_fk<8>(m, a, st);
    case 16: return launch_fk<16>(m, a, st);
    default: return launch_fk<32>(m, a, st);
  }
}

static int fk_common(const bik_model* model, int B, const float* q, const bik_frame* frames, int nframes, float* poses, float* com, float* J, void* stream) {
  if (!model || B < 0 || !q || nframes < 0 || (nframes && !frames)) return fail(BIK_ERR_INVALID, "null argument");
  if (B == 0) return BIK_OK;
  DeviceGuard g(model->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // frames travel by value in the kernel arguments (the Configuration API asks for one or two)
  if (nframes > 16) return fail(BIK_ERR_UNSUPPORTED, "at most 16 frames per bik_fk / bik_frame_jacobian call");
  FkArgs a;
  memset(&a, 0, sizeof a);
  a.B = B; a.nframes = nframes; a.q = q; a.poses = poses; a.com = com; a.J = J;
  for (int f = 0; f < nframes; ++f) {
    if (frames[f].node >= model->hm.nnode) return fail(BIK_ERR_INVALID, "frame node out of range");
    put_frame(frames[f], &a.frames[f].node, a.frames[f].lpos, a.frames[f].lquat);
  }
  return dispatch_fk(model, a, st);
}
extern "C" int bik_fk(const bik_model* model, int B, const float* q, const bik_frame* frames, int nframes, float* poses, float* com, void* stream) {
  return fk_common(model, B, q, frames, nframes, poses, com, nullptr, stream);
}
extern "C" int bik_frame_jacobian(const bik_model* model, int B, const float* q, const bik_frame* frames, int nframes, float* J, void* stream) {
  return fk_common(model, B, q, frames, nframes, nullptr, nullptr, J, stream);
}

static int check_inputs(const bik_problem* p, const bik_inputs* in, bool need_q) {
  if (!p || !in) return fail(BIK_ERR_INVALID, "null argument");
  const PHeader& h = p->h;
  if (need_q && !in->q) return fail(BIK_ERR_INVALID, "inputs.q is null");
  if (h.F > 0 && !in->frame_targets) return fail(BIK_ERR_INVALID, "No target set for FrameTask");
  if (h.P > 0 && !in->posture_targets) return fail(BIK_ERR_INVALID, "No target set for PostureTask");
  if (h.C > 0 && !in->com_targets) return fail(BIK_ERR_INVALID, "No target set for ComTask");
  return BIK_OK;
}

extern "C" int bik_fk_jac(const bik_problem* p, int B, const bik_inputs* in, float dt, float* J, float* e, float* e_posture, float* G_coll,
                          float* h_coll, void* stream) {
  if (B == 0 && p) return BIK_OK;
  int rc = check_inputs(p, in, true);
  if (rc) return rc;
  const PHeader& h = p->h;
  if (B < 0 || (h.K > 0 && (!J || !e)) || (h.P > 0 && !e_posture) || (h.npairs > 0 && (!G_coll || !h_coll))) return fail(BIK_ERR_INVALID, "null output");
  if (B == 0) return BIK_OK;
  DeviceGuard g(p->model->device);
  K1Args a{B, in->q, in->frame_targets, in->posture_targets, in->com_targets, in->posture_batched, dt, J, e, e_posture, G_coll, h_coll};
  return dispatch_k1(p, a, static_cast<cudaStream_t>(stream));
}

extern "C" int bik_qp_objective(const bik_problem* p, int B, const float* J, const float* e, const float* e_posture, double damping, double* H,
                                double* c, void* stream) {
  if (!p || !H || !c || B < 0) return fail(BIK_ERR_INVALID, "null argument");
  if (B == 0) return BIK_OK;
  DeviceGuard g(p->model->device);
  K2Args a;
  memset(&a, 0, sizeof a);
  a.B = B; a.J = J; a.e = e; a.ep = e_posture; a.dt = 1.f; a.damping = damping; a.Hout = H; a.cout = c;
  a.skip_box = 1;
  return dispatch_k2(p, a, static_cast<cudaStream_t>(stream));
}

extern "C" int bik_limits_box(const bik_problem* p, int B, const float* q, float dt, float* lo, float* hi, void* stream) {
  if (!p || !q || !lo || !hi || B < 0) return fail(BIK_ERR_INVALID, "null argument");
  if (B == 0) return BIK_OK;
  DeviceGuard g(p->model->device);
  K2Args a;
  memset(&a, 0, sizeof a);
  a.B = B; a.q = q; a.dt = dt; a.lo_out = lo; a.hi_out = hi; a.skip_objective = 1;
  return dispatch_k2(p, a, static_cast<cudaStream_t>(stream));
}

static int ensure_workspace(bik_problem* p, int B);
extern "C" int bik_solve_ex(const bik_problem* p, int B, const float* q, const float* J, const float* e, const float* e_posture, const float* G_coll,
                            const float* h_coll, float dt, double damping, float* dq, int32_t* status, int32_t* iters, void* stream);
extern "C" int bik_solve(const bik_problem* p, int B, const float* q, const float* J, const float* e, const float* e_posture, const float* G_coll,
                         const float* h_coll, float dt, double damping, float* dq, int32_t* status, void* stream) {
  return bik_solve_ex(p, B, q, J, e, e_posture, G_coll, h_coll, dt, damping, dq, status, nullptr, stream);
}
extern "C" int bik_solve_ex(const bik_problem* p, int B, const float* q, const float* J, const float* e, const float* e_posture, const float* G_coll,
                            const float* h_coll, float dt, double damping, float* dq, int32_t* status, int32_t* iters, void* stream) {
  if (!p || !q || !dq || B < 0) return fail(BIK_ERR_INVALID, "null argument");
  const PHeader& h = p->h;
  if ((h.K > 0 && (!J || !e)) || (h.P > 0 && !e_posture) || (h.npairs > 0 && (!G_coll || !h_coll))) return fail(BIK_ERR_INVALID, "null input");
  if (B == 0) return BIK_OK;
  DeviceGuard g(p->model->device);
  K2Args a;
  memset(&a, 0, sizeof a);
  a.B = B; a.q = q; a.J = J; a.e = e; a.ep = e_posture; a.Gc = G_coll; a.hc = h_coll; a.dt = dt; a.damping = damping; a.dq = dq; a.status = status; a.iters = iters; a.lockstep = p->k2_lockstep;
  if (status) CUDA_OK(cudaMemsetAsync(status, 0, sizeof(int32_t) * (size_t)B, static_cast<cudaStream_t>(stream)));
  if (use_fixed(p, a) && p->solve_double) {   // the mixed-precision path hands flagged instances over through the workspace
    bik_problem* mp = const_cast<bik_problem*>(p);
    std::lock_guard<std::mutex> lock(mp->mu);
    int rc = ensure_workspace(mp, B);
    if (rc) return rc;
    a.flag_out = mp->flags;
    return dispatch_k2(p, a, static_cast<cudaStream_t>(stream));
  }
  return dispatch_k2(p, a, static_cast<cudaStream_t>(stream));
}

extern "C" int bik_integrate(const bik_model* m, int B, float* q, const float* dq, void* stream) {
  if (!m || !q || !dq || B < 0) return fail(BIK_ERR_INVALID, "null argument");
  if (B == 0) return BIK_OK;
  DeviceGuard g(m->device);
  PHeader mh;
  memcpy(&mh, m->image.data(), sizeof mh);
  return launch_integrate(m->d_image, mh, B, q, dq, nullptr, static_cast<cudaStream_t>(stream));
}
extern "C" int bik_check_limits(const bik_model* m, int B, const float* q, float tol, int32_t* status, void* stream) {
  if (!m || !q || !status || B < 0) return fail(BIK_ERR_INVALID, "null argument");
  if (B == 0) return BIK_OK;
  DeviceGuard g(m->device);
  PHeader mh;
  memcpy(&mh, m->image.data(), sizeof mh);
  return launch_check_limits(m->d_image, mh, B, q, tol, status, 0, static_cast<cudaStream_t>(stream));
}

static int ensure_workspace(bik_problem* p, int B) {
  if ((size_t)B <= p->ws_B) return BIK_OK;
  const PHeader& h = p->h;
  cudaFree(p->J); cudaFree(p->e); cudaFree(p->ep); cudaFree(p->Gc); cudaFree(p->hc); cudaFree(p->warm); cudaFree(p->flags);
  p->J = p->e = p->ep = p->Gc = p->hc = nullptr; p->warm = nullptr; p->flags = nullptr; p->ws_B = 0;
  size_t b = (size_t)B;
  CUDA_OK(cudaMalloc(&p->J, sizeof(float) * b * (h.K > 0 ? h.K : 1) * h.nv));
  CUDA_OK(cudaMalloc(&p->e, sizeof(float) * b * (h.K > 0 ? h.K : 1)));
  CUDA_OK(cudaMalloc(&p->ep, sizeof(float) * b * (h.P > 0 ? h.P : 1) * h.nv));
  CUDA_OK(cudaMalloc(&p->Gc, sizeof(float) * b * (h.npairs > 0 ? h.npairs : 1) * h.nv));
  CUDA_OK(cudaMalloc(&p->hc, sizeof(float) * b * (h.npairs > 0 ? h.npairs : 1)));
  CUDA_OK(cudaMalloc(&p->warm, b * (size_t)(h.nu > 0 ? h.nu : 1)));
  CUDA_OK(cudaMalloc(&p->flags, sizeof(int32_t) * b));
  p->ws_B = b;
  return BIK_OK;
}

// One or more solve_ik steps on device buffers; the K1 -> K2 workspace rows [ws_off, ws_off + B) are used, so that
// chunks of one batch can be in flight on different streams (bik_step_host).  Caller holds p->mu and has sized the workspace.
static int step_core(bik_problem* p, int B, size_t ws_off, float* q, const bik_inputs* in, float dt, double damping, int nsteps, int integrate,
                     float* dq, int32_t* status, cudaStream_t st) {
  const PHeader& h = p->h;
  const size_t Kd = h.K > 0 ? h.K : 1, Pd = h.P > 0 ? h.P : 1, Nd = h.npairs > 0 ? h.npairs : 1, Ud = h.nu > 0 ? h.nu : 1;
  float* J = p->J + ws_off * Kd * h.nv; float* e = p->e + ws_off * Kd; float* ep = p->ep + ws_off * Pd * h.nv;
  float* Gc = p->Gc + ws_off * Nd * h.nv; float* hc = p->hc + ws_off * Nd;
  signed char* warm_buf = p->warm + ws_off * Ud;
  const bool warm = nsteps > 1 && !use_low_rank_static(p, damping);   // rollouts: carry the active set from step to step
  if (warm) CUDA_OK(cudaMemsetAsync(warm_buf, 0, (size_t)B * Ud, st));
  for (int s = 0; s < nsteps; ++s) {
    if (status) {  // Configuration.check_limits(safety_break=False) of solve_ik.py:99
      int rcl = launch_check_limits(p->d_image, h, B, q, 1e-6f, status, s > 0, st);
      if (rcl) return rcl;
    }
    K1Args a1{B, q, in->frame_targets, in->posture_targets, in->com_targets, in->posture_batched, dt, J, e, ep, Gc, hc};
    int rc = dispatch_k1(p, a1, st);
    if (rc) return rc;
    K2Args a2;
    memset(&a2, 0, sizeof a2);
    a2.B = B; a2.q = q; a2.J = J; a2.e = e; a2.ep = ep; a2.Gc = Gc; a2.hc = hc; a2.dt = dt; a2.damping = damping; a2.dq = dq; a2.status = status; a2.lockstep = p->k2_lockstep;
    a2.warm = warm ? warm_buf : nullptr;
    a2.flag_out = p->flags + ws_off;
    rc = dispatch_k2(p, a2, st);
    if (rc) return rc;
    if (integrate) {
      rc = launch_integrate(p->d_image, h, B, q, dq, nullptr, st);
      if (rc) return rc;
    }
  }
  return BIK_OK;
}

extern "C" int bik_step(const bik_problem* cp, int B, float* q, const bik_inputs* in, float dt, double damping, int nsteps, int integrate, float* dq,
                        int32_t* status, void* stream) {
  if (B == 0 && cp) return BIK_OK;
  int rc = check_inputs(cp, in, false);
  if (rc) return rc;
  if (!q || !dq || B < 0 || nsteps < 1) return fail(BIK_ERR_INVALID, "bad argument");
  bik_problem* p = const_cast<bik_problem*>(cp);
  std::lock_guard<std::mutex> lock(p->mu);
  DeviceGuard g(p->model->device);
  rc = ensure_workspace(p, B);
  if (rc) return rc;
  return step_core(p, B, 0, q, in, dt, damping, nsteps, integrate, dq, status, static_cast<cudaStream_t>(stream));
}

extern "C" int bik_problem_describe(const bik_problem* p, double damping, char* buf, size_t cap) {
  if (!p) return fail(BIK_ERR_INVALID, "null argument");
  const PHeader& h = p->h;
  K2Args a;
  memset(&a, 0, sizeof a);
  a.dq = reinterpret_cast<float*>(1); a.damping = damping;
  std::string k2;
  if (use_fixed(p, a)) k2 = "fixed-size thread-per-problem N=" + std::to_string(k2x_size(h.nu));
  else if (use_thread(p, a)) k2 = "small-group G=" + std::to_string(p->k2_group);
  else if (use_low_rank(p, a)) k2 = "low-rank warp-per-problem";
  else k2 = "dense warp-per-problem";
  std::string d = "k1: " + std::to_string(h.G) + " lanes/instance, " + std::to_string(h.nneeded) + "/" + std::to_string(h.nnode) + " nodes visited; nv=" +
                  std::to_string(h.nv) + " coupled=" + std::to_string(h.nu) + " (unbounded, eliminated once: " + std::to_string(use_thread(p, a) && !use_fixed(p, a) ? h.nfree : 0) + ") rows=" + std::to_string(h.K) + " pairs=" + std::to_string(h.npairs) + "; k2: " + k2 +
                  (p->solve_double ? " f64" : " f32");
  if (buf && cap) { size_t n = d.size() < cap - 1 ? d.size() : cap - 1; memcpy(buf, d.c_str(), n); buf[n] = 0; }
  return (int)d.size();
}

// solve_ik + integrate until every frame task of an instance is within (pos_threshold, ori_threshold) or max_iters steps were
// taken -- the inner loop of the reference's examples (examples/quadruped_spot.py:89-104, examples/arm_aloha.py:146-169), per
// instance.  An instance that has converged keeps its q; the batch loop ends as soon as the device counter of unconverged
// instances reads zero (checked every `check_every` steps: one 4-byte D2H copy and a stream synchronisation).
extern "C" int bik_converge(const bik_problem* cp, int B, float* q, const bik_inputs* in, float dt, double damping, int max_iters, float pos_threshold,
                            float ori_threshold, int check_every, int32_t* iters, int32_t* status, void* stream) {
  if (B == 0 && cp) return BIK_OK;
  int rc = check_inputs(cp, in, false);
  if (rc) return rc;
  if (!q || !iters || B < 0 || max_iters < 1) return fail(BIK_ERR_INVALID, "bad argument");
  if (check_every < 1) check_every = 1;
  bik_problem* p = const_cast<bik_problem*>(cp);
  std::lock_guard<std::mutex> lock(p->mu);
  DeviceGuard g(p->model->device);
  rc = ensure_workspace(p, B);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const PHeader& h = p->h;
  if ((size_t)B > p->conv_B) {
    cudaFree(p->conv_done); cudaFree(p->conv_dq); cudaFree(p->conv_count);
    p->conv_done = nullptr; p->conv_dq = nullptr; p->conv_count = nullptr; p->conv_B = 0;
    CUDA_OK(cudaMalloc(&p->conv_done, sizeof(int32_t) * (size_t)B));
    CUDA_OK(cudaMalloc(&p->conv_dq, sizeof(float) * (size_t)B * h.nv));
    CUDA_OK(cudaMalloc(&p->conv_count, sizeof(int)));
    p->conv_B = (size_t)B;
  }
  if (!p->conv_host) CUDA_OK(cudaMallocHost(&p->conv_host, sizeof(int)));
  CUDA_OK(cudaMemsetAsync(p->conv_done, 0, sizeof(int32_t) * (size_t)B, st));
  const bool warm = !use_low_rank_static(p, damping);
  if (warm) CUDA_OK(cudaMemsetAsync(p->warm, 0, (size_t)B * (size_t)(h.nu > 0 ? h.nu : 1), st));
  const int blocks = (B + 127) / 128;
  bool all_done = false;
  for (int it = 0; it <= max_iters && !all_done; ++it) {
    if (status && it < max_iters) {
      rc = launch_check_limits(p->d_image, h, B, q, 1e-6f, status, it > 0, st);
      if (rc) return rc;
    }
    K1Args a1{B, q, in->frame_targets, in->posture_targets, in->com_targets, in->posture_batched, dt, p->J, p->e, p->ep, p->Gc, p->hc};
    rc = dispatch_k1(p, a1, st);
    if (rc) return rc;
    if (it > 0) {   // errors at the configuration reached after `it` steps
      CUDA_OK(cudaMemsetAsync(p->conv_count, 0, sizeof(int), st));
      converge_check_kernel<<<blocks, 128, 0, st>>>(p->d_image, B, p->e, it, pos_threshold, ori_threshold, p->conv_done, iters, p->conv_count);
      CUDA_OK(cudaGetLastError());
      if (it == max_iters || it % check_every == 0) {
        CUDA_OK(cudaMemcpyAsync(p->conv_host, p->conv_count, sizeof(int), cudaMemcpyDeviceToHost, st));
        CUDA_OK(cudaStreamSynchronize(st));
        all_done = *p->conv_host == 0;
      }
    }
    if (it == max_iters || all_done) break;
    K2Args a2;
    memset(&a2, 0, sizeof a2);
    a2.B = B; a2.q = q; a2.J = p->J; a2.e = p->e; a2.ep = p->ep; a2.Gc = p->Gc; a2.hc = p->hc; a2.dt = dt; a2.damping = damping; a2.dq = p->conv_dq; a2.status = status;
    a2.lockstep = p->k2_lockstep; a2.warm = warm ? p->warm : nullptr; a2.flag_out = p->flags;
    rc = dispatch_k2(p, a2, st);
    if (rc) return rc;
    rc = launch_integrate(p->d_image, h, B, q, p->conv_dq, p->conv_done, st);
    if (rc) return rc;
  }
  // instances that never met the thresholds: iters = max_iters, status bit BIK_STATUS_NOT_CONVERGED
  converge_finish_kernel<<<blocks, 128, 0, st>>>(B, max_iters, p->conv_done, iters, status);
  CUDA_OK(cudaGetLastError());
  return BIK_OK;
}

// Instances one resident wave of the K2 kernel covers (0 when unknown): bik_step_host cuts its chunks at multiples of it.
static long long k2_wave_instances(const bik_problem* p) {
  K2Args a;
  memset(&a, 0, sizeof a);
  a.dq = reinterpret_cast<float*>(1); a.damping = 1.0;
  if (!use_thread(p, a) || use_fixed(p, a)) return 0;
  PView P{p->image.data()};
  const int G = p->k2_group == 8 ? 8 : 4, NS = 32 / G, ts = p->solve_double ? 8 : 4, maxt = p->solve_double ? 256 : 512;
  const size_t wb = (size_t)k2t_warp_bytes(P, ts, NS);
  int NW = maxt / 32;
  while (NW > 1 && NW * wb > (size_t)p->model->max_smem) --NW;
  int per_sm = (int)((size_t)p->model->max_smem / (NW * wb + 1024));   // estimate; only used to size chunks
  if (per_sm < 1) per_sm = 1;
  return (long long)p->model->nsm * per_sm * NW * NS;
}

// Host-buffer variant: device staging lives in the problem's workspace region (separate allocations).
extern "C" int bik_step_host(const bik_problem* cp, int B, float* q_host, const bik_inputs* in, float dt, double damping, int nsteps, int integrate,
                             float* dq_host, int32_t* status_host, size_t* h2d_bytes, size_t* d2h_bytes) {
  int rc = check_inputs(cp, in, false);
  if (rc) return rc;
  if (!q_host || !dq_host || B < 0) return fail(BIK_ERR_INVALID, "bad argument");
  if (B == 0) return BIK_OK;
  bik_problem* p = const_cast<bik_problem*>(cp);
  const PHeader& h = p->h;
  DeviceGuard g(p->model->device);
  size_t nq = h.nq, nv = h.nv, b = (size_t)B;
  size_t pt_elems = (in->posture_batched ? b : 1) * (size_t)h.P * nq;
  {
    std::lock_guard<std::mutex> lock(p->mu);
    if (b > p->host_B || pt_elems > p->host_pt_elems) {
      cudaFree(p->hq); cudaFree(p->hft); cudaFree(p->hpt); cudaFree(p->hct); cudaFree(p->hdq); cudaFree(p->hst);
      p->hq = p->hft = p->hpt = p->hct = p->hdq = nullptr; p->hst = nullptr; p->host_B = 0;
      CUDA_OK(cudaMalloc(&p->hq, 4 * b * nq));
      CUDA_OK(cudaMalloc(&p->hft, 4 * b * (h.F > 0 ? h.F : 1) * 7));
      CUDA_OK(cudaMalloc(&p->hpt, 4 * (pt_elems > 0 ? pt_elems : 1)));
      CUDA_OK(cudaMalloc(&p->hct, 4 * b * (h.C > 0 ? h.C : 1) * 3));
      CUDA_OK(cudaMalloc(&p->hdq, 4 * b * nv));
      CUDA_OK(cudaMalloc(&p->hst, 4 * b));
      p->host_B = b; p->host_pt_elems = pt_elems;
    }
  }
  // The batch is cut into chunks that alternate between two streams: the copies of one chunk overlap the kernels of
  // the other (BIK_HOST_CHUNKS, default 4; chunk boundaries at multiples of a resident K2 wave so no wave runs half empty).
  std::lock_guard<std::mutex> lock(p->mu);
  rc = ensure_workspace(p, B);
  if (rc) return rc;
  if (!p->hs[0]) { CUDA_OK(cudaStreamCreateWithFlags(&p->hs[0], cudaStreamNonBlocking)); CUDA_OK(cudaStreamCreateWithFlags(&p->hs[1], cudaStreamNonBlocking)); CUDA_OK(cudaEventCreateWithFlags(&p->hev, cudaEventDisableTiming)); }
  int NC = env_int("BIK_HOST_CHUNKS", 4);
  {
    K2Args ka;
    memset(&ka, 0, sizeof ka);
    ka.dq = reinterpret_cast<float*>(1); ka.damping = damping;
    if (NC < 1 || use_fixed(p, ka)) NC = 1;   // the fixed-size path keeps one scratch per launch grid: one chunk in flight
  }
  size_t chunk = (b + NC - 1) / NC;
  const long long wave = k2_wave_instances(p);
  if (wave > 0 && NC > 1) chunk = (size_t)(((long long)chunk + wave - 1) / wave * wave);
  if (chunk < 1024) chunk = b;
  size_t up = 0, down = 0;
  if (h.P) { CUDA_OK(cudaMemcpyAsync(p->hpt, in->posture_targets, 4 * pt_elems, cudaMemcpyHostToDevice, p->hs[0])); up += 4 * pt_elems; }
  CUDA_OK(cudaEventRecord(p->hev, p->hs[0]));
  CUDA_OK(cudaStreamWaitEvent(p->hs[1], p->hev, 0));
  int ci = 0;
  for (size_t o = 0; o < b; o += chunk, ++ci) {
    const size_t n = b - o < chunk ? b - o : chunk;
    cudaStream_t st = p->hs[ci & 1];
    CUDA_OK(cudaMemcpyAsync(p->hq + o * nq, q_host + o * nq, 4 * n * nq, cudaMemcpyHostToDevice, st)); up += 4 * n * nq;
    if (h.F) { CUDA_OK(cudaMemcpyAsync(p->hft + o * h.F * 7, in->frame_targets + o * h.F * 7, 4 * n * h.F * 7, cudaMemcpyHostToDevice, st)); up += 4 * n * h.F * 7; }
    if (h.C) { CUDA_OK(cudaMemcpyAsync(p->hct + o * h.C * 3, in->com_targets + o * h.C * 3, 4 * n * h.C * 3, cudaMemcpyHostToDevice, st)); up += 4 * n * h.C * 3; }
    bik_inputs din = *in;
    din.q = p->hq + o * nq; din.frame_targets = p->hft + o * h.F * 7; din.com_targets = p->hct + o * h.C * 3;
    din.posture_targets = in->posture_batched ? p->hpt + o * (size_t)h.P * nq : p->hpt;
    rc = step_core(p, (int)n, o, p->hq + o * nq, &din, dt, damping, nsteps, integrate, p->hdq + o * nv, status_host ? p->hst + o : nullptr, st);
    if (rc) return rc;
    CUDA_OK(cudaMemcpyAsync(dq_host + o * nv, p->hdq + o * nv, 4 * n * nv, cudaMemcpyDeviceToHost, st)); down += 4 * n * nv;
    if (integrate) { CUDA_OK(cudaMemcpyAsync(q_host + o * nq, p->hq + o * nq, 4 * n * nq, cudaMemcpyDeviceToHost, st)); down += 4 * n * nq; }
    if (status_host) { CUDA_OK(cudaMemcpyAsync(status_host + o, p->hst + o, 4 * n, cudaMemcpyDeviceToHost, st)); down += 4 * n; }
  }
  CUDA_OK(cudaStreamSynchronize(p->hs[0]));
  CUDA_OK(cudaStreamSynchronize(p->hs[1]));
  if (h2d_bytes) *h2d_bytes = up;
  if (d2h_bytes) *d2h_bytes = down;
  return BIK_OK;
}
