// bik_k2t.cu -- small-group K2 path (bik_k2t.h): G lanes per problem, 32/G problems per warp; tables are read straight from
// the (L1-resident) global image so that all of shared memory goes to the per-problem triangles; warps are independent.
// Tiles are handed out by an atomic counter (sched[0]) when the launcher has one for this stream: the pivoting count
// differs between instances (G1: 91 % need one iteration, 1 in 10 000 needs six or more), and with a static round-robin
// the warp that meets a slow instance also keeps its whole share of ordinary tiles, which set the kernel time.  The last
// CTA to leave (sched[1] counts them) rewinds the counter for the next launch on the stream.
#include "bik_dev.cuh"

using namespace bik;

template <typename T, int G, int MAXT, typename M, bool F32IO>
__global__ void __launch_bounds__(MAXT, 512 / MAXT) k2t_kernel(const uint32_t* __restrict__ gimage, int warp_bytes, K2Args a, unsigned int* sched) {
  extern __shared__ __align__(16) uint32_t smem[];
  constexpr int NS = 32 / G;
  if (a.gate && *a.gate == 0) return;   // (uniform over the grid: the tile counter is left untouched)
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
      k2t_warp_tile<T, G, NS, M, F32IO>(P, a, (long long)t * NS, wsm, lane, St, uw);
    }
    __syncthreads();
    if (threadIdx.x == 0 && atomicInc(&sched[1], gridDim.x - 1) == gridDim.x - 1) { __threadfence(); sched[0] = 0u; }
    return;
  }
  for (long long tile = (long long)blockIdx.x * nwarps + warp; tile < ntiles; tile += (long long)gridDim.x * nwarps)
    k2t_warp_tile<T, G, NS, M, F32IO>(P, a, tile * NS, wsm, lane, St, uw);
}

template <typename T, int G, typename M, bool F32IO>
static int launch_k2t_io(const bik_problem* p, const K2Args& a, unsigned int* sched, cudaStream_t st) {
  constexpr int NS = 32 / G, MAXT = sizeof(T) == 8 ? 256 : 512;
  PView P{p->image.data()};
  const size_t wb = (size_t)k2t_warp_bytes(P, sizeof(T), NS);
  int NW = MAXT / 32;
  while (NW > 1 && NW * wb > (size_t)p->model->max_smem) --NW;
  const size_t smem = NW * wb;
  int grid = 1;
  long long tiles = ((long long)a.B + NS - 1) / NS;
  int rc = bik_launch_geometry((const void*)k2t_kernel<T, G, MAXT, M, F32IO>, p->model, smem, 32 * NW, (tiles + NW - 1) / NW, &grid);
  if (rc) return rc;
  k2t_kernel<T, G, MAXT, M, F32IO><<<grid, 32 * NW, smem, st>>>(p->d_image, (int)wb, a, sched);
  CUDA_OK(cudaGetLastError());
  return BIK_OK;
}
template <typename T, int G, typename M>
static int launch_k2t(const bik_problem* p, const K2Args& a, unsigned int* sched, cudaStream_t st) {
  if (sizeof(T) == 4 || !(a.io64 || a.pk64 || a.dense64)) return launch_k2t_io<T, G, M, true>(p, a, sched, st);   // fp32 buffers everywhere
  return launch_k2t_io<T, G, M, sizeof(T) == 4>(p, a, sched, st);
}
static int group_of(const bik_problem* p) { return (p->h.nu > K2T_NMAX || p->k2_group == 8) ? 8 : 4; }
static bool use_double(const bik_problem* p, const K2Args& a) { return p->solve_double || a.io64 || a.pk64 || a.dense64 || p->h.nu > K2T_NMAX; }

bool bik_k2_group_applies(const bik_problem* p) {
  const PHeader& h = p->h;
  if (p->k2_general || h.npairs != 0 || h.nu < 1 || h.nu > K2T_NMAX_WIDE) return false;
  PView P{p->image.data()};
  const int ts = (p->solve_double || h.nu > K2T_NMAX) ? 8 : 4;
  return k2t_warp_bytes(P, ts, 32 / group_of(p)) <= p->model->max_smem;
}
long long bik_k2_group_wave(const bik_problem* p) {
  if (!bik_k2_group_applies(p)) return 0;
  PView P{p->image.data()};
  const int G = group_of(p), NS = 32 / G, ts = (p->solve_double || p->h.nu > K2T_NMAX) ? 8 : 4, maxt = ts == 8 ? 256 : 512;
  const size_t wb = (size_t)k2t_warp_bytes(P, ts, NS);
  int NW = maxt / 32;
  while (NW > 1 && NW * wb > (size_t)p->model->max_smem) --NW;
  int per_sm = (int)((size_t)p->model->max_smem / (NW * wb + 1024));   // estimate; only used to size chunks
  if (per_sm < 1) per_sm = 1;
  return (long long)p->model->nsm * per_sm * NW * NS;
}
int bik_launch_k2_group(const bik_problem* p, const K2Args& a, unsigned int* sched, cudaStream_t st) {
  if (p->h.nu > K2T_NMAX) return launch_k2t<double, 8, uint64_t>(p, a, sched, st);   // wide: 64-bit active-set masks
  if (use_double(p, a)) return group_of(p) == 8 ? launch_k2t<double, 8, uint32_t>(p, a, sched, st) : launch_k2t<double, 4, uint32_t>(p, a, sched, st);
  return group_of(p) == 8 ? launch_k2t<float, 8, uint32_t>(p, a, sched, st) : launch_k2t<float, 4, uint32_t>(p, a, sched, st);
}
const char* bik_k2_describe(const bik_problem* p, char* buf, size_t cap) {
  if (bik_k2_group_applies(p)) snprintf(buf, cap, "small-group G=%d%s", group_of(p), p->h.nu > K2T_NMAX ? " (64-bit masks)" : "");
  else {
    const int rows = p->h.nu + 1, want = p->k2_lanes;
    const int lanes = ((want == 8 || want == 0) && rows <= 8) ? 8 : (((want == 16 || want == 0) && rows <= 16) ? 16 : 32);
    if (lanes == 32) snprintf(buf, cap, "general warp-per-problem");
    else snprintf(buf, cap, "general %d-lanes-per-problem", lanes);
  }
  return buf;
}
