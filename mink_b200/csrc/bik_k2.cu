// bik_k2.cu -- general K2 path: one warp per problem (collision rows, or more than 64 coupled dofs); bik_k2.h.
#include "bik_dev.cuh"

using namespace bik;

template <typename T, int SLOTS>
__global__ void __launch_bounds__(256) k2_kernel(const uint32_t* __restrict__ gimage, int words, int use_tma, K2Args a) {
  extern __shared__ __align__(16) uint32_t smem[];
  __shared__ __align__(8) uint64_t bar;
  if (a.gate && *a.gate == 0) return;
  stage_image(smem, gimage, words, &bar, use_tma);
  PView P{smem};
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  char* wsm = reinterpret_cast<char*>(smem + words) + (size_t)warp * k2_warp_bytes(P.h(), sizeof(T));
  for (int b = blockIdx.x * nwarps + warp; b < a.B; b += gridDim.x * nwarps) k2_warp<T, 32, SLOTS>(P, a, b, wsm, lane);
}

template <typename T, int SLOTS>
static int launch_k2(const bik_problem* p, const K2Args& a, cudaStream_t st) {
  const PHeader& h = p->h;
  int NW = p->k2_warps;
  auto need = [&](int nw) { return (size_t)h.words32 * 4 + (size_t)nw * k2_warp_bytes(h, sizeof(T)); };
  while (NW > 1 && (int)need(NW) > p->model->max_smem) NW >>= 1;
  size_t smem = need(NW);
  int grid = 1;
  int rc = bik_launch_geometry((const void*)k2_kernel<T, SLOTS>, p->model, smem, 32 * NW, ((long long)a.B + NW - 1) / NW, &grid);
  if (rc) return rc;
  k2_kernel<T, SLOTS><<<grid, 32 * NW, smem, st>>>(p->d_image, h.words32, p->model->use_tma, a);
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
int bik_launch_k2_general(const bik_problem* p, const K2Args& a, cudaStream_t st) {
  const bool dbl = p->solve_double || a.io64 || a.pk64 || a.gc64 || a.dense64;
  return dbl ? dispatch_k2_slots<double>(p, a, st) : dispatch_k2_slots<float>(p, a, st);
}
