// Nearest-neighbour (Chamfer) distance between point sets -- the B200 replacement of the reference's custom TF op
// NnDistance (models/tf_ops/nn_distance/tf_nndistance.cpp:21-43 CPU `nnsearch`, tf_nndistance_g.cu:5-131 GPU kernel),
// used by the mesh metrics (test/test_cd_emd.py:300, test/test_f_score.py:253).  Brute force, HBM/L2-light:
// each CTA keeps 512 query points in registers and streams the other set through shared memory in 1024-point
// tiles.  Arithmetic follows the reference CPU kernel exactly -- float32 (dx*dx + dy*dy) + dz*dz without FMA
// contraction, first minimum wins -- so distances and indices are bit-identical to it.
#include "common.cuh"

namespace disn {
namespace {

constexpr int NN_THREADS = 256;
constexpr int NN_TILE = 1024;

__global__ void __launch_bounds__(NN_THREADS) nn_distance_kernel(const float* __restrict__ xyz1, int n,
                                                                 const float* __restrict__ xyz2, int m,
                                                                 float* __restrict__ dist, int* __restrict__ idx) {
  __shared__ float sx[NN_TILE], sy[NN_TILE], sz[NN_TILE];
  const int b = blockIdx.y;
  const float* p1 = xyz1 + (size_t)b * n * 3;
  const float* p2 = xyz2 + (size_t)b * m * 3;
  const int j = blockIdx.x * NN_THREADS + threadIdx.x;
  float x1 = 0.f, y1 = 0.f, z1 = 0.f;
  if (j < n) { x1 = p1[j * 3 + 0]; y1 = p1[j * 3 + 1]; z1 = p1[j * 3 + 2]; }
  float best = 0.f;
  int besti = 0;
  for (int k0 = 0; k0 < m; k0 += NN_TILE) {
    const int cnt = min(NN_TILE, m - k0);
    __syncthreads();
    for (int i = threadIdx.x; i < cnt; i += NN_THREADS) {
      sx[i] = p2[(size_t)(k0 + i) * 3 + 0]; sy[i] = p2[(size_t)(k0 + i) * 3 + 1]; sz[i] = p2[(size_t)(k0 + i) * 3 + 2];
    }
    __syncthreads();
    if (j < n) {
#pragma unroll 4
      for (int i = 0; i < cnt; ++i) {
        const float dx = __fsub_rn(sx[i], x1), dy = __fsub_rn(sy[i], y1), dz = __fsub_rn(sz[i], z1);
        const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        if ((k0 + i) == 0 || d < best) { best = d; besti = k0 + i; }
      }
    }
  }
  if (j < n) { dist[(size_t)b * n + j] = best; idx[(size_t)b * n + j] = besti; }
}

}  // namespace

int nn_distance(disn_ctx* c, const float* d_xyz1, int n, const float* d_xyz2, int m, int B, float* d_dist1,
                int* d_idx1, float* d_dist2, int* d_idx2) {
  dim3 g1((n + NN_THREADS - 1) / NN_THREADS, B), g2((m + NN_THREADS - 1) / NN_THREADS, B);
  nn_distance_kernel<<<g1, NN_THREADS, 0, c->stream>>>(d_xyz1, n, d_xyz2, m, d_dist1, d_idx1);
  nn_distance_kernel<<<g2, NN_THREADS, 0, c->stream>>>(d_xyz2, m, d_xyz1, n, d_dist2, d_idx2);
  c->launches += 2;
  DISN_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace disn
