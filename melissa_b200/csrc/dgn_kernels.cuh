// Device kernels shared by the fp32 and bf16 forward paths of the DGN Q-networks.
//
// Math follows the reference networks and the third-party ops they call:
//   graph_env/env/utils/networks/common.py:31-63  (obs split, radius_graph, ctrl index)
//   graph_env/env/utils/networks/{l_dgn.py:117-149, dgn_r.py:98-127, hl_dgn.py:97-117}
//   PyG GATv2Conv / TransformerConv / softmax / global_*_pool, torch_cluster radius,
//   tianshou MLP + DQNPolicy.forward / exploration_noise  (SURVEY.md Appendix B).
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

// bf16 tensor-core path (dgn_forward_bf16.cu)
size_t dgn_workspace_bytes_bf16(const MlsNetDesc* d, int n_graphs);
int dgn_forward_bf16(const MlsNetDesc* d, const MlsNetWeights* w, const MlsForwardArgs* a, void* stream);
int dgn_prepare_bf16(const MlsNetDesc* d, const MlsNetWeights* w, int flags, void* workspace, size_t workspace_bytes, void* stream);
size_t dgn_csr_cache_bytes(const MlsNetDesc* d, int n_pool_graphs);
int dgn_csr_cache_build(const MlsNetDesc* d, const float* pos_obs, int64_t obs_stride, int n_pool_graphs, void* cache, size_t cache_bytes,
                        void* stream);

namespace mls {

constexpr int kC = 128;                 // channels per head the edge kernels are written for
constexpr int kMaxNbr = 32;             // torch_cluster radius_graph max_num_neighbors
__device__ __forceinline__ float r2_threshold() { return (float)(0.2 * 0.2); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Neighbour set of target node i in graph with node rows at `g_obs` (N rows of 8 floats):
// radius_graph(pos, r=0.2, loop=False, max_num_neighbors=32) -- candidates scanned in index
// order, kept while fma(dy,dy,dx*dx) < r*r in fp32, scan stops after 33 hits (self
// included), self dropped.  Warp-cooperative; result words are warp-uniform.
template <int W>
__device__ __forceinline__ void radius_neighbours(const float* __restrict__ g_obs, int N, int i, int lane,
                                                  uint32_t (&nb)[W], const int stride = 8) {
  // `stride` floats between node rows: 8 for observation rows, 2 for a staged (x, y) array
  const float xi = g_obs[i * stride + 0], yi = g_obs[i * stride + 1];
  const float thr = r2_threshold();
  int total = 0;
#pragma unroll
  for (int w = 0; w < W; ++w) {
    const int j = w * 32 + lane;
    bool hit = false;
    if (j < N) {
      const float dx = g_obs[j * stride + 0] - xi, dy = g_obs[j * stride + 1] - yi;
      const float d2 = __fmaf_rn(dy, dy, __fmul_rn(dx, dx));
      hit = d2 < thr;
    }
    nb[w] = __ballot_sync(0xffffffffu, hit);
    total += __popc(nb[w]);
  }
  if (total > kMaxNbr + 1) {          // keep the first 33 hits in index order
    int keep = kMaxNbr + 1;
#pragma unroll
    for (int w = 0; w < W; ++w) {
      const int c = __popc(nb[w]);
      if (c <= keep) { keep -= c; }
      else {
        uint32_t m = nb[w], out = 0;
        for (int t = 0; t < keep; ++t) { out |= m & (0u - m); m &= m - 1; }
        nb[w] = out;
        keep = 0;
      }
    }
  }
#pragma unroll
  for (int w = 0; w < W; ++w) if (w == (i >> 5)) nb[w] &= ~(1u << (i & 31));
}

}  // namespace mls
