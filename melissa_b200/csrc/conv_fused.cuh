// Fused GATv2 convolution: projection GEMM (tcgen05) + attention (SIMT) in one persistent kernel.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace mls {

struct FusedConvArgs {
  int rows;                  // node rows of this pass (n_graphs * N)
  int K;                     // input features (multiple of 64)
  int N, H, n_graphs;
  int G;                     // graphs per 128-row MMA tile (G * N <= 100)
  const float* proj_bias;    // [2*H*C]  x_l bias | x_r bias
  const float* att;          // [H*C]
  const float* conv_bias;    // [H*C]
  const float* obs;          // chunk base (decision-maker mask = obs col 7)
  long long obs_stride;
  int scale_rows;            // multiply the projected rows by the decision-maker mask (conv2: x1 * dm commutes with the GEMM)
  const unsigned short* csr_ptr;   // [graphs][N+1]
  const unsigned char* csr_src;    // [graphs][N*32]
  const int* slot;           // [rows] or NULL
  int ctrl_only;
  __nv_bfloat16* x_out;      // [rows, H*C] relu(conv) or NULL
  __nv_bfloat16* z;          // snapshot rows or NULL
  int ldz, z_col;
};

// X: [rows, K] bf16 (row stride K), Wt: [2*H*C, K] bf16.  Returns MLS_ERR_UNSUPPORTED when the shape does not fit.
int fused_gatv2_conv_launch(const __nv_bfloat16* X, const __nv_bfloat16* Wt, const FusedConvArgs& a, int sm_count, cudaStream_t st);

}  // namespace mls
