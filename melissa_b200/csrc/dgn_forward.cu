// DGN-R / L-DGN / HL-DGN forward + dueling head + (epsilon-)greedy action selection.
//
// fp32 ("exact") path: layer-by-layer kernels over chunks of graphs whose intermediates
// are sized to stay L2-resident and are reused chunk after chunk (so they never need to
// go to HBM).  One GNN body evaluation per graph serves every controlling agent of that
// graph (all agents of a round observe the same obs_matrix, reference graph.py:361-371).
//
// Reference: graph_env/env/utils/networks/{common.py:6-64, l_dgn.py:92-151,
// dgn_r.py:82-129, hl_dgn.py:82-119}; PyG GATv2Conv/TransformerConv/softmax/pool,
// torch_cluster radius_graph, tianshou MLP/DQNPolicy.
#include "dgn_kernels.cuh"
#include "gemm_tcgen05.cuh"

namespace mls {

// ------------------------------------------------------------------------------------------
// encoder layer 0: h = relu(W0 f + b0), f = obs cols 2..6 (common.py:40-44), K = input_dim (5)
__global__ void enc0_kernel(const float* __restrict__ obs, int64_t obs_stride, int N, int rows, int in_dim,
                            const float* __restrict__ w0, const float* __restrict__ b0, int hidden,
                            float* __restrict__ h) {
  const int r = blockIdx.x * blockDim.y + threadIdx.y;
  if (r >= rows) return;
  const int g = r / N, i = r - g * N;
  const float* f = obs + (int64_t)g * obs_stride + i * 8 + 2;
  for (int c = threadIdx.x; c < hidden; c += blockDim.x) {
    float acc = b0[c];
    for (int k = 0; k < in_dim; ++k) acc = fmaf(w0[c * in_dim + k], f[k], acc);
    h[(size_t)r * hidden + c] = fmaxf(acc, 0.0f);
  }
}

// ------------------------------------------------------------------------------------------
// C[M, Nout] = act( (A[M, K] * row_scale) @ Wt[Nout, K]^T + bias ), fp32 FMA.
// 64x64 tile, BK 16, 256 threads, 4x4 per thread.  M may come from device memory.
constexpr int GB = 64, GK = 16;
__global__ void __launch_bounds__(256) sgemm_kernel(const float* __restrict__ A, int lda,
                                                    const float* __restrict__ obs_for_scale, int64_t obs_stride, int N,
                                                    const float* __restrict__ Wt, int ldw, const float* __restrict__ bias,
                                                    float* __restrict__ C, int ldc, int M, const int* __restrict__ m_dev,
                                                    int Nout, int K, int relu) {
  if (m_dev) M = min(M, *m_dev);
  const int m0 = blockIdx.y * GB, n0 = blockIdx.x * GB;
  if (m0 >= M) return;
  __shared__ float As[GK][GB + 4];
  __shared__ float Bs[GK][GB + 4];
  const int tid = threadIdx.x;
  const int lr = tid >> 2, lk = (tid & 3) * 4;       // loader: row 0..63, k 0,4,8,12
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0f;
  const int ar = m0 + lr;
  float scale = 1.0f;
  if (obs_for_scale && ar < M) {       // decision-maker mask = obs col 7 (common.py:40-41; l_dgn.py:128)
    const int g = ar / N, i = ar - g * N;
    scale = obs_for_scale[(int64_t)g * obs_stride + i * 8 + 7];
  }
  const int br = n0 + lr;
  for (int k0 = 0; k0 < K; k0 += GK) {
    float4 av = make_float4(0.f, 0.f, 0.f, 0.f), bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ar < M) av = *reinterpret_cast<const float4*>(A + (size_t)ar * lda + k0 + lk);
    if (br < Nout) bv = *reinterpret_cast<const float4*>(Wt + (size_t)br * ldw + k0 + lk);
    As[lk + 0][lr] = av.x * scale; As[lk + 1][lr] = av.y * scale; As[lk + 2][lr] = av.z * scale; As[lk + 3][lr] = av.w * scale;
    Bs[lk + 0][lr] = bv.x; Bs[lk + 1][lr] = bv.y; Bs[lk + 2][lr] = bv.z; Bs[lk + 3][lr] = bv.w;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GK; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) acc[x][y] = fmaf(a[x], b[y], acc[x][y]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int x = 0; x < 4; ++x) {
    const int r = m0 + ty * 4 + x;
    if (r >= M) continue;
#pragma unroll
    for (int y = 0; y < 4; ++y) {
      const int c = n0 + tx * 4 + y;
      if (c >= Nout) continue;
      float v = acc[x][y] + (bias ? bias[c] : 0.0f);
      if (relu) v = fmaxf(v, 0.0f);
      C[(size_t)r * ldc + c] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------
// fp32-grade GEMM on the tensor cores.  Every fp32 operand value is split three ways into bf16,
//     a = hi + mid + lo,  hi = bf16(a), mid = bf16(a - hi), lo = bf16(a - hi - mid)      (|a - hi - mid - lo| <= 2^-26 |a|)
// and the six significant cross products hi.hi + hi.mid + mid.hi + hi.lo + lo.hi + mid.mid (dropped terms <= 2^-26) become
// ONE bf16 GEMM by concatenating the pieces along K:
//     A' = [hi | hi | mid | hi | lo | mid]   (M x 6K),   B' = [hi | mid | hi | lo | hi | mid]   (N x 6K)
// accumulated in fp32 in TMEM by gemm_bf16_tcgen05_kernel, which then writes fp32 (GemmEpilogue::Cf).
// which = 0: A' pattern, 1: B' pattern.  One thread per (row, 4 columns).
__global__ void __launch_bounds__(256) split3_kernel(const float* __restrict__ X, int ld, long long rows, int K, int which,
                                                     __nv_bfloat16* __restrict__ out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int kq = K >> 2;
  if (t >= rows * kq) return;
  const long long r = t / kq;
  const int k = (int)(t - r * kq) * 4;
  const float4 v = *reinterpret_cast<const float4*>(X + r * ld + k);
  const float a[4] = {v.x, v.y, v.z, v.w};
  __nv_bfloat16 hi[4], mid[4], lo[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    hi[i] = __float2bfloat16_rn(a[i]);
    const float r1 = a[i] - __bfloat162float(hi[i]);
    mid[i] = __float2bfloat16_rn(r1);
    lo[i] = __float2bfloat16_rn(r1 - __bfloat162float(mid[i]));
  }
  auto pack = [](const __nv_bfloat16* p) {
    uint2 u;
    u.x = (uint32_t)__bfloat16_as_ushort(p[0]) | ((uint32_t)__bfloat16_as_ushort(p[1]) << 16);
    u.y = (uint32_t)__bfloat16_as_ushort(p[2]) | ((uint32_t)__bfloat16_as_ushort(p[3]) << 16);
    return u;
  };
  const uint2 H = pack(hi), Mi = pack(mid), L = pack(lo);
  __nv_bfloat16* o = out + r * (6ll * K) + k;
  const uint2 seqA[6] = {H, H, Mi, H, L, Mi}, seqB[6] = {H, Mi, H, L, H, Mi};
#pragma unroll
  for (int p = 0; p < 6; ++p) *reinterpret_cast<uint2*>(o + (long long)p * K) = which ? seqB[p] : seqA[p];
}

// ------------------------------------------------------------------------------------------
// GATv2Conv edge phase + bias + ReLU for one (node, head) per warp.
//   P row = [x_l (H*C) | x_r (H*C)];  e_ij = sum_c att[h,c] * leaky_relu(x_l[j] + x_r[i], 0.2)
//   alpha = softmax over {neighbours j} U {i}: exp(e - max) / (sum + 1e-16);  out = sum alpha x_l[j] + bias
template <int W>
__global__ void __launch_bounds__(256) gatv2_edge_kernel(const float* __restrict__ P, int ldp,
                                                         const float* __restrict__ obs, int64_t obs_stride, int N,
                                                         int rows, int H, const float* __restrict__ att,
                                                         const float* __restrict__ bias, float* __restrict__ out, int ldo) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp >= rows * H) return;
  const int r = warp / H, h = warp - r * H;
  const int g = r / N, i = r - g * N;
  const float* g_obs = obs + (int64_t)g * obs_stride;
  uint32_t nb[W];
  radius_neighbours<W>(g_obs, N, i, lane, nb);
#pragma unroll
  for (int w = 0; w < W; ++w) if (w == (i >> 5)) nb[w] |= 1u << (i & 31);      // add_self_loops
  const int HC = H * kC;
  const size_t base = (size_t)g * N;
  const float4 xr = *reinterpret_cast<const float4*>(P + (base + i) * ldp + HC + h * kC + lane * 4);
  const float4 a4 = *reinterpret_cast<const float4*>(att + h * kC + lane * 4);
  // pass 1: logits + max
  float e_loc[W];            // lane l keeps the logit of candidate w*32+l
  float mx = -INFINITY;
#pragma unroll
  for (int w = 0; w < W; ++w) {
    e_loc[w] = -INFINITY;
    uint32_t bits = nb[w];
    while (bits) {
      const int jl = __ffs(bits) - 1;
      bits &= bits - 1;
      const int j = w * 32 + jl;
      const float4 xl = *reinterpret_cast<const float4*>(P + (base + j) * ldp + h * kC + lane * 4);
      float s0 = xl.x + xr.x, s1 = xl.y + xr.y, s2 = xl.z + xr.z, s3 = xl.w + xr.w;
      s0 = s0 > 0.f ? s0 : 0.2f * s0; s1 = s1 > 0.f ? s1 : 0.2f * s1;
      s2 = s2 > 0.f ? s2 : 0.2f * s2; s3 = s3 > 0.f ? s3 : 0.2f * s3;
      float e = warp_sum(s0 * a4.x + s1 * a4.y + s2 * a4.z + s3 * a4.w);
      if (lane == jl) e_loc[w] = e;
      mx = fmaxf(mx, e);
    }
  }
  // pass 2: softmax weights + aggregation
  float den = 0.f;
#pragma unroll
  for (int w = 0; w < W; ++w) {
    float ex = (nb[w] >> lane) & 1u ? expf(e_loc[w] - mx) : 0.f;
    e_loc[w] = ex;
    den += warp_sum(ex);
  }
  den += 1e-16f;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int w = 0; w < W; ++w) {
    uint32_t bits = nb[w];
    while (bits) {
      const int jl = __ffs(bits) - 1;
      bits &= bits - 1;
      const int j = w * 32 + jl;
      const float a = __shfl_sync(0xffffffffu, e_loc[w], jl) / den;
      const float4 xl = *reinterpret_cast<const float4*>(P + (base + j) * ldp + h * kC + lane * 4);
      acc.x = fmaf(a, xl.x, acc.x); acc.y = fmaf(a, xl.y, acc.y); acc.z = fmaf(a, xl.z, acc.z); acc.w = fmaf(a, xl.w, acc.w);
    }
  }
  const float4 b4 = *reinterpret_cast<const float4*>(bias + h * kC + lane * 4);
  float4 o;
  o.x = fmaxf(acc.x + b4.x, 0.f); o.y = fmaxf(acc.y + b4.y, 0.f); o.z = fmaxf(acc.z + b4.z, 0.f); o.w = fmaxf(acc.w + b4.w, 0.f);
  *reinterpret_cast<float4*>(out + (base + i) * ldo + h * kC + lane * 4) = o;
}

// TransformerConv(root_weight=False) edge phase + ReLU.  P row = [q | k | v] (each H*C).
//   e_ij = <q_i, k_j>_h / sqrt(C); softmax over neighbours (no self loop); out = sum alpha v_j; isolated -> 0
template <int W>
__global__ void __launch_bounds__(256) transformer_edge_kernel(const float* __restrict__ P, int ldp,
                                                               const float* __restrict__ obs, int64_t obs_stride, int N,
                                                               int rows, int H, float* __restrict__ out, int ldo) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp >= rows * H) return;
  const int r = warp / H, h = warp - r * H;
  const int g = r / N, i = r - g * N;
  const float* g_obs = obs + (int64_t)g * obs_stride;
  uint32_t nb[W];
  radius_neighbours<W>(g_obs, N, i, lane, nb);
  const int HC = H * kC;
  const size_t base = (size_t)g * N;
  const float4 q = *reinterpret_cast<const float4*>(P + (base + i) * ldp + h * kC + lane * 4);
  const float inv_sqrt_c = 1.0f / sqrtf((float)kC);
  float e_loc[W];
  float mx = -INFINITY;
#pragma unroll
  for (int w = 0; w < W; ++w) {
    e_loc[w] = -INFINITY;
    uint32_t bits = nb[w];
    while (bits) {
      const int jl = __ffs(bits) - 1;
      bits &= bits - 1;
      const int j = w * 32 + jl;
      const float4 kk = *reinterpret_cast<const float4*>(P + (base + j) * ldp + HC + h * kC + lane * 4);
      float e = warp_sum(q.x * kk.x + q.y * kk.y + q.z * kk.z + q.w * kk.w) * inv_sqrt_c;
      if (lane == jl) e_loc[w] = e;
      mx = fmaxf(mx, e);
    }
  }
  float den = 0.f;
#pragma unroll
  for (int w = 0; w < W; ++w) {
    float ex = (nb[w] >> lane) & 1u ? expf(e_loc[w] - mx) : 0.f;
    e_loc[w] = ex;
    den += warp_sum(ex);
  }
  den += 1e-16f;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int w = 0; w < W; ++w) {
    uint32_t bits = nb[w];
    while (bits) {
      const int jl = __ffs(bits) - 1;
      bits &= bits - 1;
      const int j = w * 32 + jl;
      const float a = __shfl_sync(0xffffffffu, e_loc[w], jl) / den;
      const float4 v = *reinterpret_cast<const float4*>(P + (base + j) * ldp + 2 * HC + h * kC + lane * 4);
      acc.x = fmaf(a, v.x, acc.x); acc.y = fmaf(a, v.y, acc.y); acc.z = fmaf(a, v.z, acc.z); acc.w = fmaf(a, v.w, acc.w);
    }
  }
  float4 o;
  o.x = fmaxf(acc.x, 0.f); o.y = fmaxf(acc.y, 0.f); o.z = fmaxf(acc.z, 0.f); o.w = fmaxf(acc.w, 0.f);
  *reinterpret_cast<float4*>(out + (base + i) * ldo + h * kC + lane * 4) = o;
}

// ------------------------------------------------------------------------------------------
// HL-DGN: z[g] = pool_i( x1[g,i,:] * dm[g,i] )  (hl_dgn.py:103-108)
__global__ void pool_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ obs, int64_t obs_stride,
                            int N, int HC, int mode, float* __restrict__ z) {
  const int g = blockIdx.x;
  const float* g_obs = obs + (int64_t)g * obs_stride;
  for (int c = threadIdx.x; c < HC; c += blockDim.x) {
    float acc = mode == MLS_POOL_MAX ? -INFINITY : 0.f;
    for (int i = 0; i < N; ++i) {
      const float v = x[((size_t)g * N + i) * ldx + c] * g_obs[i * 8 + 7];
      acc = mode == MLS_POOL_MAX ? fmaxf(acc, v) : acc + v;
    }
    if (mode == MLS_POOL_MEAN) acc = acc / (float)N;
    z[(size_t)g * HC + c] = acc;
  }
}

// ------------------------------------------------------------------------------------------
// controlling-node list of a chunk.  mode 0: every set entry of ctrl_mask; mode 1: one per
// graph, clamp(obs[g][8N], 0, N-1) (common.py:63).
__global__ void ctrl_list_kernel(const uint8_t* __restrict__ ctrl_mask, const float* __restrict__ obs, int64_t obs_stride,
                                 int N, int n_graphs, int mode, int* __restrict__ idx, int* __restrict__ count) {
  const int lane = threadIdx.x & 31;
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (g >= n_graphs) return;
  if (mode == 1) {
    if (lane == 0) {
      float c = obs[(int64_t)g * obs_stride + (int64_t)N * 8];
      c = fminf(fmaxf(c, 0.f), (float)(N - 1));
      idx[g] = g * N + (int)(long long)c;
      if (g == 0) *count = n_graphs;
    }
    return;
  }
  for (int i0 = 0; i0 < N; i0 += 32) {
    const int i = i0 + lane;
    const bool c = i < N && ctrl_mask[(size_t)g * N + i] != 0;
    const uint32_t bal = __ballot_sync(0xffffffffu, c);
    const int n = __popc(bal);
    int slot = 0;
    if (lane == 0 && n) slot = atomicAdd(count, n);
    slot = __shfl_sync(0xffffffffu, slot, 0);
    if (c) idx[slot + __popc(bal & ((1u << lane) - 1))] = g * N + i;
  }
}

// z[t] = [x0[idx] | x1[idx] | x2[idx]]  (l_dgn.py:121-139): x1 snapshot is PRE-mask
__global__ void gather_kernel(const int* __restrict__ idx, const int* __restrict__ count, const float* __restrict__ x0,
                              int d0, const float* __restrict__ x1, int d1, const float* __restrict__ x2, int d2,
                              float* __restrict__ z) {
  const int t = blockIdx.x;
  if (t >= *count) return;
  const int r = idx[t];
  const int D = d0 + d1 + d2;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float v;
    if (c < d0) v = x0[(size_t)r * d0 + c];
    else if (c < d0 + d1) v = x1[(size_t)r * d1 + (c - d0)];
    else v = x2[(size_t)r * d2 + (c - d0 - d1)];
    z[(size_t)t * D + c] = v;
  }
}

struct ActArgs {
  float eps;
  uint64_t seed, offset;
  const double* rand3;
  const unsigned long long* offset_dev;   // optional device-side addend (round counter under CUDA-graph replay)
  uint64_t row0;                          // added to the row index that keys the Philox draw (sub-batch calls)
};

__device__ __forceinline__ int select_action(float q0, float q1, const ActArgs& a, uint64_t row) {
  // tianshou DQNPolicy.forward: argmax (first max wins); mask [1,1] leaves the logits unchanged
  int act = q1 > q0 ? 1 : 0;
  // exploration_noise: skipped when eps ~ 0 (np.isclose(eps, 0.0): |eps| <= 1e-8)
  if (fabsf(a.eps) > 1e-8f) {
    double ue, u0, u1;
    if (a.rand3) { ue = a.rand3[row * 3 + 0]; u0 = a.rand3[row * 3 + 1]; u1 = a.rand3[row * 3 + 2]; }
    else {
      Philox4 r = philox4x32_10(a.seed, row + a.row0, a.offset + (a.offset_dev ? *a.offset_dev : 0ull));
      ue = u01_from_u32x2(r.v[0], r.v[1]);
      u0 = (double)r.v[2] * (1.0 / 4294967296.0);
      u1 = (double)r.v[3] * (1.0 / 4294967296.0);
    }
    if (ue < (double)a.eps) act = (u1 + 1.0) > (u0 + 1.0) ? 1 : 0;   // argmax(rand(2) + mask)
  }
  return act;
}

// last layers of the dueling head: q = Wq2 hq + bq2 (2), v = Wv2 hv + bv2 (1);
// out = q - mean(q) + v; act = (eps-)greedy.  One warp per controlling row.
// hid row = [Q hidden (hh) | V hidden (hh)].
__global__ void head_out_kernel(const float* __restrict__ hid, int hh, const int* __restrict__ idx,
                                const int* __restrict__ count, int max_rows, const float* __restrict__ wq,
                                const float* __restrict__ bq, const float* __restrict__ wv, const float* __restrict__ bv,
                                int64_t row0, int per_graph_N, float* __restrict__ q_out, int8_t* __restrict__ act_out,
                                int out_mode, ActArgs aa) {
  const int lane = threadIdx.x & 31;
  const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int n = count ? min(*count, max_rows) : max_rows;
  if (t >= n) return;
  const float* hq = hid + (size_t)t * 2 * hh;
  const float* hv = hq + hh;
  float s0 = 0.f, s1 = 0.f, sv = 0.f;
  for (int c = lane; c < hh; c += 32) {
    s0 = fmaf(hq[c], wq[c], s0);
    s1 = fmaf(hq[c], wq[hh + c], s1);
    sv = fmaf(hv[c], wv[c], sv);
  }
  s0 = warp_sum(s0) + bq[0]; s1 = warp_sum(s1) + bq[1]; sv = warp_sum(sv) + bv[0];
  if (lane == 0) {
    const float mean = (s0 + s1) / 2.0f;
    const float o0 = (s0 - mean) + sv, o1 = (s1 - mean) + sv;
    // out_mode 0: q[global node row][2]; 1: q[global graph][2]; 2: per-graph scratch q (HL-DGN, scattered later)
    int64_t orow;
    if (out_mode == 0) orow = row0 + idx[t];
    else if (out_mode == 1) orow = row0 / per_graph_N + idx[t] / per_graph_N;
    else orow = t;
    q_out[orow * 2 + 0] = o0;
    q_out[orow * 2 + 1] = o1;
    if (act_out && out_mode != 2) act_out[orow] = (int8_t)select_action(o0, o1, aa, (uint64_t)orow);
  }
}

// Network without dueling heads: q = out_linear(latent) (l_dgn.py:88-90,149).  One warp per row of z.
__global__ void head_linear_kernel(const float* __restrict__ z, int latent, const int* __restrict__ idx, const int* __restrict__ count,
                                   int max_rows, const float* __restrict__ w, const float* __restrict__ b, int64_t row0, int per_graph_N,
                                   float* __restrict__ q_out, int8_t* __restrict__ act_out, int out_mode, ActArgs aa) {
  const int lane = threadIdx.x & 31;
  const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int n = count ? min(*count, max_rows) : max_rows;
  if (t >= n) return;
  const float* zr = z + (size_t)t * latent;
  float s0 = 0.f, s1 = 0.f;
  for (int c = lane; c < latent; c += 32) {
    s0 = fmaf(zr[c], w[c], s0);
    s1 = fmaf(zr[c], w[latent + c], s1);
  }
  s0 = warp_sum(s0) + b[0]; s1 = warp_sum(s1) + b[1];
  if (lane == 0) {
    int64_t orow;
    if (out_mode == 0) orow = row0 + idx[t];
    else if (out_mode == 1) orow = row0 / per_graph_N + idx[t] / per_graph_N;
    else orow = t;
    q_out[orow * 2 + 0] = s0;
    q_out[orow * 2 + 1] = s1;
    if (act_out && out_mode != 2) act_out[orow] = (int8_t)select_action(s0, s1, aa, (uint64_t)orow);
  }
}

// HL-DGN: the Q-values of a graph are shared by all its controlling agents (hl_dgn.py:108).
__global__ void hl_scatter_kernel(const float* __restrict__ qg, const uint8_t* __restrict__ ctrl_mask, int N, int n_graphs,
                                  int64_t graph0, int mode, float* __restrict__ q_out, int8_t* __restrict__ act_out,
                                  ActArgs aa) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (mode == 1) {
    if (t >= n_graphs) return;
    const float o0 = qg[t * 2], o1 = qg[t * 2 + 1];
    const int64_t orow = graph0 + t;
    q_out[orow * 2] = o0; q_out[orow * 2 + 1] = o1;
    if (act_out) act_out[orow] = (int8_t)select_action(o0, o1, aa, (uint64_t)orow);
    return;
  }
  if (t >= (int64_t)n_graphs * N) return;
  const int64_t g = t / N;
  const int64_t orow = graph0 * N + t;
  if (!ctrl_mask[t]) return;
  const float o0 = qg[g * 2], o1 = qg[g * 2 + 1];
  q_out[orow * 2] = o0; q_out[orow * 2 + 1] = o1;
  if (act_out) act_out[orow] = (int8_t)select_action(o0, o1, aa, (uint64_t)orow);
}

}  // namespace mls

// ================================================================================== host
namespace {
using namespace mls;

struct Ws {      // fp32 workspace carve-up for a chunk of Gc graphs (R = Gc*N rows)
  float *h, *x0, *P, *x1, *x2, *z, *hid1, *hid2, *qg;
  int *idx, *count;
  // tensor-core route (option fp32_tc): 3-way bf16 splits of the weights (B' operands) and of the current activation (A')
  __nv_bfloat16 *asplit, *w_enc1, *w_c1[3], *w_c2[3], *w_q0, *w_v0, *w_q1, *w_v1;
};

bool fp32_tc_enabled() { return mls_get_option("fp32_tc") != 0; }

size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

int chunk_graphs(const MlsNetDesc* d, int n_graphs) {
  // tensor-core route: 32768-row chunks (256 GEMM tiles of 128 rows); SIMT route: L2-sized chunks
  int gc = (fp32_tc_enabled() ? 32768 : 8192) / d->n_nodes;
  if (gc < 1) gc = 1;
  return n_graphs < gc ? n_graphs : gc;
}

size_t carve(const MlsNetDesc* d, int Gc, unsigned char* base, Ws* ws) {
  const size_t R = (size_t)Gc * d->n_nodes;
  const int hid = d->hidden, HC = d->hidden * d->heads;
  const int nproj = d->kind == MLS_NET_DGN_R ? 3 : 2;
  const int latent = d->kind == MLS_NET_HL_DGN ? HC : hid + 2 * HC;
  const size_t head_rows = d->kind == MLS_NET_HL_DGN ? (size_t)Gc : R;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes); return o; };
  size_t o_h = take(R * hid * 4), o_x0 = take(R * hid * 4), o_P = take(R * nproj * HC * 4), o_x1 = take(R * HC * 4);
  size_t o_x2 = take(d->kind == MLS_NET_HL_DGN ? 0 : R * HC * 4);
  size_t o_z = take(head_rows * latent * 4), o_h1 = take(head_rows * 2 * d->head_hidden * 4);
  size_t o_h2 = take(head_rows * 2 * d->head_hidden * 4), o_qg = take((size_t)Gc * 2 * 4);
  size_t o_idx = take(R * 4), o_cnt = take(4);
  const bool tc = fp32_tc_enabled();
  const int hh = d->head_hidden;
  const size_t kmax = (size_t)(latent > HC ? latent : HC);
  size_t o_as = take(tc ? R * 6 * kmax * 2 : 0), o_we = take(tc ? (size_t)hid * 6 * hid * 2 : 0);
  size_t o_wc1[3], o_wc2[3];
  for (int t = 0; t < 3; ++t) o_wc1[t] = take(tc && t < nproj ? (size_t)HC * 6 * hid * 2 : 0);
  for (int t = 0; t < 3; ++t) o_wc2[t] = take(tc && t < nproj && d->kind != MLS_NET_HL_DGN ? (size_t)HC * 6 * HC * 2 : 0);
  size_t o_wq0 = take(tc ? (size_t)hh * 6 * latent * 2 : 0), o_wv0 = take(tc ? (size_t)hh * 6 * latent * 2 : 0);
  size_t o_wq1 = take(tc ? (size_t)hh * 6 * hh * 2 : 0), o_wv1 = take(tc ? (size_t)hh * 6 * hh * 2 : 0);
  if (ws) {
    auto B16 = [&](size_t o) { return reinterpret_cast<__nv_bfloat16*>(base + o); };
    ws->asplit = B16(o_as); ws->w_enc1 = B16(o_we);
    for (int t = 0; t < 3; ++t) { ws->w_c1[t] = B16(o_wc1[t]); ws->w_c2[t] = B16(o_wc2[t]); }
    ws->w_q0 = B16(o_wq0); ws->w_v0 = B16(o_wv0); ws->w_q1 = B16(o_wq1); ws->w_v1 = B16(o_wv1);
    ws->h = (float*)(base + o_h); ws->x0 = (float*)(base + o_x0); ws->P = (float*)(base + o_P);
    ws->x1 = (float*)(base + o_x1); ws->x2 = (float*)(base + o_x2); ws->z = (float*)(base + o_z);
    ws->hid1 = (float*)(base + o_h1); ws->hid2 = (float*)(base + o_h2); ws->qg = (float*)(base + o_qg);
    ws->idx = (int*)(base + o_idx); ws->count = (int*)(base + o_cnt);
  }
  return off;
}

void sgemm(cudaStream_t st, const float* A, int lda, const float* obs_scale, int64_t obs_stride, int N, const float* Wt,
           int ldw, const float* bias, float* C, int ldc, int M, const int* m_dev, int Nout, int K, int relu) {
  dim3 grid((Nout + GB - 1) / GB, (M + GB - 1) / GB);
  sgemm_kernel<<<grid, 256, 0, st>>>(A, lda, obs_scale, obs_stride, N, Wt, ldw, bias, C, ldc, M, m_dev, Nout, K, relu);
  mls_count_launch();
}

void split3(cudaStream_t st, const float* X, int ld, long long rows, int K, int which, __nv_bfloat16* out) {
  const long long n = rows * (K >> 2);
  split3_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(X, ld, rows, K, which, out);
  mls_count_launch();
}

// Same contract as sgemm, on the tensor cores: A is split into `scratch` (M x 6K bf16), Wsplit is the pre-split weight.
int tcgemm(cudaStream_t st, int sms, const float* A, int lda, const float* obs_scale, int64_t obs_stride, int N,
           const __nv_bfloat16* Wsplit, const float* bias, float* C, int ldc, int M, const int* m_dev, int Nout, int K, int relu,
           __nv_bfloat16* scratch) {
  split3(st, A, lda, M, K, 0, scratch);
  GemmEpilogue e{};
  e.C = nullptr; e.ldc = 0; e.bias = bias; e.obs = obs_scale; e.obs_stride = obs_stride; e.nodes = N; e.relu = relu;
  e.Cf = C; e.ldcf = ldc;
  return gemm_bf16_launch(scratch, 6 * K, Wsplit, 6 * K, GemmShape{M, Nout, 6 * K, m_dev}, e, sms, st);
}

template <int W>
void launch_gat_edge(cudaStream_t st, const float* P, int ldp, const float* obs, int64_t os, int N, int rows, int H,
                     const float* att, const float* bias, float* out, int ldo) {
  const long long warps = (long long)rows * H;
  gatv2_edge_kernel<W><<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(P, ldp, obs, os, N, rows, H, att, bias, out, ldo);
  mls_count_launch();
}
template <int W>
void launch_tr_edge(cudaStream_t st, const float* P, int ldp, const float* obs, int64_t os, int N, int rows, int H,
                    float* out, int ldo) {
  const long long warps = (long long)rows * H;
  transformer_edge_kernel<W><<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(P, ldp, obs, os, N, rows, H, out, ldo);
  mls_count_launch();
}

}  // namespace

int bf16_chunk_graphs(const MlsNetDesc* d, int n_graphs);
extern "C" int mls_dgn_chunk_graphs(const MlsNetDesc* desc, int32_t n_graphs) {
  if (!desc || n_graphs <= 0 || desc->n_nodes <= 0) return 0;
  return desc->precision == MLS_PREC_BF16 ? bf16_chunk_graphs(desc, n_graphs) : chunk_graphs(desc, n_graphs);
}

extern "C" size_t mls_dgn_workspace_bytes(const MlsNetDesc* desc, int32_t n_graphs) {
  if (!desc || n_graphs <= 0 || desc->n_nodes <= 0) return 0;
  if (desc->precision == MLS_PREC_BF16) return dgn_workspace_bytes_bf16(desc, n_graphs);
  return carve(desc, chunk_graphs(desc, n_graphs), nullptr, nullptr);
}

static int check_desc(const MlsNetDesc* d) {
  MLS_CHECK_ARG(d->kind >= MLS_NET_DGN_R && d->kind <= MLS_NET_HL_DGN, "unknown network kind %d", d->kind);
  MLS_CHECK_ARG(d->n_nodes >= 1 && d->n_nodes <= MLS_MAX_NODES, "n_nodes out of range: %d", d->n_nodes);
  MLS_CHECK_ARG(d->hidden == kC, "hidden_dim must be %d (got %d)", kC, d->hidden);
  MLS_CHECK_ARG(d->heads >= 1 && d->heads <= 8, "num_heads must be in [1,8]");
  MLS_CHECK_ARG(d->input_dim >= 1 && d->input_dim <= 5, "input_dim must be <= 5 (obs rows hold 5 features)");
  MLS_CHECK_ARG(d->head_hidden % 16 == 0 && d->head_hidden >= 16, "head hidden size must be a multiple of 16");
  return MLS_OK;
}

extern "C" int mls_dgn_prepare(const MlsNetDesc* d, const MlsNetWeights* w, int32_t flags, void* workspace, size_t workspace_bytes,
                               void* stream) {
  MLS_CHECK_ARG(d && w, "NULL argument");
  if (int rc = check_desc(d)) return rc;
  if (d->precision != MLS_PREC_BF16) return MLS_OK;        // the fp32 path reads the parameters in place
  return dgn_prepare_bf16(d, w, flags, workspace, workspace_bytes, stream);
}

extern "C" size_t mls_dgn_csr_cache_bytes(const MlsNetDesc* d, int32_t n_pool_graphs) {
  if (!d || n_pool_graphs <= 0 || d->n_nodes <= 0 || d->n_nodes > MLS_MAX_NODES) return 0;
  return dgn_csr_cache_bytes(d, n_pool_graphs);
}

extern "C" int mls_dgn_csr_cache_build(const MlsNetDesc* d, const float* pos_obs, int64_t obs_stride, int32_t n_pool_graphs, void* cache,
                                       size_t cache_bytes, void* stream) {
  MLS_CHECK_ARG(d && pos_obs && cache && n_pool_graphs > 0, "NULL argument");
  MLS_CHECK_ARG(d->n_nodes >= 1 && d->n_nodes <= MLS_MAX_NODES, "n_nodes out of range: %d", d->n_nodes);
  MLS_CHECK_ARG(obs_stride >= (int64_t)d->n_nodes * 8 - 6, "position rows must be 8 floats apart");
  return dgn_csr_cache_build(d, pos_obs, obs_stride, n_pool_graphs, cache, cache_bytes, stream);
}

extern "C" int mls_dgn_forward(const MlsNetDesc* d, const MlsNetWeights* w, const MlsForwardArgs* a, void* stream) {
  MLS_CHECK_ARG(d && w && a, "NULL argument");
  MLS_CHECK_ARG(d->kind >= MLS_NET_DGN_R && d->kind <= MLS_NET_HL_DGN, "unknown network kind %d", d->kind);
  MLS_CHECK_ARG(d->n_nodes >= 1 && d->n_nodes <= MLS_MAX_NODES, "n_nodes out of range: %d", d->n_nodes);
  MLS_CHECK_ARG(d->hidden == kC, "hidden_dim must be %d (got %d)", kC, d->hidden);
  MLS_CHECK_ARG(d->heads >= 1 && d->heads <= 8, "num_heads must be in [1,8]");
  MLS_CHECK_ARG(d->input_dim >= 1 && d->input_dim <= 5, "input_dim must be <= 5 (obs rows hold 5 features)");
  MLS_CHECK_ARG(d->head_hidden % 16 == 0 && d->head_hidden >= 16, "head hidden size must be a multiple of 16");
  MLS_CHECK_ARG(a->obs && a->q && a->n_graphs >= 0, "obs/q missing");
  MLS_CHECK_ARG(a->ctrl_mode == 1 || a->ctrl_mask, "ctrl_mask missing");
  MLS_CHECK_ARG(a->obs_stride >= (int64_t)d->n_nodes * 8 + (a->ctrl_mode == 1 ? 1 : 0),
                "Expected %d feature cols for nodes, got %lld", d->n_nodes * 8, (long long)a->obs_stride - 1);
  if (a->n_graphs == 0) return MLS_OK;
  if (d->precision == MLS_PREC_BF16) return dgn_forward_bf16(d, w, a, stream);
  if (d->precision != MLS_PREC_FP32) {
    mls_set_error("unknown precision %d", d->precision);
    return MLS_ERR_UNSUPPORTED;
  }
  const int N = d->n_nodes, hid = d->hidden, H = d->heads, HC = hid * H, hh = d->head_hidden;
  const int Gc = chunk_graphs(d, a->n_graphs);
  MLS_CHECK_ARG(a->workspace && a->workspace_bytes >= carve(d, Gc, nullptr, nullptr), "workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  Ws ws;
  carve(d, Gc, (unsigned char*)a->workspace, &ws);
  const int Wn = mls_words_per_row(N);
  const bool hl = d->kind == MLS_NET_HL_DGN, tr = d->kind == MLS_NET_DGN_R;
  const int nproj = tr ? 3 : 2;
  const int latent = hl ? HC : hid + 2 * HC;
  ActArgs aa{a->eps, a->philox_seed, a->philox_offset, a->rand3, reinterpret_cast<const unsigned long long*>(a->philox_offset_dev), a->philox_row0};
  if (a->ctrl_mode == 0) {
    MLS_CUDA(cudaMemsetAsync(a->q, 0, (size_t)a->n_graphs * N * 2 * sizeof(float), st));
    if (a->act) MLS_CUDA(cudaMemsetAsync(a->act, 0xFF, (size_t)a->n_graphs * N, st));
  }

  cudaEvent_t ev0 = reinterpret_cast<cudaEvent_t>(a->prof_start), ev1 = reinterpret_cast<cudaEvent_t>(a->prof_stop);
  bool first_chunk = true;
  auto prof_begin = [&](int which) { if (first_chunk && ev0 && ev1 && a->prof_kernel == which) cudaEventRecord(ev0, st); };
  auto prof_end = [&](int which) { if (first_chunk && ev0 && ev1 && a->prof_kernel == which) cudaEventRecord(ev1, st); };
  // dense layers: tensor cores through the 3-way bf16 split (fp32-grade), or the SIMT sgemm (option fp32_tc = 0)
  const bool tc = fp32_tc_enabled() && hid % 64 == 0 && hh % 128 == 0;
  int sms = 0, rc_g = MLS_OK;
  if (tc) {
    int dev = 0;
    MLS_CUDA(cudaGetDevice(&dev));
    MLS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    split3(st, w->enc_w1, hid, hid, hid, 1, ws.w_enc1);
    const float* c1w[3] = {w->c1_wa, w->c1_wb, w->c1_wc};
    const float* c2w[3] = {w->c2_wa, w->c2_wb, w->c2_wc};
    for (int t = 0; t < nproj; ++t) {
      split3(st, c1w[t], hid, HC, hid, 1, ws.w_c1[t]);
      if (!hl) split3(st, c2w[t], HC, HC, HC, 1, ws.w_c2[t]);
    }
    if (!w->out_w) {
      split3(st, w->q_w0, latent, hh, latent, 1, ws.w_q0);
      split3(st, w->v_w0, latent, hh, latent, 1, ws.w_v0);
      split3(st, w->q_w1, hh, hh, hh, 1, ws.w_q1);
      split3(st, w->v_w1, hh, hh, hh, 1, ws.w_v1);
    }
  }
  auto gemm32 = [&](const float* A, int lda, const float* obs_scale, const float* Wt, const __nv_bfloat16* Wsplit, int ldw,
                    const float* bias, float* C, int ldc, int M, const int* m_dev, int Nout, int K, int relu) {
    if (tc) {
      const int rc = tcgemm(st, sms, A, lda, obs_scale, a->obs_stride, N, Wsplit, bias, C, ldc, M, m_dev, Nout, K, relu, ws.asplit);
      if (rc) rc_g = rc;
    } else {
      sgemm(st, A, lda, obs_scale, a->obs_stride, N, Wt, ldw, bias, C, ldc, M, m_dev, Nout, K, relu);
    }
  };
  auto edge = [&](const float* P, const float* obs, int rows, const float* att, const float* bias, float* out) {
    if (tr) {
      switch (Wn) {
        case 1: launch_tr_edge<1>(st, P, nproj * HC, obs, a->obs_stride, N, rows, H, out, HC); break;
        case 2: launch_tr_edge<2>(st, P, nproj * HC, obs, a->obs_stride, N, rows, H, out, HC); break;
        case 4: launch_tr_edge<4>(st, P, nproj * HC, obs, a->obs_stride, N, rows, H, out, HC); break;
        default: launch_tr_edge<8>(st, P, nproj * HC, obs, a->obs_stride, N, rows, H, out, HC); break;
      }
    } else {
      switch (Wn) {
        case 1: launch_gat_edge<1>(st, P, nproj * HC, obs, a->obs_stride, N, rows, H, att, bias, out, HC); break;
        case 2: launch_gat_edge<2>(st, P, nproj * HC, obs, a->obs_stride, N, rows, H, att, bias, out, HC); break;
        case 4: launch_gat_edge<4>(st, P, nproj * HC, obs, a->obs_stride, N, rows, H, att, bias, out, HC); break;
        default: launch_gat_edge<8>(st, P, nproj * HC, obs, a->obs_stride, N, rows, H, att, bias, out, HC); break;
      }
    }
  };
  // projection weights are separate tensors in the state_dict: one GEMM per tensor, written side by side
  auto project = [&](const float* X, int K, const float* obs_scale, int prof_id, const float* wa, const float* ba,
                     const float* wb, const float* bb, const float* wc, const float* bc, int rows, __nv_bfloat16* const* wsplit) {
    const float* ws_[3] = {wa, wb, wc};
    const float* bs_[3] = {ba, bb, bc};
    if (tc) split3(st, X, K, rows, K, 0, ws.asplit);                 // one split of the input serves all projections
    for (int t = 0; t < nproj; ++t) {
      if (t == 0) prof_begin(prof_id);
      if (tc) {
        GemmEpilogue e{};
        e.bias = bs_[t]; e.obs = obs_scale; e.obs_stride = a->obs_stride; e.nodes = N; e.Cf = ws.P + (size_t)t * HC; e.ldcf = nproj * HC;
        const int rc = gemm_bf16_launch(ws.asplit, 6 * K, wsplit[t], 6 * K, GemmShape{rows, HC, 6 * K, nullptr}, e, sms, st);
        if (rc) rc_g = rc;
      } else {
        sgemm(st, X, K, obs_scale, a->obs_stride, N, ws_[t], K, bs_[t], ws.P + (size_t)t * HC, nproj * HC, rows, nullptr, HC, K, 0);
      }
      if (t == 0) prof_end(prof_id);
    }
  };

  for (int g0 = 0; g0 < a->n_graphs; g0 += Gc) {
    const int gc = (a->n_graphs - g0) < Gc ? (a->n_graphs - g0) : Gc;
    const int rows = gc * N;
    const float* obs = a->obs + (int64_t)g0 * a->obs_stride;
    const uint8_t* cm = a->ctrl_mode == 0 ? a->ctrl_mask + (size_t)g0 * N : nullptr;
    // encoder (tianshou MLP + extra relu): x0 = relu(W1 relu(W0 f + b0) + b1)
    {
      dim3 blk(32, 8);
      enc0_kernel<<<(rows + 7) / 8, blk, 0, st>>>(obs, a->obs_stride, N, rows, d->input_dim, w->enc_w0, w->enc_b0, hid, ws.h);
      mls_count_launch();
      gemm32(ws.h, hid, nullptr, w->enc_w1, ws.w_enc1, hid, w->enc_b1, ws.x0, hid, rows, nullptr, hid, hid, 1);
    }
    // conv1 (+relu)
    if (tr) project(ws.x0, hid, nullptr, MLS_PROF_PROJ1, w->c1_wa, w->c1_ba, w->c1_wb, w->c1_bb, w->c1_wc, w->c1_bc, rows, ws.w_c1);
    else project(ws.x0, hid, nullptr, MLS_PROF_PROJ1, w->c1_wa, w->c1_ba, w->c1_wb, w->c1_bb, nullptr, nullptr, rows, ws.w_c1);
    prof_begin(MLS_PROF_EDGE1);
    edge(ws.P, obs, rows, w->c1_att, w->c1_bias, ws.x1);
    prof_end(MLS_PROF_EDGE1);
    if (!hl) {
      // conv2 on x1 * dm (+relu); the x1 snapshot used by the head stays unmasked
      if (tr) project(ws.x1, HC, obs, MLS_PROF_PROJ2, w->c2_wa, w->c2_ba, w->c2_wb, w->c2_bb, w->c2_wc, w->c2_bc, rows, ws.w_c2);
      else project(ws.x1, HC, obs, MLS_PROF_PROJ2, w->c2_wa, w->c2_ba, w->c2_wb, w->c2_bb, nullptr, nullptr, rows, ws.w_c2);
      prof_begin(MLS_PROF_EDGE2);
      edge(ws.P, obs, rows, w->c2_att, w->c2_bias, ws.x2);
      prof_end(MLS_PROF_EDGE2);
      MLS_CUDA(cudaMemsetAsync(ws.count, 0, sizeof(int), st));
      ctrl_list_kernel<<<(gc * 32 + 255) / 256, 256, 0, st>>>(cm, obs, a->obs_stride, N, gc, a->ctrl_mode, ws.idx, ws.count);
      gather_kernel<<<rows, 128, 0, st>>>(ws.idx, ws.count, ws.x0, hid, ws.x1, HC, ws.x2, HC, ws.z);
      mls_count_launch(2);
    } else {
      pool_kernel<<<gc, 128, 0, st>>>(ws.x1, HC, obs, a->obs_stride, N, HC, d->pool, ws.z);
      mls_count_launch();
    }
    const int head_rows = hl ? gc : rows;
    const int* m_dev = hl ? nullptr : ws.count;
    if (w->out_w) {                                           // no dueling heads: one linear layer on the latent row
      if (!hl) {
        head_linear_kernel<<<(rows * 32 + 255) / 256, 256, 0, st>>>(ws.z, latent, ws.idx, ws.count, rows, w->out_w, w->out_b,
                                                                    (int64_t)g0 * N, N, a->q, a->act, a->ctrl_mode, aa);
        mls_count_launch();
      } else {
        head_linear_kernel<<<(gc * 32 + 255) / 256, 256, 0, st>>>(ws.z, latent, nullptr, nullptr, gc, w->out_w, w->out_b, 0, N, ws.qg,
                                                                  nullptr, 2, aa);
        const long long nthr = a->ctrl_mode == 1 ? gc : (long long)gc * N;
        hl_scatter_kernel<<<(unsigned)((nthr + 255) / 256), 256, 0, st>>>(ws.qg, cm, N, gc, g0, a->ctrl_mode, a->q, a->act, aa);
        mls_count_launch(2);
      }
      MLS_LAUNCH_CHECK();
      first_chunk = false;
      continue;
    }
    // dueling head: Q = MLP(latent->hh->hh->2), V = MLP(latent->hh->hh->1)
    prof_begin(MLS_PROF_HEAD0);
    gemm32(ws.z, latent, nullptr, w->q_w0, ws.w_q0, latent, w->q_b0, ws.hid1, 2 * hh, head_rows, m_dev, hh, latent, 1);
    prof_end(MLS_PROF_HEAD0);
    gemm32(ws.z, latent, nullptr, w->v_w0, ws.w_v0, latent, w->v_b0, ws.hid1 + hh, 2 * hh, head_rows, m_dev, hh, latent, 1);
    gemm32(ws.hid1, 2 * hh, nullptr, w->q_w1, ws.w_q1, hh, w->q_b1, ws.hid2, 2 * hh, head_rows, m_dev, hh, hh, 1);
    gemm32(ws.hid1 + hh, 2 * hh, nullptr, w->v_w1, ws.w_v1, hh, w->v_b1, ws.hid2 + hh, 2 * hh, head_rows, m_dev, hh, hh, 1);
    if (rc_g) return rc_g;
    if (!hl) {
      head_out_kernel<<<(rows * 32 + 255) / 256, 256, 0, st>>>(ws.hid2, hh, ws.idx, ws.count, rows, w->q_w2, w->q_b2, w->v_w2,
                                                               w->v_b2, (int64_t)g0 * N, N, a->q, a->act, a->ctrl_mode, aa);
      mls_count_launch();
    } else {
      head_out_kernel<<<(gc * 32 + 255) / 256, 256, 0, st>>>(ws.hid2, hh, nullptr, nullptr, gc, w->q_w2, w->q_b2, w->v_w2,
                                                             w->v_b2, 0, N, ws.qg, nullptr, 2, aa);
      const long long nthr = a->ctrl_mode == 1 ? gc : (long long)gc * N;
      hl_scatter_kernel<<<(unsigned)((nthr + 255) / 256), 256, 0, st>>>(ws.qg, cm, N, gc, g0, a->ctrl_mode, a->q, a->act, aa);
      mls_count_launch(2);
    }
    MLS_LAUNCH_CHECK();
    first_chunk = false;
  }
  return MLS_OK;
}
