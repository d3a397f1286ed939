// Training-side GATv2 attention (forward + backward) on per-sample edge lists (sm_100a).
//
// The gradient step evaluates the Q-network on SAMPLED agent observations (tianshou DQNPolicy.learn; reference
// policies/dgn.py:22-71, l_dgn.py:92-151).  One observation controls ONE node c, so only what c reads is evaluated:
//     conv2 at c                        <- sources  S1 = {c} + radius-neighbours(c)
//     conv1 at the nodes of S1          <- sources  {i} + radius-neighbours(i)
// (same structure as melissa_b200/networks/autograd.py).  The dense layers stay torch GEMMs; the edge phase of
// PyG's GATv2Conv -- gather x_l / x_r per edge, leaky_relu, <att, .>, segment softmax exp(e - max) / (sum + 1e-16),
// weighted scatter-add, and the backward of all that -- is three kernels here instead of ~60 torch launches over
// [edges, heads, 128] tensors:
//   train_lists_kernel   radius rule of torch_cluster.radius_graph (candidates in index order, fma(dy, dy, dx * dx) < r^2 in
//                        fp32, at most 32 neighbours) -> S1 slots and, per slot, its source rows (self loop first)
//   gatv2_fwd_kernel     one warp per (target, head): logits, softmax, aggregation; alpha kept for the backward
//   gatv2_bwd_kernel     d x_l (atomic: a source feeds several targets), d x_r, d att (per-block partial sums)
// All fp32.  Nothing here is used by the rollout path.
#include "common.cuh"

namespace {

constexpr int kTC = 128;         // channels per head (hidden_dim of every reference script)
constexpr int kCap = 36;         // list capacity: 1 self + 32 neighbours, rounded up

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Neighbours of node i in index order under the radius rule, emitted through `emit(position, j)`; returns their count.
template <typename Emit>
__device__ __forceinline__ int scan_neighbours(const float* __restrict__ row, int N, int i, float r2, int lane, Emit emit) {
  const float xi = row[i * 8], yi = row[i * 8 + 1];
  int hits = 0, emitted = 0;                     // hits counts the node itself too (radius_graph drops loops afterwards)
  for (int j0 = 0; j0 < N; j0 += 32) {
    const int j = j0 + lane;
    bool h = false;
    if (j < N) {
      const float dx = row[j * 8] - xi, dy = row[j * 8 + 1] - yi;
      h = fmaf(dy, dy, dx * dx) < r2;
    }
    const uint32_t bal = __ballot_sync(0xffffffffu, h);
    const uint32_t lt = (1u << lane) - 1u;
    const bool keep = h && (hits + __popc(bal & lt) + 1 <= 33) && j != i;
    const uint32_t kb = __ballot_sync(0xffffffffu, keep);
    if (keep) emit(emitted + __popc(kb & lt), j);
    hits += __popc(bal);
    emitted += __popc(kb);
  }
  return emitted;
}

// mode 0: s1_cnt[b] = |S1|.   mode 1: fill the compact slot arrays at slot_base[b].   mode 2: lists of ALL nodes, slot = node row.
__global__ void __launch_bounds__(128) train_lists_kernel(const float* __restrict__ obs, long long stride, int bs, int N, float r2, int mode,
                                                          int* __restrict__ s1_cnt, const long long* __restrict__ slot_base,
                                                          int* __restrict__ tgt_row, int* __restrict__ src_row, int* __restrict__ src_cnt,
                                                          uint8_t* __restrict__ used) {
  __shared__ int s1[4][kCap];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * 4 + warp;
  if (b >= bs) return;
  const float* row = obs + (long long)b * stride;
  if (mode == 2) {                               // every node is a target (HL-DGN): slot = node row, no compaction
    for (int i = 0; i < N; ++i) {
      const long long slot = (long long)b * N + i;
      int* dst = src_row + slot * kCap;
      const int d = scan_neighbours(row, N, i, r2, lane, [&](int pos, int j) { dst[1 + pos] = b * N + j; });
      if (lane == 0) { tgt_row[slot] = b * N + i; dst[0] = b * N + i; src_cnt[slot] = 1 + d; }
    }
    return;
  }
  int c = (int)row[(long long)N * 8];
  c = c < 0 ? 0 : (c > N - 1 ? N - 1 : c);
  int* my = s1[warp];
  if (lane == 0) my[0] = c;
  const int deg = scan_neighbours(row, N, c, r2, lane, [&](int pos, int j) { my[1 + pos] = j; });
  const int cnt = 1 + deg;
  if (mode == 0) {
    if (lane == 0) s1_cnt[b] = cnt;
    return;
  }
  __syncwarp();
  const long long base = slot_base[b];
  for (int s = 0; s < cnt; ++s) {
    const int i = my[s];
    const long long slot = base + s;
    int* dst = src_row + slot * kCap;
    const int d = scan_neighbours(row, N, i, r2, lane, [&](int pos, int j) {
      dst[1 + pos] = b * N + j;
      if (used) used[b * N + j] = 1;             // node rows whose encoder output / conv1 source projection is read
    });
    if (lane == 0) {
      tgt_row[slot] = b * N + i;
      dst[0] = b * N + i;                        // GATv2Conv(add_self_loops=True)
      src_cnt[slot] = 1 + d;
      if (used) used[b * N + i] = 1;
    }
  }
}

struct EdgeArgs {
  const float* xl; long long ldl;                // source-side projections, one row per source
  const float* xr; long long ldr;                // target-side projections
  const float* att;                              // [H][128]
  const int* tgt_row;                            // [T] row of xr (negative: no such target, output zeros)
  const int* src_row;                            // [T][kCap] rows of xl
  const int* src_cnt;                            // [T]
  int T, H;
};

// blockDim = 32 * H: warp = head.  lane = 4 channels.
__global__ void __launch_bounds__(128) gatv2_fwd_kernel(const EdgeArgs a, float* __restrict__ out, float* __restrict__ alpha) {
  __shared__ float lg[4][kCap];
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int HC = a.H * kTC, col = h * kTC + lane * 4;
  const float4 att = *reinterpret_cast<const float4*>(a.att + col);
  float* my = lg[h];
  for (int t = blockIdx.x; t < a.T; t += gridDim.x) {
    const int r = a.tgt_row[t];
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r >= 0) {
      const int cnt = a.src_cnt[t];
      const int* src = a.src_row + (long long)t * kCap;
      const float4 xr = *reinterpret_cast<const float4*>(a.xr + (long long)r * a.ldr + col);
      float mx = -INFINITY;
      for (int e = 0; e < cnt; ++e) {
        const float4 xs = *reinterpret_cast<const float4*>(a.xl + (long long)src[e] * a.ldl + col);
        float v0 = xs.x + xr.x, v1 = xs.y + xr.y, v2 = xs.z + xr.z, v3 = xs.w + xr.w;
        v0 = v0 > 0.f ? v0 : 0.2f * v0; v1 = v1 > 0.f ? v1 : 0.2f * v1; v2 = v2 > 0.f ? v2 : 0.2f * v2; v3 = v3 > 0.f ? v3 : 0.2f * v3;
        const float p = warp_sum(fmaf(v0, att.x, fmaf(v1, att.y, fmaf(v2, att.z, v3 * att.w))));
        if (lane == 0) my[e] = p;
        mx = fmaxf(mx, p);
      }
      __syncwarp();
      float sum = 0.f;
      for (int e = 0; e < cnt; ++e) sum += expf(my[e] - mx);
      const float inv = 1.f / (sum + 1e-16f);
      for (int e = 0; e < cnt; ++e) {
        const float w = expf(my[e] - mx) * inv;
        const float4 xs = *reinterpret_cast<const float4*>(a.xl + (long long)src[e] * a.ldl + col);
        acc.x = fmaf(w, xs.x, acc.x); acc.y = fmaf(w, xs.y, acc.y); acc.z = fmaf(w, xs.z, acc.z); acc.w = fmaf(w, xs.w, acc.w);
        if (lane == 0) alpha[((long long)t * kCap + e) * a.H + h] = w;
      }
      __syncwarp();
    }
    *reinterpret_cast<float4*>(out + (long long)t * HC + col) = acc;
  }
}

__global__ void __launch_bounds__(128) gatv2_bwd_kernel(const EdgeArgs a, const float* __restrict__ alpha, const float* __restrict__ dout,
                                                        float* __restrict__ d_xl, float* __restrict__ d_xr, float* __restrict__ d_att_part) {
  __shared__ float ge[4][kCap];
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int HC = a.H * kTC, col = h * kTC + lane * 4;
  const float4 att = *reinterpret_cast<const float4*>(a.att + col);
  float* my = ge[h];
  float4 datt = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int t = blockIdx.x; t < a.T; t += gridDim.x) {
    const int r = a.tgt_row[t];
    if (r < 0) continue;
    const int cnt = a.src_cnt[t];
    const int* src = a.src_row + (long long)t * kCap;
    const float* al = alpha + (long long)t * kCap * a.H + h;
    const float4 g = *reinterpret_cast<const float4*>(dout + (long long)t * HC + col);
    const float4 xr = *reinterpret_cast<const float4*>(a.xr + (long long)r * a.ldr + col);
    float dot = 0.f;
    for (int e = 0; e < cnt; ++e) {
      const float4 xs = *reinterpret_cast<const float4*>(a.xl + (long long)src[e] * a.ldl + col);
      const float p = warp_sum(fmaf(g.x, xs.x, fmaf(g.y, xs.y, fmaf(g.z, xs.z, g.w * xs.w))));
      if (lane == 0) my[e] = p;
      dot = fmaf(al[e * a.H], p, dot);
    }
    __syncwarp();
    float4 dxr = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int e = 0; e < cnt; ++e) {
      const float w = al[e * a.H];
      const float dl = w * (my[e] - dot);                  // d loss / d logit_e
      const float4 xs = *reinterpret_cast<const float4*>(a.xl + (long long)src[e] * a.ldl + col);
      const float v0 = xs.x + xr.x, v1 = xs.y + xr.y, v2 = xs.z + xr.z, v3 = xs.w + xr.w;
      const float s0 = v0 > 0.f ? 1.f : 0.2f, s1 = v1 > 0.f ? 1.f : 0.2f, s2 = v2 > 0.f ? 1.f : 0.2f, s3 = v3 > 0.f ? 1.f : 0.2f;
      const float d0 = dl * att.x * s0, d1 = dl * att.y * s1, d2 = dl * att.z * s2, d3 = dl * att.w * s3;
      float* dst = d_xl + (long long)src[e] * HC + col;
      atomicAdd(dst, fmaf(w, g.x, d0)); atomicAdd(dst + 1, fmaf(w, g.y, d1));
      atomicAdd(dst + 2, fmaf(w, g.z, d2)); atomicAdd(dst + 3, fmaf(w, g.w, d3));
      dxr.x += d0; dxr.y += d1; dxr.z += d2; dxr.w += d3;
      datt.x = fmaf(dl, v0 * s0, datt.x); datt.y = fmaf(dl, v1 * s1, datt.y);
      datt.z = fmaf(dl, v2 * s2, datt.z); datt.w = fmaf(dl, v3 * s3, datt.w);
    }
    __syncwarp();
    *reinterpret_cast<float4*>(d_xr + (long long)r * HC + col) = dxr;     // one target per x_r row
  }
  *reinterpret_cast<float4*>(d_att_part + (long long)blockIdx.x * HC + col) = datt;
}

// ---- TransformerConv(root_weight=False, beta=False) edge phase (DGN-R): logit_e = <q_t, k_e> / sqrt(128), out = sum_e a_e v_e;
// no self loop: the lists' first entry (the node itself) is skipped.
struct TrArgs {
  const float* k; const float* v; long long lds;   // source-side projections
  const float* q; long long ldq;                   // target-side projection
  const int* tgt_row; const int* src_row; const int* src_cnt;
  int T, H;
};

__global__ void __launch_bounds__(128) transformer_fwd_kernel(const TrArgs a, float* __restrict__ out, float* __restrict__ alpha) {
  __shared__ float lg[4][kCap];
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int HC = a.H * kTC, col = h * kTC + lane * 4;
  const float scale = rsqrtf((float)kTC);
  float* my = lg[h];
  for (int t = blockIdx.x; t < a.T; t += gridDim.x) {
    const int r = a.tgt_row[t];
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r >= 0) {
      const int cnt = a.src_cnt[t];
      const int* src = a.src_row + (long long)t * kCap;
      const float4 q = *reinterpret_cast<const float4*>(a.q + (long long)r * a.ldq + col);
      float mx = -INFINITY;
      for (int e = 1; e < cnt; ++e) {
        const float4 ke = *reinterpret_cast<const float4*>(a.k + (long long)src[e] * a.lds + col);
        const float p = warp_sum(fmaf(q.x, ke.x, fmaf(q.y, ke.y, fmaf(q.z, ke.z, q.w * ke.w)))) * scale;
        if (lane == 0) my[e] = p;
        mx = fmaxf(mx, p);
      }
      __syncwarp();
      float sum = 0.f;
      for (int e = 1; e < cnt; ++e) sum += expf(my[e] - mx);
      const float inv = 1.f / (sum + 1e-16f);
      for (int e = 1; e < cnt; ++e) {
        const float w = expf(my[e] - mx) * inv;
        const float4 ve = *reinterpret_cast<const float4*>(a.v + (long long)src[e] * a.lds + col);
        acc.x = fmaf(w, ve.x, acc.x); acc.y = fmaf(w, ve.y, acc.y); acc.z = fmaf(w, ve.z, acc.z); acc.w = fmaf(w, ve.w, acc.w);
        if (lane == 0) alpha[((long long)t * kCap + e) * a.H + h] = w;
      }
      __syncwarp();
    }
    *reinterpret_cast<float4*>(out + (long long)t * HC + col) = acc;
  }
}

__global__ void __launch_bounds__(128) transformer_bwd_kernel(const TrArgs a, const float* __restrict__ alpha, const float* __restrict__ dout,
                                                              float* __restrict__ d_k, float* __restrict__ d_v, float* __restrict__ d_q) {
  __shared__ float ge[4][kCap];
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int HC = a.H * kTC, col = h * kTC + lane * 4;
  const float scale = rsqrtf((float)kTC);
  float* my = ge[h];
  for (int t = blockIdx.x; t < a.T; t += gridDim.x) {
    const int r = a.tgt_row[t];
    if (r < 0) continue;
    const int cnt = a.src_cnt[t];
    const int* src = a.src_row + (long long)t * kCap;
    const float* al = alpha + (long long)t * kCap * a.H + h;
    const float4 g = *reinterpret_cast<const float4*>(dout + (long long)t * HC + col);
    const float4 q = *reinterpret_cast<const float4*>(a.q + (long long)r * a.ldq + col);
    float dot = 0.f;
    for (int e = 1; e < cnt; ++e) {
      const float4 ve = *reinterpret_cast<const float4*>(a.v + (long long)src[e] * a.lds + col);
      const float p = warp_sum(fmaf(g.x, ve.x, fmaf(g.y, ve.y, fmaf(g.z, ve.z, g.w * ve.w))));
      if (lane == 0) my[e] = p;
      dot = fmaf(al[e * a.H], p, dot);
    }
    __syncwarp();
    float4 dq = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int e = 1; e < cnt; ++e) {
      const float w = al[e * a.H];
      const float dl = w * (my[e] - dot) * scale;          // d loss / d <q, k_e>
      const float4 ke = *reinterpret_cast<const float4*>(a.k + (long long)src[e] * a.lds + col);
      float* dv = d_v + (long long)src[e] * HC + col;
      float* dk = d_k + (long long)src[e] * HC + col;
      atomicAdd(dv, w * g.x); atomicAdd(dv + 1, w * g.y); atomicAdd(dv + 2, w * g.z); atomicAdd(dv + 3, w * g.w);
      atomicAdd(dk, dl * q.x); atomicAdd(dk + 1, dl * q.y); atomicAdd(dk + 2, dl * q.z); atomicAdd(dk + 3, dl * q.w);
      dq.x = fmaf(dl, ke.x, dq.x); dq.y = fmaf(dl, ke.y, dq.y); dq.z = fmaf(dl, ke.z, dq.z); dq.w = fmaf(dl, ke.w, dq.w);
    }
    __syncwarp();
    *reinterpret_cast<float4*>(d_q + (long long)r * HC + col) = dq;
  }
}

int edge_grid(int T) {
  int g = 148 * 8;
  return T < g ? (T > 0 ? T : 1) : g;
}

}  // namespace

extern "C" int mls_train_list_capacity(void) { return kCap; }

extern "C" int mls_train_lists(const float* obs_rows, int64_t row_stride, int32_t n_samples, int32_t n_nodes, float r2, int32_t mode,
                               int32_t* s1_cnt, const int64_t* slot_base, int32_t* tgt_row, int32_t* src_row, int32_t* src_cnt,
                               uint8_t* used, void* stream) {
  MLS_CHECK_ARG(obs_rows && n_nodes > 0 && row_stride >= (int64_t)n_nodes * 8 + 1, "bad observation rows");
  MLS_CHECK_ARG(mode == 0 ? s1_cnt != nullptr : ((slot_base || mode == 2) && tgt_row && src_row && src_cnt), "NULL output");
  MLS_CHECK_ARG((int64_t)n_samples * n_nodes < (1ll << 31), "too many node rows for 32-bit row indices");
  if (n_samples <= 0) return MLS_OK;
  train_lists_kernel<<<(n_samples + 3) / 4, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      obs_rows, row_stride, n_samples, n_nodes, r2, mode, s1_cnt, reinterpret_cast<const long long*>(slot_base), tgt_row, src_row, src_cnt, used);
  mls_count_launch();
  MLS_LAUNCH_CHECK();
  return MLS_OK;
}

extern "C" int mls_gatv2_edge_fwd(const float* xl, int64_t ldl, const float* xr, int64_t ldr, const float* att, const int32_t* tgt_row,
                                  const int32_t* src_row, const int32_t* src_cnt, int32_t n_targets, int32_t heads, float* out,
                                  float* alpha, void* stream) {
  MLS_CHECK_ARG(xl && xr && att && tgt_row && src_row && src_cnt && out && alpha, "NULL argument");
  MLS_CHECK_ARG(heads >= 1 && heads <= 4 && ldl % 4 == 0 && ldr % 4 == 0, "heads must be 1..4 and rows 16-byte aligned");
  if (n_targets <= 0) return MLS_OK;
  EdgeArgs a{xl, ldl, xr, ldr, att, tgt_row, src_row, src_cnt, n_targets, heads};
  gatv2_fwd_kernel<<<edge_grid(n_targets), 32 * heads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a, out, alpha);
  mls_count_launch();
  MLS_LAUNCH_CHECK();
  return MLS_OK;
}

extern "C" int mls_gatv2_edge_bwd_blocks(int32_t n_targets) { return edge_grid(n_targets); }

extern "C" int mls_gatv2_edge_bwd(const float* xl, int64_t ldl, const float* xr, int64_t ldr, const float* att, const int32_t* tgt_row,
                                  const int32_t* src_row, const int32_t* src_cnt, int32_t n_targets, int32_t heads, const float* alpha,
                                  const float* dout, float* d_xl, float* d_xr, float* d_att_part, void* stream) {
  MLS_CHECK_ARG(xl && xr && att && tgt_row && src_row && src_cnt && alpha && dout && d_xl && d_xr && d_att_part, "NULL argument");
  MLS_CHECK_ARG(heads >= 1 && heads <= 4 && ldl % 4 == 0 && ldr % 4 == 0, "heads must be 1..4 and rows 16-byte aligned");
  if (n_targets <= 0) return MLS_OK;
  EdgeArgs a{xl, ldl, xr, ldr, att, tgt_row, src_row, src_cnt, n_targets, heads};
  gatv2_bwd_kernel<<<edge_grid(n_targets), 32 * heads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a, alpha, dout, d_xl, d_xr, d_att_part);
  mls_count_launch();
  MLS_LAUNCH_CHECK();
  return MLS_OK;
}

extern "C" int mls_transformer_edge_fwd(const float* k, const float* v, int64_t lds, const float* q, int64_t ldq, const int32_t* tgt_row,
                                        const int32_t* src_row, const int32_t* src_cnt, int32_t n_targets, int32_t heads, float* out,
                                        float* alpha, void* stream) {
  MLS_CHECK_ARG(k && v && q && tgt_row && src_row && src_cnt && out && alpha, "NULL argument");
  MLS_CHECK_ARG(heads >= 1 && heads <= 4 && lds % 4 == 0 && ldq % 4 == 0, "heads must be 1..4 and rows 16-byte aligned");
  if (n_targets <= 0) return MLS_OK;
  TrArgs a{k, v, lds, q, ldq, tgt_row, src_row, src_cnt, n_targets, heads};
  transformer_fwd_kernel<<<edge_grid(n_targets), 32 * heads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a, out, alpha);
  mls_count_launch();
  MLS_LAUNCH_CHECK();
  return MLS_OK;
}

extern "C" int mls_transformer_edge_bwd(const float* k, const float* v, int64_t lds, const float* q, int64_t ldq, const int32_t* tgt_row,
                                        const int32_t* src_row, const int32_t* src_cnt, int32_t n_targets, int32_t heads,
                                        const float* alpha, const float* dout, float* d_k, float* d_v, float* d_q, void* stream) {
  MLS_CHECK_ARG(k && v && q && tgt_row && src_row && src_cnt && alpha && dout && d_k && d_v && d_q, "NULL argument");
  MLS_CHECK_ARG(heads >= 1 && heads <= 4 && lds % 4 == 0 && ldq % 4 == 0, "heads must be 1..4 and rows 16-byte aligned");
  if (n_targets <= 0) return MLS_OK;
  TrArgs a{k, v, lds, q, ldq, tgt_row, src_row, src_cnt, n_targets, heads};
  transformer_bwd_kernel<<<edge_grid(n_targets), 32 * heads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a, alpha, dout, d_k, d_v, d_q);
  mls_count_launch();
  MLS_LAUNCH_CHECK();
  return MLS_OK;
}
