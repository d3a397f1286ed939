// conv2 attention (GATv2Conv / TransformerConv, second layer of L-DGN / DGN-R) on the compacted row sets (sm_100a).
//
// Only controlling nodes are conv2 targets (l_dgn.py:133-139, dgn_r.py:113-118: the network reads x[ctrl] only), and
// only their radius-graph sources are conv2 sources ("needed" rows, ctrl_need_list_kernel).  One work item =
// (graph, head); a persistent 128-thread CTA (head fixed per CTA) runs per item:
//   stage     cp.async of the head's 128-channel fp16 slices: the needed source rows into a [source][channel]
//             MN-major 128B-swizzled operand (GATv2: x_l, which is BOTH the logit operand and the value), the
//             target rows (x_r / q) and (Transformer) the key rows into padded rows; CSR block of the graph
//   logits    one LANE per edge (self loop first for GATv2), 128 channels in packed half2 math
//               GATv2:       e_ij = 0.6 (a_j + b_i) + 0.4 sum_c att_c |x_l[j,c] + x_r[i,c]|   (leaky_relu(s,.2) = .6 s + .4 |s|;
//                            the linear parts come from the projection GEMMs' epilogue dots, fp32)
//               Transformer: e_ij = <q_i, k_j> / sqrt(C)
//             all in base 2 (pre-scaled by log2 e)
//   softmax   4 lanes per target: 2^(e - max) / (sum + 1e-16) (PyG softmax), written as fp16 into the K-major
//             128B-swizzled weight matrix W[target][source]
//   aggregate ONE tcgen05.mma chain, transposed: out^T[channel][target] = X^T[channel][source] x W^T[source][target]
//             (M = 128 channels = the head, N = targets rounded to 16, K = 16 sources per step), fp32 in TMEM
//   epilogue  tcgen05.ld (lane = channel, registers = targets) + conv bias + ReLU -> bf16, 64 contiguous bytes per
//             warp store, straight into the controlling-node snapshot rows z[slot] (slots of a graph are consecutive)
// Reference math: PyG GATv2Conv / TransformerConv(root_weight=False) as called by l_dgn.py:133, dgn_r.py:113.
#include "conv2_attn.cuh"

#include "dgn_kernels.cuh"
#include "tcgen05_ptx.cuh"

namespace mls {

namespace {

constexpr int kThreads = 128;
constexpr int kPanel = 64 * 128;          // one 64-channel panel of the value operand: 64 source rows x 128 B
constexpr int kXBytes = 2 * kPanel;       // the head's 128 channels
constexpr int kWBytes = 64 * 128;         // W[target][source]: 64 rows of 64 fp16
constexpr int kRowPad = kC * 2 + 16;      // padded row of 128 fp16 (bank-conflict-free 16 B reads by row-owning lanes)
constexpr int kCols = 64;                 // TMEM columns per CTA

__device__ __forceinline__ void cp16(unsigned char* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcpf(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint16_t relu_bf16_bits(float x) {
  uint16_t d;
  asm("cvt.rn.relu.bf16.f32 %0, %1;" : "=h"(d) : "f"(x));
  return d;
}
// same descriptor helpers as attn_table.cu (SWIZZLE_128B; A operand MN-major)
__device__ __forceinline__ uint64_t desc_mn(uint32_t smem_addr, uint32_t lbo16, uint32_t sbo16) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo16 & 0x3FFFu) << 16;
  d |= (uint64_t)(sbo16 & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

size_t conv2_smem_bytes(int N, bool tr) {
  const int edge_cap = (N * (kMaxNbr + 1) + 15) & ~15;
  return 1024 + kXBytes + kWBytes + (size_t)64 * kRowPad * (tr ? 2 : 1) + 256 /*att*/ + 64 * kMaxNbr /*csr*/ + 16 /*barrier, tmem*/ +
         512 /*as, bt*/ + (size_t)edge_cap * 4 + 68 * 4 + 72 * 2 + 128 + (size_t)edge_cap * 2;
}

template <bool TR>
__global__ void __launch_bounds__(kThreads, TR ? 3 : 4) conv2_attn_kernel(const Conv2Args a, const int edge_cap) {
  extern __shared__ __align__(16) unsigned char sm_raw[];
  unsigned char* sm = sm_raw + ((1024u - (smem_u32(sm_raw) & 1023u)) & 1023u);
  unsigned char* sX = sm;                                   // values [source][channel], MN-major SW128 (GATv2: also logit operand)
  unsigned char* sW = sX + kXBytes;                         // weights [target][source], K-major SW128
  unsigned char* sT = sW + kWBytes;                         // targets (x_r / q): [64][kRowPad]
  unsigned char* sK = sT + 64 * kRowPad;                    // Transformer keys: [64][kRowPad] by node
  unsigned char* p = sK + (TR ? 64 * kRowPad : 0);
  __half* att_s = reinterpret_cast<__half*>(p); p += 256;
  uint8_t* s_src = p; p += 64 * kMaxNbr;
  uint64_t* bar = reinterpret_cast<uint64_t*>(p); p += 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(p); p += 8;
  float* s_as = reinterpret_cast<float*>(p); p += 256;
  float* s_bt = reinterpret_cast<float*>(p); p += 256;
  float* s_e = reinterpret_cast<float*>(p); p += (size_t)edge_cap * 4;
  int* s_eptr = reinterpret_cast<int*>(p); p += 68 * 4;
  uint16_t* s_ptr = reinterpret_cast<uint16_t*>(p); p += 72 * 2;
  uint8_t* s_tl = p; p += 64;
  uint8_t* s_nl = p; p += 64;
  uint8_t* s_esrc = p; p += edge_cap;
  uint8_t* s_etgt = p;

  const int N = a.N, H = a.H, HC = H * kC;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = blockIdx.x % H;                             // gridDim.x is a multiple of H: the head is fixed per CTA
  constexpr float kLog2e = 1.4426950408889634f;
  const float k06 = 0.6f * kLog2e;
  const float tr_scale = kLog2e / sqrtf((float)kC);

  // stale value rows only ever meet zero weights, so they just have to stay finite: zero once
  for (int u = tid; u < kXBytes / 16; u += kThreads) reinterpret_cast<uint4*>(sX)[u] = make_uint4(0, 0, 0, 0);
  if (!TR) att_s[tid] = __float2half_rn(a.att[h * kC + tid] * (0.4f * kLog2e));
  const float bias_c = a.bias ? a.bias[h * kC + tid] : 0.f;  // this thread's output channel in the epilogue
  if (tid == 0) { mbar_init(smem_u32(bar), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), kCols);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  uint32_t parity = 0;
  const int self = TR ? 0 : 1;

  for (int g = blockIdx.x / H; g < a.n_graphs; g += gridDim.x / H) {
    const int cnt = a.gcnt[g];
    if (cnt == 0) continue;                                  // uniform over the CTA
    const int first = a.gfirst[g], nf = a.nfirst[g], nc = a.ncnt[g];
    const int base = g * N;
    const size_t cg = a.graph_id ? (size_t)a.graph_id[(size_t)g * a.gid_stride] : (size_t)g;
    const int nmma = cnt <= 16 ? 16 : ((cnt + 15) & ~15);
    // ---------------------------------------------------------------- stage
    if (tid < nc) s_nl[tid] = (uint8_t)(a.nidx[nf + tid] - base);
    if (tid < cnt) {
      s_tl[tid] = (uint8_t)(a.idx[first + tid] - base);
      if (!TR) s_bt[tid] = a.bt[(size_t)(first + tid) * H + h] * k06;
    }
    {
      const uint16_t* gp = a.csr_ptr + cg * (N + 1);
      for (int t = tid; t <= N; t += kThreads) s_ptr[t] = gp[t];
      const uint8_t* gs = a.csr_src + cg * N * kMaxNbr;
      for (int t = tid; t < N * 2; t += kThreads) cp16(s_src + t * 16, gs + t * 16);
    }
    for (int u = tid; u < nmma * 8; u += kThreads) reinterpret_cast<uint4*>(sW)[u] = make_uint4(0, 0, 0, 0);
    for (int t = tid; t < cnt * 16; t += kThreads) {
      const int k = t >> 4, c = t & 15;
      cp16(sT + k * kRowPad + c * 16, a.Pt + (size_t)(first + k) * a.ldt + h * kC + c * 8);
    }
    __syncthreads();
    for (int t = tid; t < nc * 16; t += kThreads) {
      const int k = t >> 4, c = t & 15, j = s_nl[k];
      const __half* row = a.Ps + (size_t)(nf + k) * a.lds + h * kC + c * 8;
      unsigned char* xd = sX + (c >> 3) * kPanel + j * 128 + (((c & 7) ^ (j & 7)) << 4);
      if (TR) {
        cp16(sK + j * kRowPad + c * 16, row);
        cp16(xd, row + HC);
      } else {
        cp16(xd, row);
      }
    }
    if (!TR && tid < nc) s_as[s_nl[tid]] = a.as[(size_t)(nf + tid) * H + h] * k06;
    if (warp == 0) {                                         // edge offsets: exclusive scan of (degree + self) over the targets
      int run = 0;
      for (int c0 = 0; c0 < cnt; c0 += 32) {
        const int tk = c0 + lane;
        int v = 0;
        if (tk < cnt) { const int i = s_tl[tk]; v = (int)s_ptr[i + 1] - (int)s_ptr[i] + self; }
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += t;
        }
        if (tk < cnt) s_eptr[tk] = run + incl - v;
        run += __shfl_sync(0xffffffffu, incl, 31);
      }
      if (lane == 0) s_eptr[cnt] = run;
    }
    cp_wait_all();
    __syncthreads();
    if (tid < cnt) {                                         // edge list: (source node, target slot) per entry
      const int i = s_tl[tid], r0 = s_ptr[i], d = (int)s_ptr[i + 1] - r0;
      int e0 = s_eptr[tid];
      if (!TR) { s_esrc[e0] = (uint8_t)i; s_etgt[e0] = (uint8_t)tid; ++e0; }
      for (int k = 0; k < d; ++k) { s_esrc[e0 + k] = s_src[r0 + k]; s_etgt[e0 + k] = (uint8_t)tid; }
    }
    __syncthreads();
    // ---------------------------------------------------------------- logits: one lane per edge
    const int E = s_eptr[cnt];
    for (int e = tid; e < E; e += kThreads) {
      const int tk = s_etgt[e], j = s_esrc[e];
      const unsigned char* tr = sT + tk * kRowPad;
      __half2 acc[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[q] = __float2half2_rn(0.f);
      if (!TR) {
        const unsigned char* xr = sX + j * 128;
        const int jx = j & 7;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const uint4 xs = *reinterpret_cast<const uint4*>(xr + (c >> 3) * kPanel + (((c & 7) ^ jx) << 4));
          const uint4 xt = *reinterpret_cast<const uint4*>(tr + c * 16);
          const uint4 at = *reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(att_s) + c * 16);
          const __half2* x2 = reinterpret_cast<const __half2*>(&xs);
          const __half2* t2 = reinterpret_cast<const __half2*>(&xt);
          const __half2* a2 = reinterpret_cast<const __half2*>(&at);
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[q] = __hfma2(a2[q], __habs2(__hadd2(x2[q], t2[q])), acc[q]);
        }
      } else {
        const unsigned char* kr = sK + j * kRowPad;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const uint4 xs = *reinterpret_cast<const uint4*>(kr + c * 16);
          const uint4 xt = *reinterpret_cast<const uint4*>(tr + c * 16);
          const __half2* x2 = reinterpret_cast<const __half2*>(&xs);
          const __half2* t2 = reinterpret_cast<const __half2*>(&xt);
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[q] = __hfma2(x2[q], t2[q], acc[q]);
        }
      }
      float s = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) { const float2 f = __half22float2(acc[q]); s += f.x + f.y; }
      s_e[e] = TR ? s * tr_scale : s + (s_as[j] + s_bt[tk]);
    }
    __syncthreads();
    // ---------------------------------------------------------------- softmax -> W (4 lanes per target)
    for (int it = 0; it * kThreads < cnt * 4; ++it) {
      const int u = tid + it * kThreads, tk = u >> 2, l = u & 3;
      int lo = 0, hi = 0;
      if (tk < cnt) { lo = s_eptr[tk]; hi = s_eptr[tk + 1]; }
      float mx = -INFINITY;
      for (int x = lo + l; x < hi; x += 4) mx = fmaxf(mx, s_e[x]);
      __syncwarp();
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      float sum = 0.f;
      for (int x = lo + l; x < hi; x += 4) { const float pv = ex2f(s_e[x] - mx); s_e[x] = pv; sum += pv; }
      __syncwarp();
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      const float inv = rcpf(sum + 1e-16f);
      unsigned char* wrow = sW + tk * 128;
      for (int x = lo + l; x < hi; x += 4) {
        const int j = s_esrc[x];
        *reinterpret_cast<__half*>(wrow + ((((j >> 3) ^ tk) & 7) << 4) + (j & 7) * 2) = __float2half_rn(s_e[x] * inv);
      }
    }
    fence_proxy_async();
    __syncthreads();
    // ---------------------------------------------------------------- aggregate on the tensor core
    if (tid == 0) {
      tc_fence_after();
      // both operands fp16 (format 0), A (values) MN-major (bit 15), see attn_table.cu
      const uint32_t idesc = (make_idesc(128, nmma) & ~((7u << 7) | (7u << 10))) | (1u << 15);
      const uint64_t dv = desc_mn(smem_u32(sX), kPanel >> 4, 1024 >> 4);
      const uint64_t dw = make_smem_desc(smem_u32(sW));
      const int ksteps = (N + 15) >> 4;
      for (int k = 0; k < ksteps; ++k) umma_bf16(tmem_base, dv + (uint64_t)(k * 128), dw + (uint64_t)(k * 2), idesc, k ? 1u : 0u);
      umma_commit(smem_u32(bar));
    }
    mbar_wait(smem_u32(bar), parity);
    parity ^= 1u;
    tc_fence_after();
    // ---------------------------------------------------------------- epilogue: lane = channel, registers = targets
    {
      uint16_t* zo = reinterpret_cast<uint16_t*>(a.z) + (size_t)first * a.ldz + a.z_col + h * kC + tid;
      for (int c0 = 0; c0 < nmma; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
        for (int t = 0; t < 32; ++t)
          if (c0 + t < cnt) zo[(size_t)(c0 + t) * a.ldz] = relu_bf16_bits(__uint_as_float(v[t]) + bias_c);
      }
    }
    tc_fence_before();
    __syncthreads();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, kCols);
}

template <bool TR>
int launch(const Conv2Args& a, int sm_count, cudaStream_t st) {
  const size_t smem = conv2_smem_bytes(a.N, TR);
  const int edge_cap = (a.N * (kMaxNbr + 1) + 15) & ~15;
  static size_t configured = 0;
  if (smem > configured) {
    MLS_CUDA(cudaFuncSetAttribute(conv2_attn_kernel<TR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MLS_CUDA(cudaFuncSetAttribute(conv2_attn_kernel<TR>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    configured = smem;
  }
  const int per_sm = TR ? 3 : 4;
  long long items = (long long)a.n_graphs * a.H;
  long long grid = (long long)sm_count * per_sm;
  grid -= grid % a.H;
  if (grid < a.H) grid = a.H;
  if (grid > items) grid = items;                           // items is a multiple of H
  if (grid <= 0) return MLS_OK;
  conv2_attn_kernel<TR><<<(unsigned)grid, kThreads, smem, st>>>(a, edge_cap);
  mls_count_launch();
  MLS_LAUNCH_CHECK();
  return MLS_OK;
}

}  // namespace

int conv2_attn_launch(const Conv2Args& a, int sm_count, cudaStream_t st) {
  if (!conv2_attn_supported(a.N, a.H)) {
    mls_set_error("conv2 attention: unsupported shape (N=%d, H=%d)", a.N, a.H);
    return MLS_ERR_UNSUPPORTED;
  }
  return a.transformer ? launch<true>(a, sm_count, st) : launch<false>(a, sm_count, st);
}

}  // namespace mls
