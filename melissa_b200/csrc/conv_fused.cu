// Fused GATv2 convolution for sm_100a: per tile of G graphs (<= 100 node rows)
//
//   for every head h:
//     TMA  -> smem ring : X tile [128 x K] and the head's weight rows [W_l,h ; W_r,h] (256 x K), 64-wide k blocks
//     tcgen05.mma       : [128 x 256] fp32 accumulator in TMEM (double buffered over heads)
//     16 compute warps  : tcgen05.ld the accumulator, add bias (x decision-maker mask), write x_l | x_r
//                         as fp32 rows into shared memory and accumulate <att_h, x_l>, <att_h, x_r>;
//                         then run the attention of that head out of shared memory (same algorithm as
//                         edge_bf16_kernel) and write relu(conv + bias) / the controlling-node snapshot.
//
// The projections never exist in global memory, there is no staging/convert pass, and the MMAs of head
// h+1 overlap the SIMT attention of head h.  Reference math: PyG GATv2Conv (l_dgn.py:125,133), see
// dgn_forward.cu / dgn_forward_bf16.cu.
#include "conv_fused.cuh"

#include "dgn_kernels.cuh"
#include "gemm_tcgen05.cuh"
#include "tcgen05_ptx.cuh"

namespace mls {

namespace {

constexpr int kFStages = 2;
constexpr int kFA = 128 * 64 * 2;            // 16 KiB  X k-block
constexpr int kFB = 256 * 64 * 2;            // 32 KiB  [W_l,h ; W_r,h] k-block
constexpr int kFStage = kFA + kFB;
constexpr int kComputeWarps = 16;
constexpr int kFThreads = 128 + kComputeWarps * 32;   // 640
constexpr int kPitch = kC + 4;               // fp32 row pitch of the staged x_l / x_r rows (lane = row stores are conflict free)
constexpr int kMaxRowsPerTile = 100;

__device__ __forceinline__ float f_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float f_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ void bar_compute() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

__device__ __forceinline__ void st_bf16x4_g(__nv_bfloat16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}

struct FSmem {      // byte offsets from the 1024-aligned base, computed on host and device identically
  int stA, stT, s_a, s_b, s_dm, s_slot, s_ptr, s_src, bias, att, bars, total;
};
__host__ __device__ inline FSmem fused_layout(int rt, int G, int N) {
  FSmem L;
  int off = kFStages * kFStage;
  L.stA = off; off += rt * kPitch * 4;
  L.stT = off; off += rt * kPitch * 4;
  L.s_a = off; off += 2 * rt * 4;            // double buffered over heads
  L.s_b = off; off += 2 * rt * 4;
  L.s_dm = off; off += rt * 4;
  L.s_slot = off; off += rt * 4;
  L.bias = off; off += 2 * 256 * 4;          // double buffered over heads
  L.att = off; off += 2 * kC * 4;
  L.s_ptr = off; off += ((G * (N + 1) * 2 + 15) / 16) * 16;
  L.s_src = off; off += ((G * N * kMaxNbr + 15) / 16) * 16;
  L.bars = off; off += 128;
  L.total = off + 1024;
  return L;
}

__global__ void __launch_bounds__(kFThreads, 1)
fused_gatv2_conv_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const FusedConvArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // aligned by offset so that the compiler keeps the shared address space (LDS / STS instead of generic LD / ST)
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int N = a.N, H = a.H, HC = H * kC, G = a.G;
  const int rt_max = G * N;
  const FSmem L = fused_layout(rt_max, G, N);
  float* stA = reinterpret_cast<float*>(smem + L.stA);
  float* stT = reinterpret_cast<float*>(smem + L.stT);
  float* s_a = reinterpret_cast<float*>(smem + L.s_a);
  float* s_b = reinterpret_cast<float*>(smem + L.s_b);
  float* s_dm = reinterpret_cast<float*>(smem + L.s_dm);
  int* s_slot = reinterpret_cast<int*>(smem + L.s_slot);
  float* bias_s = reinterpret_cast<float*>(smem + L.bias);
  float* att_s = reinterpret_cast<float*>(smem + L.att);
  uint16_t* s_ptr = reinterpret_cast<uint16_t*>(smem + L.s_ptr);
  uint8_t* s_src = smem + L.s_src;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kFStages + 4);
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (kFStages + s); };
  auto tfull_bar = [&](int b) { return bar0 + 8u * (2 * kFStages + b); };
  auto tempty_bar = [&](int b) { return bar0 + 8u * (2 * kFStages + 2 + b); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (a.n_graphs + G - 1) / G;
  const int n_kb = a.K / 64;

  if (warp == 0 && lane == 0) { prefetch_tmap(&tmX); prefetch_tmap(&tmW); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kFStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), kComputeWarps * 32); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int m0 = tile * G * N;
        for (int h = 0; h < H; ++h) {
          for (int kb = 0; kb < n_kb; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1);
            const uint32_t sa = smem_u32(smem + stage * kFStage);
            mbar_expect_tx(full_bar(stage), kFStage);
            tma_load_2d(sa, &tmX, full_bar(stage), kb * 64, m0);
            tma_load_2d(sa + kFA, &tmW, full_bar(stage), kb * 64, h * kC);                 // W_l rows of head h
            tma_load_2d(sa + kFA + kFB / 2, &tmW, full_bar(stage), kb * 64, HC + h * kC);  // W_r rows of head h
            if (++stage == kFStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(128, 256);
      int stage = 0, it = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int h = 0; h < H; ++h, ++it) {
          const int acc = it & 1;
          mbar_wait(tempty_bar(acc), ((it >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 256);
          for (int kb = 0; kb < n_kb; ++kb) {
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + stage * kFStage);
            const uint64_t da = make_smem_desc(sa), db = make_smem_desc(sa + kFA);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
            umma_commit(empty_bar(stage));
            if (++stage == kFStages) { stage = 0; phase ^= 1; }
          }
          umma_commit(tfull_bar(acc));
        }
      }
    }
  } else if (warp >= 4) {
    // ===================================================================== compute warps
    const int cw = warp - 4;                      // 0..15
    const int quarter = cw & 3, colgrp = cw >> 2; // TMEM lane quarter (tile rows 32q..), 64-column group of the 256
    const int ct = threadIdx.x - 128;             // 0..511
    constexpr float kLog2e = 1.4426950408889634f;
    const int grp = lane >> 3, sub = lane & 7;
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int g0 = tile * G;
      const int gt = min(G, a.n_graphs - g0);
      const int rt = gt * N;
      const size_t m0 = (size_t)g0 * N;
      // per-tile scalars and CSR lists (the previous tile's readers are past their last barrier)
      for (int t = ct; t < rt; t += kComputeWarps * 32) {
        const int gl = t / N, i = t - gl * N;
        s_slot[t] = a.slot ? a.slot[m0 + t] : -1;
        s_dm[t] = a.obs[(long long)(g0 + gl) * a.obs_stride + i * 8 + 7];
        s_a[t] = 0.f; s_b[t] = 0.f;               // buffer 0 (head 0)
      }
      for (int t = ct; t < gt * (N + 1); t += kComputeWarps * 32) s_ptr[t] = a.csr_ptr[(size_t)g0 * (N + 1) + t];
      for (int gl = 0; gl < gt; ++gl) {
        const int E = a.csr_ptr[(size_t)(g0 + gl) * (N + 1) + N];
        const uint8_t* gs = a.csr_src + (size_t)(g0 + gl) * N * kMaxNbr;
        for (int t = ct; t < E; t += kComputeWarps * 32) s_src[gl * N * kMaxNbr + t] = gs[t];
      }
      for (int h = 0; h < H; ++h, ++it) {
        const int acc = it & 1, ab = h & 1;
        // head constants (double buffered: the other buffer may still be read by slow warps of the previous head)
        float* bias_h = bias_s + ab * 256;
        float* att_h = att_s + ab * kC;
        if (ct < 256) bias_h[ct] = a.proj_bias[(ct < 128 ? 0 : HC) + h * kC + (ct & 127)];
        else if (ct < 384) att_h[ct - 256] = a.att[h * kC + (ct - 256)];
        bar_compute();                             // scalars / constants visible; previous head fully consumed
        // ---- accumulator -> shared memory (fp32), + bias, x dm, + <att, row> partial sums
        mbar_wait(tfull_bar(acc), (it >> 1) & 1);
        tc_fence_after();
        {
          const int r = quarter * 32 + lane;       // tile row == TMEM lane
          const bool rv = r < rt;
          const float scale = (rv && a.scale_rows) ? s_dm[r] : 1.0f;
          float* dstrow = (colgrp < 2 ? stA : stT) + r * kPitch + (colgrp & 1) * 64;
          float dot = 0.f;
#pragma unroll 1
          for (int c0 = 0; c0 < 64; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * 256 + colgrp * 64 + c0), v);
            if (rv) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 b4 = *reinterpret_cast<const float4*>(bias_h + colgrp * 64 + c0 + j);
                const float4 a4 = *reinterpret_cast<const float4*>(att_h + (colgrp & 1) * 64 + c0 + j);
                float4 x;
                x.x = fmaf(__uint_as_float(v[j]), scale, b4.x); x.y = fmaf(__uint_as_float(v[j + 1]), scale, b4.y);
                x.z = fmaf(__uint_as_float(v[j + 2]), scale, b4.z); x.w = fmaf(__uint_as_float(v[j + 3]), scale, b4.w);
                dot = fmaf(x.x, a4.x, dot); dot = fmaf(x.y, a4.y, dot); dot = fmaf(x.z, a4.z, dot); dot = fmaf(x.w, a4.w, dot);
                *reinterpret_cast<float4*>(dstrow + c0 + j) = x;
              }
            }
          }
          if (rv) atomicAdd((colgrp < 2 ? s_a : s_b) + ab * rt_max + r, dot * (0.6f * kLog2e));
        }
        tc_fence_before();
        mbar_arrive(tempty_bar(acc));              // the MMA warp may reuse this accumulator
        bar_compute();                             // staged rows + logit scalars complete
        // zero the other scalar buffer for the next head while this head is consumed
        for (int t = ct; t < rt; t += kComputeWarps * 32) { s_a[(ab ^ 1) * rt_max + t] = 0.f; s_b[(ab ^ 1) * rt_max + t] = 0.f; }
        // ---- attention of head h out of shared memory (see edge_bf16_kernel)
        float4 attn[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 v4 = *reinterpret_cast<const float4*>(att_h + (q * 8 + sub) * 4);
          const float sc = 0.4f * kLog2e;
          attn[q] = make_float4(v4.x * sc, v4.y * sc, v4.z * sc, v4.w * sc);
        }
        const float4 bias4 = *reinterpret_cast<const float4*>(a.conv_bias + h * kC + lane * 4);
        const float* sa_h = s_a + ab * rt_max;
        const float* sb_h = s_b + ab * rt_max;
        for (int t = cw; t < rt; t += kComputeWarps) {
          const int sl = s_slot[t];
          if (a.ctrl_only && sl < 0) continue;
          const int gl = t / N, i = t - gl * N;
          const int rbase = gl * N;               // first tile row of this graph
          const uint16_t* ptr = s_ptr + gl * (N + 1);
          const uint8_t* src = s_src + gl * N * kMaxNbr;
          const int r0 = ptr[i];
          const int d = (int)ptr[i + 1] - r0 + 1;  // + self loop (slot 0)
          const float* trow = stT + t * kPitch + sub * 4;
          const float b_i = sb_h[t];
          float mx = -INFINITY, den = 0.f;
          float4 accv[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) accv[q] = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int kb = 0; kb < d; kb += 4) {
            const int k = kb + grp;
            const bool valid = k < d;
            int j = t;
            if (valid && k >= 1) j = rbase + src[r0 + k - 1];
            const float* xrow = stA + j * kPitch + sub * 4;
            float4 x[4];
            float part = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              x[q] = *reinterpret_cast<const float4*>(xrow + q * 32);
              const float4 t4 = *reinterpret_cast<const float4*>(trow + q * 32);
              part = fmaf(attn[q].x, fabsf(x[q].x + t4.x), part); part = fmaf(attn[q].y, fabsf(x[q].y + t4.y), part);
              part = fmaf(attn[q].z, fabsf(x[q].z + t4.z), part); part = fmaf(attn[q].w, fabsf(x[q].w + t4.w), part);
            }
            part += __shfl_xor_sync(0xffffffffu, part, 1);
            part += __shfl_xor_sync(0xffffffffu, part, 2);
            part += __shfl_xor_sync(0xffffffffu, part, 4);
            float e = part + (sa_h[j] + b_i);
            if (!valid) e = -INFINITY;
            float m_r = fmaxf(e, __shfl_xor_sync(0xffffffffu, e, 8));
            m_r = fmaxf(m_r, __shfl_xor_sync(0xffffffffu, m_r, 16));
            if (m_r > mx) {
              const float resc = f_ex2(mx - m_r);
              den *= resc;
#pragma unroll
              for (int q = 0; q < 4; ++q) { accv[q].x *= resc; accv[q].y *= resc; accv[q].z *= resc; accv[q].w *= resc; }
              mx = m_r;
            }
            const float p = f_ex2(e - mx);
            den += p;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              accv[q].x = fmaf(p, x[q].x, accv[q].x); accv[q].y = fmaf(p, x[q].y, accv[q].y);
              accv[q].z = fmaf(p, x[q].z, accv[q].z); accv[q].w = fmaf(p, x[q].w, accv[q].w);
            }
          }
          den += __shfl_xor_sync(0xffffffffu, den, 8);
          den += __shfl_xor_sync(0xffffffffu, den, 16);
          float4 lo, hi;
          {
            const bool up = grp >= 2;
            const float4 k0 = up ? accv[2] : accv[0], k1 = up ? accv[3] : accv[1];
            const float4 s0 = up ? accv[0] : accv[2], s1 = up ? accv[1] : accv[3];
            lo.x = k0.x + __shfl_xor_sync(0xffffffffu, s0.x, 16); lo.y = k0.y + __shfl_xor_sync(0xffffffffu, s0.y, 16);
            lo.z = k0.z + __shfl_xor_sync(0xffffffffu, s0.z, 16); lo.w = k0.w + __shfl_xor_sync(0xffffffffu, s0.w, 16);
            hi.x = k1.x + __shfl_xor_sync(0xffffffffu, s1.x, 16); hi.y = k1.y + __shfl_xor_sync(0xffffffffu, s1.y, 16);
            hi.z = k1.z + __shfl_xor_sync(0xffffffffu, s1.z, 16); hi.w = k1.w + __shfl_xor_sync(0xffffffffu, s1.w, 16);
          }
          float4 mine;
          {
            const bool odd = grp & 1;
            const float4 kp = odd ? hi : lo, sd = odd ? lo : hi;
            mine.x = kp.x + __shfl_xor_sync(0xffffffffu, sd.x, 8); mine.y = kp.y + __shfl_xor_sync(0xffffffffu, sd.y, 8);
            mine.z = kp.z + __shfl_xor_sync(0xffffffffu, sd.z, 8); mine.w = kp.w + __shfl_xor_sync(0xffffffffu, sd.w, 8);
          }
          const float inv_den = f_rcp(den + 1e-16f);
          float4 o;
          o.x = fmaxf(fmaf(mine.x, inv_den, bias4.x), 0.f); o.y = fmaxf(fmaf(mine.y, inv_den, bias4.y), 0.f);
          o.z = fmaxf(fmaf(mine.z, inv_den, bias4.z), 0.f); o.w = fmaxf(fmaf(mine.w, inv_den, bias4.w), 0.f);
          if (a.x_out) st_bf16x4_g(a.x_out + (m0 + t) * HC + h * kC + lane * 4, o);
          if (a.z && sl >= 0) st_bf16x4_g(a.z + (size_t)sl * a.ldz + a.z_col + h * kC + lane * 4, o);
        }
        // the barrier at the top of the next head (or tile) separates these reads from the next overwrite
      }
      bar_compute();                               // all warps done with this tile's scalars / lists
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

}  // namespace

int fused_gatv2_conv_launch(const __nv_bfloat16* X, const __nv_bfloat16* Wt, const FusedConvArgs& a, int sm_count, cudaStream_t st) {
  if (a.G < 1 || a.G * a.N > kMaxRowsPerTile || a.K % 64 != 0 || a.H * kC * 2 > 65536) {
    mls_set_error("fused GATv2 conv: unsupported shape (N=%d, G=%d, K=%d)", a.N, a.G, a.K);
    return MLS_ERR_UNSUPPORTED;
  }
  const FSmem L = fused_layout(a.G * a.N, a.G, a.N);
  if (L.total > 227 * 1024) {
    mls_set_error("fused GATv2 conv needs %d bytes of shared memory", L.total);
    return MLS_ERR_UNSUPPORTED;
  }
  static int configured = 0;
  if (L.total > configured) {
    MLS_CUDA(cudaFuncSetAttribute(fused_gatv2_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
    configured = L.total;
  }
  CUtensorMap tx, tw;
  int rc = make_tmap_bf16(&tx, X, a.rows, a.K, a.K, 128);
  if (rc) return rc;
  rc = make_tmap_bf16(&tw, Wt, 2 * a.H * kC, a.K, a.K, 128);
  if (rc) return rc;
  const int n_tiles = (a.n_graphs + a.G - 1) / a.G;
  const int grid = n_tiles < sm_count ? n_tiles : sm_count;
  if (grid <= 0) return MLS_OK;
  fused_gatv2_conv_kernel<<<grid, kFThreads, L.total, st>>>(tx, tw, a);
  mls_count_launch();
  MLS_LAUNCH_CHECK();
  return MLS_OK;
}

}  // namespace mls
