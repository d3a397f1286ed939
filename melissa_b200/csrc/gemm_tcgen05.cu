// tcgen05 / TMEM / TMA bf16 GEMM for the dense projections of the DGN networks (sm_100a).
// See gemm_tcgen05.cuh for the contract.  Replaces the cuBLAS sgemm calls behind
// torch.nn.Linear in PyG GATv2Conv / TransformerConv and tianshou MLP
// (reference l_dgn.py:125,133,142-149; dgn_r.py:105,113; hl_dgn.py:101,111-117).
#include "gemm_tcgen05.cuh"

#include <cuda_fp16.h>

#include "tcgen05_ptx.cuh"

namespace mls {

template <int BN>
struct GemmSmem {
  static constexpr int kNumStages = BN == 256 ? 3 : 4;
  static constexpr int kABytes = kBM * kBK * 2;      // 16 KiB
  static constexpr int kBBytes = BN * kBK * 2;       // 16 / 32 KiB
  static constexpr int kStageBytes = kABytes + kBBytes;
  // epilogue staging: 16 warps, each 32 rows x BN/4 bf16 (+16 B pad: conflict-free 16 B stores by row-owning lanes)
  static constexpr int kOutRowBytes = BN / 2 + 16;
  static constexpr int kOutWarpBytes = 32 * kOutRowBytes;
  static constexpr int kOutOffset = kNumStages * kStageBytes;
  // bias / dot vectors of a tile's column slice, double buffered: the next tile's slices arrive by cp.async while this
  // tile's accumulator is drained (the epilogue, not the MMA, bounds the tile rate: ncu showed the MMA warp waiting for
  // a free accumulator and the epilogue warps exposed to three dependent global-load latencies per tile)
  static constexpr int kBiasOffset = kOutOffset + kGemmEpiWarps * kOutWarpBytes;      // 2 x BN floats
  static constexpr int kDotOffset = kBiasOffset + 2 * BN * 4;             // 2 x BN floats
  static constexpr int kDot2Offset = kDotOffset + 2 * BN * 4;             // 2 x BN floats
  static constexpr int kBarOffset = kDot2Offset + 2 * BN * 4;
  static constexpr int kTotal = kBarOffset + 256 + 1024;   // barriers + alignment slack
};

// MODE fixes the epilogue options at compile time for the hot launches (0 = read them from `epi` at run time):
//   1 = one dot vector on the raw values, C written (GATv2 projections)   2 = ReLU, C written, no dots (hidden layers)
//   3 = ReLU + two dot vectors on the rectified values, C not written (last hidden layer with the output layer fused)
// The epilogue bounds the tile rate, and the run-time selects were a fifth of its instructions.
template <int BN, int MODE>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const __grid_constant__ CUtensorMap tmA2, const GemmShape shape, const GemmEpilogue epi) {
  using S = GemmSmem<BN>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment
  // aligned by offset so that the compiler keeps the shared address space (LDS / STS instead of generic LD / ST)
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kBarOffset);
  // bars[0..S) full, [S..2S) empty, [2S..2S+2) tmem_full, [2S+2..2S+4) tmem_empty
  constexpr int kStages = S::kNumStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (kStages + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * kStages + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * kStages + 2 + a); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int M = shape.M;
  if (shape.m_dev) M = min(M, *shape.m_dev);
  const int n_m = (M + kBM - 1) / kBM, n_n = shape.N / BN;
  const int n_tiles = n_m * n_n;
  const int n_kb = shape.K / kBK;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), kGemmEpiWarps * 32); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(tmem_slot), 2 * BN);      // 2 accumulator stages of BN fp32 columns
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
        const int m0 = (tile / n_n) * kBM, n0 = (tile % n_n) * BN;
        for (int kb = 0; kb < n_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_u32(smem + stage * S::kStageBytes);
          mbar_expect_tx(full_bar(stage), S::kStageBytes);
          if (kb * kBK >= shape.k2_lo && kb * kBK < shape.k2_hi) tma_load_2d(sa, &tmA2, full_bar(stage), kb * kBK - shape.k2_lo, m0);
          else tma_load_2d(sa, &tmA, full_bar(stage), kb * kBK, m0);
          tma_load_2d(sa + S::kABytes, &tmB, full_bar(stage), kb * kBK, n0);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (one thread)
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(kBM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);          // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < n_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);                // TMA bytes have landed
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * S::kStageBytes);
          const uint64_t da = make_smem_desc(sa), db = make_smem_desc(sa + S::kABytes);
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k)            // +32 B per K=16 step inside the swizzle atom
            umma_bf16(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
          umma_commit(empty_bar(stage));                    // frees the smem slot when these MMAs retire
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull_bar(acc));                        // accumulator complete -> epilogue
      }
    }
  } else if (warp >= 4) {
    // ===================================================================== epilogue
    // TMEM -> registers (lane = row) -> scale/bias/ReLU -> bf16 -> per-warp smem tile -> coalesced row stores
    // sixteen warps (the epilogue, not the MMA, bounds the tile rate with fewer): warp % 4 selects the TMEM lane quarter
    // (rows) a warp may read, (warp - 4) / 4 the column quarter
    const int ew = (warp - 4) & 3, half = (warp - 4) >> 2;
    const bool has_dot = MODE == 0 ? epi.dotvec != nullptr : (MODE == 1 || MODE == 3);
    const bool has_dot2 = MODE == 0 ? epi.dotvec2 != nullptr : MODE == 3;
    const bool dot_relu = MODE == 0 ? epi.dot_relu != 0 : MODE == 3;
    const bool relu = MODE == 0 ? epi.relu != 0 : (MODE == 2 || MODE == 3);
    const bool has_cf = MODE == 0 ? epi.Cf != nullptr : false;
    const bool has_c = MODE == 0 ? epi.C != nullptr : MODE != 3;
    constexpr int HN = BN / 4;
    const int cb = half * HN;
    unsigned char* stage_out = smem + S::kOutOffset + (warp - 4) * S::kOutWarpBytes;
    float* bias_s = reinterpret_cast<float*>(smem + S::kBiasOffset);
    float* dot_s = reinterpret_cast<float*>(smem + S::kDotOffset);
    float* dot2_s = reinterpret_cast<float*>(smem + S::kDot2Offset);
    const int et = threadIdx.x - 128;                       // 0..511 within the epilogue warps
    auto cp4 = [](float* dst, const float* src) {
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
    };
    // column slices (bias, dot vectors) of `tile` -> buffer `buf`, asynchronously
    auto fill = [&](int buf, int tile) {
      if (tile >= n_tiles) return;
      const int fn0 = (tile % n_n) * BN;
      for (int c = et; c < BN; c += kGemmEpiWarps * 32) {
        if (epi.bias) cp4(bias_s + buf * BN + c, epi.bias + fn0 + c);
        else bias_s[buf * BN + c] = 0.0f;
        if (has_dot) cp4(dot_s + buf * BN + c, epi.dotvec + fn0 + c);
        if (has_dot2) cp4(dot2_s + buf * BN + c, epi.dotvec2 + fn0 + c);
      }
    };
    // row scale of this thread's row in `tile`: node row index (one tile ahead), then the observation's dm flag
    auto ld_index = [&](int tile) -> int {
      if (!epi.obs || tile >= n_tiles) return -1;
      const int rr = (tile / n_n) * kBM + ew * 32 + lane;
      if (rr >= M) return -1;
      return epi.row_index ? __ldg(epi.row_index + rr) : rr;
    };
    auto ld_scale = [&](int nr) -> float {
      if (nr < 0) return 1.0f;
      const int g = nr / epi.nodes, i = nr - g * epi.nodes;
      return __ldg(epi.obs + (long long)g * epi.obs_stride + i * 8 + 7);
    };
    fill(0, blockIdx.x);
    int nr_next = ld_index(blockIdx.x);
    float scale_cur = ld_scale(nr_next);
    nr_next = ld_index(blockIdx.x + gridDim.x);
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int m0 = (tile / n_n) * kBM, n0 = (tile % n_n) * BN;
      const int r = m0 + ew * 32 + lane;
      const float scale = scale_cur;
      // software pipeline: the next tile's scale (its index arrived during the previous tile) and the index after that
      const float scale_nx = ld_scale(nr_next);
      const int nr_nx2 = ld_index(tile + 2 * gridDim.x);
      const float* bias_b = bias_s + (it & 1) * BN;
      const float* dot_b = dot_s + (it & 1) * BN;
      const float* dot2_b = dot2_s + (it & 1) * BN;
      asm volatile("cp.async.wait_all;" ::: "memory");      // this thread's part of this tile's slices has landed
      asm volatile("bar.sync 1, 512;" ::: "memory");        // ... everybody's has, and the previous tile's readers are done
      fill((it & 1) ^ 1, tile + gridDim.x);
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      float dot = 0.f, dot2 = 0.f;
#pragma unroll 1
      for (int c0 = cb; c0 < cb + HN; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * BN + c0), v);
        uint32_t packed[16];
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(bias_b + c0 + j);
          float x0 = fmaf(__uint_as_float(v[j]), scale, b4.x), x1 = fmaf(__uint_as_float(v[j + 1]), scale, b4.y);
          float x2 = fmaf(__uint_as_float(v[j + 2]), scale, b4.z), x3 = fmaf(__uint_as_float(v[j + 3]), scale, b4.w);
          if (has_dot) {
            const float4 d4 = *reinterpret_cast<const float4*>(dot_b + c0 + j);
            const float y0 = dot_relu ? fmaxf(x0, 0.f) : x0, y1 = dot_relu ? fmaxf(x1, 0.f) : x1;
            const float y2 = dot_relu ? fmaxf(x2, 0.f) : x2, y3 = dot_relu ? fmaxf(x3, 0.f) : x3;
            dot = fmaf(y0, d4.x, dot); dot = fmaf(y1, d4.y, dot); dot = fmaf(y2, d4.z, dot); dot = fmaf(y3, d4.w, dot);
            if (has_dot2) {
              const float4 e4 = *reinterpret_cast<const float4*>(dot2_b + c0 + j);
              dot2 = fmaf(y0, e4.x, dot2); dot2 = fmaf(y1, e4.y, dot2); dot2 = fmaf(y2, e4.z, dot2); dot2 = fmaf(y3, e4.w, dot2);
            }
          }
          if (relu) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); x2 = fmaxf(x2, 0.f); x3 = fmaxf(x3, 0.f); }
          if (has_cf && r < M)                                // lane = row: 16-byte stores, 128 contiguous bytes per lane and chunk
            *reinterpret_cast<float4*>(epi.Cf + (size_t)r * epi.ldcf + n0 + c0 + j) = make_float4(x0, x1, x2, x3);
          if (!has_c) {
          } else if (epi.c_fp16) {
            asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(packed[j >> 1]) : "f"(x1), "f"(x0));
            asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(packed[(j >> 1) + 1]) : "f"(x3), "f"(x2));
          } else {
            __nv_bfloat162 p0 = __floats2bfloat162_rn(x0, x1), p1 = __floats2bfloat162_rn(x2, x3);
            packed[j >> 1] = *reinterpret_cast<uint32_t*>(&p0);
            packed[(j >> 1) + 1] = *reinterpret_cast<uint32_t*>(&p1);
          }
        }
        if (has_c) {
          uint4* dst = reinterpret_cast<uint4*>(stage_out + lane * S::kOutRowBytes + (c0 - cb) * 2);
#pragma unroll
          for (int q = 0; q < 4; ++q) dst[q] = make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
        }
        if (has_dot && (c0 & 32) == 32) {                 // end of this warp's half (64 columns) of a 128-column group
          // the group's two halves come from two warps: two atomic adds onto a zeroed cell (launcher) -- with exactly two
          // terms the sum does not depend on their order (0 + x + y == 0 + y + x), so the result is deterministic
          if (r < M) {
            atomicAdd(epi.dots + (size_t)r * (shape.N >> 7) + ((n0 + c0) >> 7), dot);
            if (has_dot2) atomicAdd(epi.dots2 + (size_t)r * (shape.N >> 7) + ((n0 + c0) >> 7), dot2);
          }
          dot = 0.f; dot2 = 0.f;
        }
      }
      // accumulator drained: hand it back to the MMA warp before the global stores
      tc_fence_before();
      mbar_arrive(tempty_bar(acc));
      __syncwarp();
      // each instruction now writes whole row segments of BN/2 bf16 (BN bytes contiguous each)
      constexpr int kLanesPerRow = HN * 2 / 16;             // 8 (BN=128) or 16 (BN=256) lanes cover one row segment
      constexpr int kRowsPerIter = 32 / kLanesPerRow;
      const int sub = lane / kLanesPerRow, cl = lane % kLanesPerRow;
      if (has_c) {
        const int gr0 = m0 + ew * 32 + sub;
        __nv_bfloat16* crow = epi.C + (size_t)gr0 * epi.ldc + n0 + cb + cl * 8;
        const size_t cstep = (size_t)kRowsPerIter * epi.ldc;
        const unsigned char* srow = stage_out + sub * S::kOutRowBytes + cl * 16;
#pragma unroll 4
        for (int rr = 0; rr < 32; rr += kRowsPerIter) {
          const uint4 val = *reinterpret_cast<const uint4*>(srow + rr * S::kOutRowBytes);
          if (gr0 + rr < M) *reinterpret_cast<uint4*>(crow) = val;
          crow += cstep;
        }
      }
      __syncwarp();
      scale_cur = scale_nx;
      nr_next = nr_nx2;
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 2 * BN);
}

// ------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_tmap_bf16(CUtensorMap* tm, const void* base, int rows, int cols, int ld_elems, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { mls_set_error("cuTensorMapEncodeTiled not available from the driver"); return MLS_ERR_CUDA; }
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld_elems * 2};
  cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { mls_set_error("cuTensorMapEncodeTiled failed with %d (rows=%d cols=%d ld=%d)", (int)r, rows, cols, ld_elems); return MLS_ERR_CUDA; }
  return MLS_OK;
}

size_t gemm_smem_bytes(int BN) { return BN == 256 ? GemmSmem<256>::kTotal : GemmSmem<128>::kTotal; }

template <int BN, int MODE>
static int launch_bn(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& ta2, GemmShape shape, GemmEpilogue epi, int sm_count, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    MLS_CUDA(cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmSmem<BN>::kTotal));
    configured = true;
  }
  const int n_tiles = ((shape.M + kBM - 1) / kBM) * (shape.N / BN);
  const int grid = n_tiles < sm_count ? n_tiles : sm_count;
  if (grid <= 0) return MLS_OK;
  gemm_bf16_tcgen05_kernel<BN, MODE><<<grid, kGemmThreads, GemmSmem<BN>::kTotal, st>>>(ta, tb, ta2, shape, epi);
  mls_count_launch();
  MLS_LAUNCH_CHECK();
  return MLS_OK;
}

int gemm_bf16_launch(const __nv_bfloat16* A, int lda, const __nv_bfloat16* B, int ldb, GemmShape shape, GemmEpilogue epi,
                     int sm_count, cudaStream_t st) {
  MLS_CHECK_ARG(shape.K % kBK == 0 && shape.K >= kBK, "GEMM K must be a multiple of %d (got %d)", kBK, shape.K);
  MLS_CHECK_ARG(shape.N % 128 == 0, "GEMM N must be a multiple of 128 (got %d)", shape.N);
  MLS_CHECK_ARG(lda % 8 == 0 && ldb % 8 == 0 && (!epi.C || epi.ldc % 8 == 0), "GEMM leading dimensions must be multiples of 8 elements");
  MLS_CHECK_ARG(!epi.Cf || (epi.ldcf % 4 == 0 && (reinterpret_cast<uintptr_t>(epi.Cf) & 15) == 0), "fp32 GEMM output must be 16-byte aligned");
  MLS_CHECK_ARG(!epi.dotvec || shape.N % 256 == 0, "the fused per-group dot products need N to be a multiple of 256");
  if (shape.M <= 0) return MLS_OK;
  if (epi.dotvec) {                                         // the epilogue accumulates two partial sums per cell
    MLS_CUDA(cudaMemsetAsync(epi.dots, 0, (size_t)shape.M * (shape.N >> 7) * sizeof(float), st));
    if (epi.dotvec2) MLS_CUDA(cudaMemsetAsync(epi.dots2, 0, (size_t)shape.M * (shape.N >> 7) * sizeof(float), st));
  }
  const int BN = (shape.N % 256 == 0) ? 256 : 128;
  CUtensorMap ta, tb;
  int rc = make_tmap_bf16(&ta, A, shape.M, shape.K, lda, kBM);
  if (rc) return rc;
  rc = make_tmap_bf16(&tb, B, shape.N, shape.K, ldb, BN);
  if (rc) return rc;
  CUtensorMap ta2 = ta;
  if (shape.A2) {
    MLS_CHECK_ARG(shape.k2_lo % kBK == 0 && shape.k2_hi % kBK == 0 && shape.k2_lo >= 0 && shape.k2_lo < shape.k2_hi && shape.k2_hi <= shape.K && shape.lda2 % 8 == 0,
                  "GEMM second A matrix: K range must be multiples of %d inside [0, K)", kBK);
    rc = make_tmap_bf16(&ta2, shape.A2, shape.M, shape.k2_hi - shape.k2_lo, shape.lda2, kBM);
    if (rc) return rc;
  } else {
    shape.k2_lo = shape.k2_hi = 0;
  }
  int mode = 0;
  if (!epi.Cf) {
    if (epi.C && epi.dotvec && !epi.dotvec2 && !epi.dot_relu && !epi.relu) mode = 1;
    else if (epi.C && !epi.dotvec && epi.relu) mode = 2;
    else if (!epi.C && epi.dotvec && epi.dotvec2 && epi.dot_relu && epi.relu) mode = 3;
  }
  if (BN == 256) {
    switch (mode) {
      case 1: return launch_bn<256, 1>(ta, tb, ta2, shape, epi, sm_count, st);
      case 2: return launch_bn<256, 2>(ta, tb, ta2, shape, epi, sm_count, st);
      case 3: return launch_bn<256, 3>(ta, tb, ta2, shape, epi, sm_count, st);
      default: return launch_bn<256, 0>(ta, tb, ta2, shape, epi, sm_count, st);
    }
  }
  return launch_bn<128, 0>(ta, tb, ta2, shape, epi, sm_count, st);
}

}  // namespace mls

// Standalone entry for the GEMM unit test (tests/test_gemm_gpu.py); not part of the reference-facing ABI.
extern "C" int mls_test_gemm_bf16(const void* A, const void* B, const float* bias, const float* obs, long long obs_stride,
                                  int nodes, void* C, int M, int N, int K, int relu, const int* m_dev, void* stream) {
  int dev = 0, sms = 0;
  MLS_CUDA(cudaGetDevice(&dev));
  MLS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  mls::GemmShape shape{M, N, K, m_dev};
  mls::GemmEpilogue epi{reinterpret_cast<__nv_bfloat16*>(C), N, bias, obs, obs_stride, nodes, relu, nullptr, nullptr};
  return mls::gemm_bf16_launch(reinterpret_cast<const __nv_bfloat16*>(A), K, reinterpret_cast<const __nv_bfloat16*>(B), K, shape,
                               epi, sms, reinterpret_cast<cudaStream_t>(stream));
}

// Same, with every epilogue option (tests/test_gemm_gpu.py): compacted row scale, one or two dot vectors per
// 128-column group (optionally on max(value, 0)), C optional.
extern "C" int mls_test_gemm_bf16_ex(const void* A, const void* B, const float* bias, const float* obs, long long obs_stride, int nodes,
                                     const int* row_index, const float* dotvec, const float* dotvec2, float* dots, float* dots2,
                                     int dot_relu, void* C, int M, int N, int K, int relu, const int* m_dev, void* stream) {
  int dev = 0, sms = 0;
  MLS_CUDA(cudaGetDevice(&dev));
  MLS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  mls::GemmShape shape{M, N, K, m_dev};
  mls::GemmEpilogue epi{reinterpret_cast<__nv_bfloat16*>(C), N, bias, obs, obs_stride, nodes, relu, dotvec, dots, row_index, dotvec2, dots2,
                        dot_relu};
  return mls::gemm_bf16_launch(reinterpret_cast<const __nv_bfloat16*>(A), K, reinterpret_cast<const __nv_bfloat16*>(B), K, shape,
                               epi, sms, reinterpret_cast<cudaStream_t>(stream));
}
