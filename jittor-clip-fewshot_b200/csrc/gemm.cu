// Persistent, warp-specialised bf16 GEMM for sm_100a:   C[M,N] = A[M,K] * B[N,K]^T  (+ fused epilogue)
//
//   warp 0      TMA producer   : cp.async.bulk.tensor 2-D tiles (128B swizzle) into a STAGES-deep smem ring
//   warp 1      MMA issuer     : one elected thread issues tcgen05.mma (UMMA 128 x BN x 16, bf16 -> fp32)
//                                accumulating in TMEM; tcgen05.commit releases smem slots / publishes tiles
//   warps 2..5  epilogue       : tcgen05.ld TMEM -> registers, bias / QuickGELU / residual / pos-embed,
//                                vectorised global stores; TMEM accumulators are double buffered so the
//                                epilogue of tile i overlaps the main loop of tile i+1
//
// This replaces the reference's fp32 `nn.Linear` calls on the hot path: packed QKV projection
// (jclip/mha.py:129-146, test.py:557-559), attention out-proj (jclip/mha.py:461, test.py:594), MLP
// c_fc / c_proj (jclip/model.py:38-39) and the patch-embed conv lowered to an im2col GEMM
// (jclip/model.py:105-108).  Residual adds (jclip/model.py:60-61), QuickGELU (jclip/model.py:27) and the
// positional-embedding add (jclip/model.py:114) are fused into the epilogues.
#include <cstdio>
#include <cudaTypedefs.h>

#include "kernels.h"
#include "ptx.cuh"

namespace jcb {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;   // 64 bf16 = 128 B = one swizzle atom row
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 192;
constexpr int EPI_WARPS = 4;
constexpr int SMEM_BUDGET = 196608;  // operand ring; barriers + bias tile come on top

template <int BN>
struct TileCfg {
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_BYTES = BN * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = SMEM_BUDGET / STAGE_BYTES;
  static constexpr int TMEM_COLS = 2 * BN;  // two accumulator stages; power of two for BN in {128,256}
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 4 * BN * 4 + 256 + 1024;  // + per-warp bias + barriers + align slack
};

struct GemmDev {
  int M, N, K;
  const float* bias;
  void* out;
  long long ldo;
  const float* pos;
  int tokens_in, tokens_out;
  int* status;
};

__device__ __forceinline__ float quick_gelu(float x) {
  // x * sigmoid(1.702 x)  (reference jclip/model.py:27)
  return __fdividef(x, 1.0f + __expf(-1.702f * x));
}

template <int BN, int EPI>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const GemmDev p) {
  using Cfg = TileCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  // 128B-swizzled TMA/UMMA tiles need 1024-byte alignment in the shared address space.
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* ring = smem;
  float* s_bias_all = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES + 4 * BN * 4);
  uint64_t* full_bar = bars;                   // [STAGES]  TMA -> MMA
  uint64_t* empty_bar = bars + STAGES;         // [STAGES]  MMA -> TMA
  uint64_t* tmem_full_bar = bars + 2 * STAGES;      // [2]  MMA -> epilogue
  uint64_t* tmem_empty_bar = bars + 2 * STAGES + 2; // [2]  epilogue -> MMA
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp_idx = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  const int m_tiles = (p.M + BLOCK_M - 1) / BLOCK_M;
  const int n_tiles = p.N / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int num_kb = p.K / BLOCK_K;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], EPI_WARPS);
    }
    fence_mbar_init();
    fence_proxy_async_smem();
  }
  if (warp_idx == 1) {
    tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_ptr_smem);

  if (warp_idx == 0) {
    // ===================================================================== TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      bool ok = true;
      for (int tile = blockIdx.x; tile < num_tiles && ok; tile += gridDim.x) {
        const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
        for (int kb = 0; kb < num_kb; ++kb) {
          if (!mbar_wait(&empty_bar[stage], phase ^ 1u, p.status, JCB_DEV_TIMEOUT_PRODUCER)) { ok = false; break; }
          uint8_t* sa = ring + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          tma_load_2d(sa, &tmA, &full_bar[stage], kb * BLOCK_K, m_blk * BLOCK_M);
          tma_load_2d(sb, &tmB, &full_bar[stage], kb * BLOCK_K, n_blk * BN);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ===================================================================== MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16_f32(BLOCK_M, BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      bool ok = true;
      for (int tile = blockIdx.x; tile < num_tiles && ok; tile += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        if (!mbar_wait(&tmem_empty_bar[as], aphase ^ 1u, p.status, JCB_DEV_TIMEOUT_MMA)) { ok = false; break; }
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          if (!mbar_wait(&full_bar[stage], phase, p.status, JCB_DEV_TIMEOUT_MMA)) { ok = false; break; }
          tc_fence_after();
          const uint32_t sa = smem_u32(ring + stage * Cfg::STAGE_BYTES);
          const uint64_t da = umma_desc_sw128(sa);
          const uint64_t db = umma_desc_sw128(sa + Cfg::A_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // +32 B per UMMA_K step inside the 128-B swizzle atom = +2 in the (addr >> 4) field
            umma_bf16(tmem_d, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc,
                      (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        if (ok) umma_commit(&tmem_full_bar[as]);  // accumulator complete -> epilogue
      }
    }
  } else {
    // ===================================================================== epilogue (warps 2..5)
    const int q = warp_idx & 3;  // TMEM lane quarter this warp may access
    // Each epilogue warp keeps a private copy of the tile's bias slice: no cross-warp barrier in
    // the epilogue, so a warp that abandons its loop on a pipeline error cannot strand the others.
    float* s_bias = s_bias_all + q * BN;
    int it = 0;
    bool ok = true;
    for (int tile = blockIdx.x; tile < num_tiles && ok; tile += gridDim.x, ++it) {
      const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      __syncwarp();
      for (int i = lane; i < BN; i += 32) s_bias[i] = p.bias ? __ldg(p.bias + n_blk * BN + i) : 0.0f;
      __syncwarp();

      ok = mbar_wait(&tmem_full_bar[as], aphase, p.status, JCB_DEV_TIMEOUT_EPILOGUE);
      ok = __all_sync(0xffffffffu, ok);
      if (!ok) break;
      tc_fence_after();

      const int row = m_blk * BLOCK_M + q * 32 + lane;
      const bool valid = row < p.M;
      long long orow = row;
      int tok = 0;
      if (EPI == EPI_PATCH_F32) {
        tok = row % p.tokens_in + 1;
        orow = static_cast<long long>(row / p.tokens_in) * p.tokens_out + tok;
      }
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as * BN);
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + static_cast<uint32_t>(c * 32), v);
        const int n0 = n_blk * BN + c * 32;
        if (EPI == EPI_BIAS_RESID_F32) {
          // prefetch the residual row segment while the TMEM load is in flight
          float4 r[8];
          float4* dst = reinterpret_cast<float4*>(static_cast<float*>(p.out) + orow * p.ldo + n0);
          if (valid) {
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = dst[j];
          }
          tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b = *reinterpret_cast<const float4*>(s_bias + c * 32 + 4 * j);
              r[j].x += __uint_as_float(v[4 * j + 0]) + b.x;
              r[j].y += __uint_as_float(v[4 * j + 1]) + b.y;
              r[j].z += __uint_as_float(v[4 * j + 2]) + b.z;
              r[j].w += __uint_as_float(v[4 * j + 3]) + b.w;
              dst[j] = r[j];
            }
          }
        } else if (EPI == EPI_PATCH_F32 || EPI == EPI_F32) {
          tmem_ld_wait();
          if (valid) {
            float4* dst = reinterpret_cast<float4*>(static_cast<float*>(p.out) + orow * p.ldo + n0);
            const float4* pe = reinterpret_cast<const float4*>(
                EPI == EPI_PATCH_F32 ? p.pos + static_cast<long long>(tok) * p.N + n0 : nullptr);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b = *reinterpret_cast<const float4*>(s_bias + c * 32 + 4 * j);
              float4 o;
              o.x = __uint_as_float(v[4 * j + 0]) + b.x;
              o.y = __uint_as_float(v[4 * j + 1]) + b.y;
              o.z = __uint_as_float(v[4 * j + 2]) + b.z;
              o.w = __uint_as_float(v[4 * j + 3]) + b.w;
              if (EPI == EPI_PATCH_F32) {
                const float4 e = __ldg(pe + j);
                o.x += e.x; o.y += e.y; o.z += e.z; o.w += e.w;
              }
              dst[j] = o;
            }
          }
        } else {
          tmem_ld_wait();
          if (valid) {
            uint4* dst = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + orow * p.ldo + n0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float f[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                f[e] = __uint_as_float(v[8 * j + e]) + s_bias[c * 32 + 8 * j + e];
                if (EPI == EPI_BIAS_GELU_BF16) f[e] = quick_gelu(f[e]);
              }
              uint4 o;
              o.x = pack_bf16x2(f[0], f[1]);
              o.y = pack_bf16x2(f[2], f[3]);
              o.z = pack_bf16x2(f[4], f[5]);
              o.w = pack_bf16x2(f[6], f[7]);
              dst[j] = o;
            }
          }
        }
      }
      // all TMEM reads of this warp are complete (wait::ld above): hand the accumulator stage back
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[as]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

PFN_cuTensorMapEncodeTiled_v12000 g_encode_tiled = nullptr;
char g_driver_err[256] = {0};

bool make_tmap_bf16_2d(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                       uint32_t box_rows, uint32_t box_cols) {
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld_elems * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box,
                              estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

template <int BN, int EPI>
cudaError_t launch_cfg(const GemmArgs& a, int* dev_status, int num_sms, cudaStream_t stream) {
  using Cfg = TileCfg<BN>;
  CUtensorMap tmA, tmB;
  if (!make_tmap_bf16_2d(&tmA, a.A, a.M, a.K, a.lda, BLOCK_M, BLOCK_K)) return cudaErrorInvalidValue;
  if (!make_tmap_bf16_2d(&tmB, a.B, a.N, a.K, a.ldb, BN, BLOCK_K)) return cudaErrorInvalidValue;
  auto kern = gemm_bf16_tcgen05_kernel<BN, EPI>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  GemmDev p;
  p.M = a.M; p.N = a.N; p.K = a.K;
  p.bias = a.bias; p.out = a.out; p.ldo = a.ldo; p.pos = a.pos;
  p.tokens_in = a.tokens_in; p.tokens_out = a.tokens_out; p.status = dev_status;
  const int m_tiles = (a.M + BLOCK_M - 1) / BLOCK_M;
  const int tiles = m_tiles * (a.N / BN);
  const int grid = tiles < num_sms ? tiles : num_sms;
  kern<<<grid, NUM_THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, p);
  return cudaGetLastError();
}

template <int BN>
cudaError_t launch_bn(const GemmArgs& a, int* st, int sms, cudaStream_t s) {
  switch (a.epilogue) {
    case EPI_BIAS_BF16: return launch_cfg<BN, EPI_BIAS_BF16>(a, st, sms, s);
    case EPI_BIAS_GELU_BF16: return launch_cfg<BN, EPI_BIAS_GELU_BF16>(a, st, sms, s);
    case EPI_BIAS_RESID_F32: return launch_cfg<BN, EPI_BIAS_RESID_F32>(a, st, sms, s);
    case EPI_PATCH_F32: return launch_cfg<BN, EPI_PATCH_F32>(a, st, sms, s);
    case EPI_F32: return launch_cfg<BN, EPI_F32>(a, st, sms, s);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace

const char* gemm_init_driver_api() {
  if (g_encode_tiled) return nullptr;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || fn == nullptr) {
    snprintf(g_driver_err, sizeof(g_driver_err), "cuTensorMapEncodeTiled not available: %s",
             cudaGetErrorString(e));
    return g_driver_err;
  }
  g_encode_tiled = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  return nullptr;
}

cudaError_t launch_gemm(const GemmArgs& a, int* dev_status, int num_sms, cudaStream_t stream) {
  if (!g_encode_tiled) return cudaErrorNotReady;
  if (a.M <= 0 || a.N <= 0 || a.K <= 0 || a.K % BLOCK_K != 0 || a.N % 128 != 0) return cudaErrorInvalidValue;
  if ((reinterpret_cast<uintptr_t>(a.A) & 15) || (reinterpret_cast<uintptr_t>(a.B) & 15) || (a.lda % 8) ||
      (a.ldb % 8) || (a.ldo % 8))
    return cudaErrorInvalidValue;
  // 128 x 256 tiles whenever N allows it; 128 x 128 otherwise (only used by generic/test shapes).
  if (a.N % 256 == 0) return launch_bn<256>(a, dev_status, num_sms, stream);
  return launch_bn<128>(a, dev_status, num_sms, stream);
}

}  // namespace jcb
