// Fused multi-head self-attention for seq_len = 128, head_dim = 64 (BERT-base, BASELINE.json
// config #5): one CTA per (batch, head), every GEMM on tcgen05 with TMEM accumulators, the
// S x S score / probability matrix never leaves the SM.
//
//   forward   S = Q K^T  (TMEM)  ->  row softmax in registers (one thread owns a row)  ->  P as bf16
//             into a 128B-swizzled smem A-operand tile  ->  O = P V (V as MN-major B operand)  ->
//             O / rowsum -> global.  Saves lse[row] = max + log(sum) for the backward pass.
//   backward  recompute S and P = exp(S - lse); dP = dO V^T; dS = P (dP - delta) with
//             delta = rowsum(dO * O); then three GEMMs out of smem-resident P / dS:
//             dQ = dS K,  dK = dS^T Q,  dV = P^T dO  (P and dS written once K-major and once
//             MN-major: the same rows, two block arrangements).
//
// q, k, v, o and their gradients are [B*S, H*D] row-major matrices (the projection outputs as
// they are): head h of batch b is the TMA box {64 columns from h*64, 128 rows from b*128}, so no
// head transpose exists anywhere.  The unfused path this replaces (ops/nn.py, round 1) ran two
// batched GEMMs, a softmax kernel and four transposes and wrote the S x S probabilities to HBM.
// No reference counterpart (the reference model is a 5x2 softmax regression, python-sdk/main.py:113-120).
#include <cuda_bf16.h>

#include "bflc_kernels.h"
#include "epi_common.cuh"
#include "launch.cuh"
#include "sm100_ptx.cuh"

namespace bflc {

namespace {

using epi::st_sw128;
__device__ __forceinline__ uint32_t pack2(float a, float b) { return epi::pack_bf16x2(a, b); }

constexpr int kS = 128, kD = 64;
constexpr int kTile = kS * 128;          // one [128 rows x 64 bf16] operand tile: 16 KB
constexpr int kThreads = 192;            // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue (one thread per row)
constexpr float kLog2e = 1.4426950408889634f;

struct AttnP {
  int H; long long ld; float scale;
  __nv_bfloat16* o; float* lse;                     // forward outputs ([B*S, ld], [B*H*S])
  const __nv_bfloat16* o_in; const __nv_bfloat16* dout_g;   // backward: saved O, dO (for delta)
  __nv_bfloat16* dq; __nv_bfloat16* dk; __nv_bfloat16* dv;
};

constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO 1024, v1, SWIZZLE_128B
__device__ __forceinline__ uint64_t desc(uint32_t lo) { return (static_cast<uint64_t>(kDescHi) << 32) | lo; }

// 4 (K = 64) or 8 (K = 128) UMMA K-steps of a [128 x N] tile.
//   K-major operand  : K-block kb at base + kb * 16 KB, K-step + 32 bytes
//   MN-major B (N=64): tile [K rows][128 B], K-step (16 rows) + 2048 bytes
//   MN-major A (M=128): K-block kb = two 8 KB M-chunks at base + kb * 16 KB, LBO = 8 KB, K-step + 2048
__device__ __forceinline__ void mma_kmaj_kmaj(uint32_t tmem, uint32_t a_addr, uint32_t b_addr, int n_cols, int kblocks) {
  const uint32_t idesc = ptx::make_idesc(1u, 0u, 0u, 128, static_cast<uint32_t>(n_cols));
  for (int kb = 0; kb < kblocks; ++kb)
#pragma unroll
    for (uint32_t k = 0; k < 4; ++k)
      ptx::umma_f16(tmem, desc(((a_addr + kb * kTile) >> 4 | (1u << 16)) + k * 2u),
                    desc(((b_addr + kb * kTile) >> 4 | (1u << 16)) + k * 2u), idesc, (kb > 0 || k > 0) ? 1u : 0u);
}
__device__ __forceinline__ void mma_kmaj_bmn(uint32_t tmem, uint32_t a_addr, uint32_t b_addr) {   // K = 128, N = 64
  const uint32_t idesc = ptx::make_idesc(1u, 0u, 1u, 128, 64);
#pragma unroll
  for (uint32_t ks = 0; ks < 8; ++ks)
    ptx::umma_f16(tmem, desc(((a_addr + (ks >> 2) * kTile) >> 4 | (1u << 16)) + (ks & 3u) * 2u),
                  desc((b_addr >> 4 | ((8192u >> 4) << 16)) + ks * (2048u >> 4)), idesc, ks > 0 ? 1u : 0u);
}
__device__ __forceinline__ void mma_amn_bmn(uint32_t tmem, uint32_t a_addr, uint32_t b_addr) {    // M = 128, K = 128, N = 64
  const uint32_t idesc = ptx::make_idesc(1u, 1u, 1u, 128, 64);
#pragma unroll
  for (uint32_t ks = 0; ks < 8; ++ks)
    ptx::umma_f16(tmem, desc(((a_addr + (ks >> 2) * 2 * 8192u) >> 4 | ((8192u >> 4) << 16)) + (ks & 3u) * (2048u >> 4)),
                  desc((b_addr >> 4 | ((8192u >> 4) << 16)) + ks * (2048u >> 4)), idesc, ks > 0 ? 1u : 0u);
}

// row m, 32-column chunk c of a [128 x 128] bf16 matrix held by thread m:
//   K-major A tile (K = column index):  K-block c/2, row m
//   MN-major A tile (K = row index):    K-block m/64, M-chunk c/2, row m % 64
__device__ __forceinline__ void put_kmaj(uint8_t* base, int m, int c, const uint4 (&u)[4]) {
  uint8_t* t = base + (c >> 1) * kTile;
#pragma unroll
  for (int jj = 0; jj < 4; ++jj) st_sw128(t, m, (c & 1) * 4 + jj, u[jj]);
}
__device__ __forceinline__ void put_mnmaj(uint8_t* base, int m, int c, const uint4 (&u)[4]) {
  uint8_t* t = base + (m >> 6) * kTile + (c >> 1) * 8192;
#pragma unroll
  for (int jj = 0; jj < 4; ++jj) st_sw128(t, m & 63, (c & 1) * 4 + jj, u[jj]);
}
__device__ __forceinline__ void pack32(const float (&v)[32], uint4 (&u)[4]) {
#pragma unroll
  for (int jj = 0; jj < 4; ++jj)
    u[jj] = make_uint4(pack2(v[8 * jj], v[8 * jj + 1]), pack2(v[8 * jj + 2], v[8 * jj + 3]),
                       pack2(v[8 * jj + 4], v[8 * jj + 5]), pack2(v[8 * jj + 6], v[8 * jj + 7]));
}
// accumulator rows -> bf16 global rows (64 columns = 128 bytes per thread), optionally scaled
__device__ __forceinline__ void store_rows64(uint32_t taddr, __nv_bfloat16* dst_row, float mul) {
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    uint32_t r[32];
    ptx::tmem_ld_32x32b_x32(taddr + c * 32, r);
    ptx::tmem_ld_wait();
    float v[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) v[k] = __uint_as_float(r[k]) * mul;
    uint4 u[4];
    pack32(v, u);
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) reinterpret_cast<uint4*>(dst_row + c * 32)[jj] = u[jj];
  }
}

// ------------------------------------------------------------------------------------ forward
constexpr int kFwdSmem = 3 * kTile + 2 * kTile + 256 + 1024;

__global__ void __launch_bounds__(kThreads, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnP p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem; uint8_t* sK = smem + kTile; uint8_t* sV = smem + 2 * kTile; uint8_t* sP = smem + 3 * kTile;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 5 * kTile);
  uint64_t* bar_in = bars; uint64_t* bar_s = bars + 1; uint64_t* bar_p = bars + 2; uint64_t* bar_o = bars + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  ptx::pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
  if (warp == 0 && lane == 0) {
    ptx::mbar_init(bar_in, 1); ptx::mbar_init(bar_s, 1); ptx::mbar_init(bar_o, 1);
    ptx::mbar_init(bar_p, 128);
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, 256);
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  ptx::pdl_wait();
  if (warp == 0) {
    if (ptx::elect_one()) {
      ptx::mbar_expect_tx(bar_in, 3 * kTile);
      ptx::tma_load_3d(sQ, &tmQ, bar_in, h * kD, b * kS, 0);
      ptx::tma_load_3d(sK, &tmK, bar_in, h * kD, b * kS, 0);
      ptx::tma_load_3d(sV, &tmV, bar_in, h * kD, b * kS, 0);
    }
    __syncwarp();
  } else if (warp == 1) {
    ptx::mbar_wait(bar_in, 0);
    ptx::tc_fence_after_sync();
    if (ptx::elect_one()) {
      mma_kmaj_kmaj(tmem, ptx::smem_u32(sQ), ptx::smem_u32(sK), 128, 1);      // S = Q K^T
      ptx::umma_commit(bar_s);
    }
    __syncwarp();
    ptx::mbar_wait(bar_p, 0);
    ptx::tc_fence_after_sync();
    if (ptx::elect_one()) {
      mma_kmaj_bmn(tmem + 128, ptx::smem_u32(sP), ptx::smem_u32(sV));         // O = P V
      ptx::umma_commit(bar_o);
    }
    __syncwarp();
  } else {
    const int q = warp & 3, m = q * 32 + lane;
    const uint32_t taddr = tmem + (static_cast<uint32_t>(q * 32) << 16);
    const float sc = p.scale * kLog2e;
    ptx::mbar_wait(bar_s, 0);
    ptx::tc_fence_after_sync();
    float mx = -INFINITY;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t r[32];
      ptx::tmem_ld_32x32b_x32(taddr + c * 32, r);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int k = 0; k < 32; ++k) mx = fmaxf(mx, __uint_as_float(r[k]));
    }
    float sum = 0.f;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t r[32];
      ptx::tmem_ld_32x32b_x32(taddr + c * 32, r);
      ptx::tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        v[k] = exp2f((__uint_as_float(r[k]) - mx) * sc);
        sum += v[k];
      }
      uint4 u[4];
      pack32(v, u);
      put_kmaj(sP, m, c, u);
    }
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before_sync();
    ptx::mbar_arrive(bar_p);
    const long long grow = static_cast<long long>(b) * kS + m;
    p.lse[static_cast<long long>(blockIdx.x) * kS + m] = mx * p.scale + __logf(sum);
    ptx::mbar_wait(bar_o, 0);
    ptx::tc_fence_after_sync();
    store_rows64(taddr + 128, p.o + grow * p.ld + h * kD, 1.f / sum);
    ptx::tc_fence_before_sync();
  }
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem, 256);
  }
}

// ----------------------------------------------------------------------------------- backward
constexpr int kBwdSmem = 4 * kTile + 3 * 2 * kTile + 256 + 1024;

__global__ void __launch_bounds__(kThreads, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                const AttnP p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem; uint8_t* sK = smem + kTile; uint8_t* sV = smem + 2 * kTile; uint8_t* sDO = smem + 3 * kTile;
  uint8_t* sPt = smem + 4 * kTile;        // P, MN-major arrangement (A of dV = P^T dO)
  uint8_t* sDSk = smem + 6 * kTile;       // dS, K-major (A of dQ = dS K)
  uint8_t* sDSt = smem + 8 * kTile;       // dS, MN-major (A of dK = dS^T Q)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 10 * kTile);
  uint64_t* bar_in = bars; uint64_t* bar_sdp = bars + 1; uint64_t* bar_ds = bars + 2; uint64_t* bar_out = bars + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  ptx::pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
  if (warp == 0 && lane == 0) {
    ptx::mbar_init(bar_in, 1); ptx::mbar_init(bar_sdp, 1); ptx::mbar_init(bar_out, 1);
    ptx::mbar_init(bar_ds, 128);
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, 512);
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  ptx::pdl_wait();
  if (warp == 0) {
    if (ptx::elect_one()) {
      ptx::mbar_expect_tx(bar_in, 4 * kTile);
      ptx::tma_load_3d(sQ, &tmQ, bar_in, h * kD, b * kS, 0);
      ptx::tma_load_3d(sK, &tmK, bar_in, h * kD, b * kS, 0);
      ptx::tma_load_3d(sV, &tmV, bar_in, h * kD, b * kS, 0);
      ptx::tma_load_3d(sDO, &tmDO, bar_in, h * kD, b * kS, 0);
    }
    __syncwarp();
  } else if (warp == 1) {
    ptx::mbar_wait(bar_in, 0);
    ptx::tc_fence_after_sync();
    if (ptx::elect_one()) {
      mma_kmaj_kmaj(tmem, ptx::smem_u32(sQ), ptx::smem_u32(sK), 128, 1);         // S  = Q K^T
      mma_kmaj_kmaj(tmem + 128, ptx::smem_u32(sDO), ptx::smem_u32(sV), 128, 1);  // dP = dO V^T
      ptx::umma_commit(bar_sdp);
    }
    __syncwarp();
    ptx::mbar_wait(bar_ds, 0);
    ptx::tc_fence_after_sync();
    if (ptx::elect_one()) {
      mma_kmaj_bmn(tmem + 256, ptx::smem_u32(sDSk), ptx::smem_u32(sK));          // dQ = dS K
      mma_amn_bmn(tmem + 320, ptx::smem_u32(sDSt), ptx::smem_u32(sQ));           // dK = dS^T Q
      mma_amn_bmn(tmem + 384, ptx::smem_u32(sPt), ptx::smem_u32(sDO));           // dV = P^T dO
      ptx::umma_commit(bar_out);
    }
    __syncwarp();
  } else {
    const int q = warp & 3, m = q * 32 + lane;
    const uint32_t taddr = tmem + (static_cast<uint32_t>(q * 32) << 16);
    const long long grow = static_cast<long long>(b) * kS + m;
    // delta = rowsum(dO * O) = rowsum(dP * P): from global while the first GEMMs run
    float delta = 0.f;
    {
      const uint4* o4 = reinterpret_cast<const uint4*>(p.o_in + grow * p.ld + h * kD);
      const uint4* d4 = reinterpret_cast<const uint4*>(p.dout_g + grow * p.ld + h * kD);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint4 a = o4[j], c = d4[j];
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, cw[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 fa = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&aw[e]));
          const float2 fc = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&cw[e]));
          delta += fa.x * fc.x + fa.y * fc.y;
        }
      }
    }
    const float lse = p.lse[static_cast<long long>(blockIdx.x) * kS + m];
    const float sc = p.scale * kLog2e, lse2 = lse * kLog2e;
    ptx::mbar_wait(bar_sdp, 0);
    ptx::tc_fence_after_sync();
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t rs[32], rp[32];
      ptx::tmem_ld_32x32b_x32(taddr + c * 32, rs);
      ptx::tmem_ld_32x32b_x32(taddr + 128 + c * 32, rp);
      ptx::tmem_ld_wait();
      float pv[32], ds[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        pv[k] = exp2f(__uint_as_float(rs[k]) * sc - lse2);
        ds[k] = pv[k] * (__uint_as_float(rp[k]) - delta) * p.scale;
      }
      uint4 u[4];
      pack32(pv, u);
      put_mnmaj(sPt, m, c, u);
      pack32(ds, u);
      put_kmaj(sDSk, m, c, u);
      put_mnmaj(sDSt, m, c, u);
    }
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before_sync();
    ptx::mbar_arrive(bar_ds);
    ptx::mbar_wait(bar_out, 0);
    ptx::tc_fence_after_sync();
    // accumulator row m of dQ is query row m; of dK / dV it is key row m
    store_rows64(taddr + 256, p.dq + grow * p.ld + h * kD, 1.f);
    store_rows64(taddr + 320, p.dk + grow * p.ld + h * kD, 1.f);
    store_rows64(taddr + 384, p.dv + grow * p.ld + h * kD, 1.f);
    ptx::tc_fence_before_sync();
  }
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem, 512);
  }
}

cudaError_t head_map(CUtensorMap* out, const void* ptr, long long ld, long long rows, int hd) {
  GemmOperand op{ptr, ld, 0, false};
  return gemm_make_operand_map(out, op, DType::BF16, static_cast<int>(rows), hd, 1, kS);
}

}  // namespace

cudaError_t attention_fwd_sm100(const void* q, const void* k, const void* v, void* o, float* lse, int B, int S,
                                int H, int D, long long ld, float scale, cudaStream_t stream) {
  bind_context_once();
  if (S != kS || D != kD || ld % 8 != 0 || B <= 0 || H <= 0) return cudaErrorNotSupported;
  CUtensorMap tq, tk, tv;
  cudaError_t e;
  const long long rows = static_cast<long long>(B) * S;
  if ((e = head_map(&tq, q, ld, rows, H * D)) != cudaSuccess) return e;
  if ((e = head_map(&tk, k, ld, rows, H * D)) != cudaSuccess) return e;
  if ((e = head_map(&tv, v, ld, rows, H * D)) != cudaSuccess) return e;
  AttnP p{};
  p.H = H; p.ld = ld; p.scale = scale; p.o = static_cast<__nv_bfloat16*>(o); p.lse = lse;
  static bool cfg = false;
  if (!cfg) {
    e = cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmem);
    if (e != cudaSuccess) return e;
    cfg = true;
  }
  note_launch();
  return launch_pdl(attn_fwd_kernel, dim3(B * H), dim3(kThreads), kFwdSmem, stream, tq, tk, tv, p);
}

cudaError_t attention_bwd_sm100(const void* q, const void* k, const void* v, const void* o, const void* dout,
                                const float* lse, void* dq, void* dk, void* dv, int B, int S, int H, int D,
                                long long ld, float scale, cudaStream_t stream) {
  bind_context_once();
  if (S != kS || D != kD || ld % 8 != 0 || B <= 0 || H <= 0) return cudaErrorNotSupported;
  CUtensorMap tq, tk, tv, tdo;
  cudaError_t e;
  const long long rows = static_cast<long long>(B) * S;
  if ((e = head_map(&tq, q, ld, rows, H * D)) != cudaSuccess) return e;
  if ((e = head_map(&tk, k, ld, rows, H * D)) != cudaSuccess) return e;
  if ((e = head_map(&tv, v, ld, rows, H * D)) != cudaSuccess) return e;
  if ((e = head_map(&tdo, dout, ld, rows, H * D)) != cudaSuccess) return e;
  AttnP p{};
  p.H = H; p.ld = ld; p.scale = scale; p.lse = const_cast<float*>(lse);
  p.o_in = static_cast<const __nv_bfloat16*>(o); p.dout_g = static_cast<const __nv_bfloat16*>(dout);
  p.dq = static_cast<__nv_bfloat16*>(dq); p.dk = static_cast<__nv_bfloat16*>(dk); p.dv = static_cast<__nv_bfloat16*>(dv);
  static bool cfg = false;
  if (!cfg) {
    e = cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem);
    if (e != cudaSuccess) return e;
    cfg = true;
  }
  note_launch();
  return launch_pdl(attn_bwd_kernel, dim3(B * H), dim3(kThreads), kBwdSmem, stream, tq, tk, tv, tdo, p);
}

}  // namespace bflc
