// Block-scaled fp8 GEMM (OCP MXFP8: e4m3 elements, one UE8M0 scale per 32 K-elements) on
// tcgen05:  D (M x N) = alpha * (A .* SFA) (B .* SFB)^T [+ bias] [act]
//
//   tcgen05.mma.cta_group::1.kind::mxf8f6f4.block_scale [d], adesc, bdesc, idesc, [sfa], [sfb], p
//
// A [M, K] and B [N, K] are K-major e4m3, staged by TMA into 128-byte-swizzled smem (one
// K-block = 128 elements = one swizzle row).  The scale factors never pass through registers:
// the quantiser (k_quantize_mx8 below) writes them to global memory already in the layout the
// tensor core wants -- per (128-row block, 128-K block) a 512-byte chunk whose byte
// [r%32][r/32][k/32] is the scale of row r, K-group k -- a `cp.async.bulk` drops the chunk into
// smem with the same mbarrier transaction as the operand tiles, and ONE
// `tcgen05.cp.32x128b.warpx4` per chunk copies it into 4 TMEM columns (replicated over the four
// lane quarters).  The 4 UMMAs (K = 32 each) of a K-block select their scale byte with the
// a_sf_id / b_sf_id fields of the instruction descriptor.  tcgen05.cp and tcgen05.mma execute in
// issue order, so the single SF TMEM region is reused by every stage without extra barriers.
//
//   warp 0  producer (TMA tiles + bulk SF chunks)     warp 1  TMEM alloc, cp + UMMA issue
//   warps 2-5  epilogue (tcgen05.ld -> bias/act -> smem staging -> coalesced stores)
//
// Reference parity: the reference trains in fp32 on CPU (python-sdk/main.py:120-123); BASELINE.json
// names block-scaled fp8 for the MLP / LeNet-5 configs, this is that compute path.
#include <cuda_bf16.h>
#include <cuda_fp8.h>

#include <cstring>

#include "bflc_kernels.h"
#include "epi_common.cuh"
#include "launch.cuh"
#include "sm100_ptx.cuh"

namespace bflc {

namespace {

constexpr int kBM = 128;
constexpr int kBK = 128;                 // fp8 elements per K-block (= 128 bytes)
using epi::kSfChunk;                     // bytes of scale factors per (128 rows, 128 K)
using epi::kStgLd;
using epi::kStgBytes;
constexpr int kBarBytes = 256;
constexpr int kThreads = 192;

template <int BN> struct Cfg {
  static constexpr int kStages = BN == 256 ? 4 : 6;
  static constexpr int kABytes = kBM * 128, kBBytes = BN * 128;
  static constexpr int kTileStage = kABytes + kBBytes;
  static constexpr int kSfbBytes = (BN / 128) * kSfChunk;
  static constexpr int kSfStage = kSfChunk + kSfbBytes;
  static constexpr int kTiles = kStages * kTileStage;
  static constexpr int kSfOff = kTiles;
  static constexpr int kBarOff = kSfOff + kStages * kSfStage;
  static constexpr int kStgOff = kBarOff + kBarBytes;
  static constexpr int kBiasOff = kStgOff + kStgBytes;
  static constexpr int kSmem = kBiasOff + BN * 4 + 1024;
  static constexpr int kTmemCols = BN == 256 ? 512 : 256;   // accumulator + SFA(4) + SFB(BN/32)
  static constexpr uint32_t kTxBytes = kTileStage + kSfStage;
};

struct PM {
  int M, N, K, k_blocks;
  const uint8_t* sfa; const uint8_t* sfb;   // canonical chunk arrays [row_block][k_block][512]
  void* d; int d_dtype; long long ldd; float alpha;
  const float* bias; int act;
};

using epi::bulk_g2s;
using epi::sf_desc;
using epi::utccp_32x128b_warpx4;
using epi::umma_mx8;
using epi::make_idesc_mx8;
__device__ __forceinline__ uint32_t pack2(float a, float b) { return epi::pack_bf16x2(a, b); }
__device__ __forceinline__ float gelu_f(float x) {
  return 0.5f * x * (1.f + erff(x * 0.70710678118654752f));
}

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_mx8_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const PM p) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::kBarOff);
  uint64_t* empty_bar = full_bar + C::kStages;
  uint64_t* accum_bar = empty_bar + C::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);
  float* stage_base = reinterpret_cast<float*>(smem + C::kStgOff);
  float* sbias = reinterpret_cast<float*>(smem + C::kBiasOff);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * kBM, n0 = blockIdx.y * BN;
  ptx::pdl_launch_dependents();

  if (warp == 0 && lane == 0) {
    ptx::tma_prefetch_desc(&tmA);
    ptx::tma_prefetch_desc(&tmB);
    for (int s = 0; s < C::kStages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    ptx::mbar_init(accum_bar, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, C::kTmemCols);
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const int n_kb = p.k_blocks;
  ptx::pdl_wait();

  if (warp == 0) {
    const uint8_t* sfa_src = p.sfa + static_cast<long long>(blockIdx.x) * n_kb * kSfChunk;
    const uint8_t* sfb_src = p.sfb + static_cast<long long>(blockIdx.y) * (BN / 128) * n_kb * kSfChunk;
    for (int i = 0; i < n_kb; ++i) {
      const int s = i % C::kStages;
      const uint32_t ph = (i / C::kStages) & 1;
      ptx::mbar_wait(&empty_bar[s], ph ^ 1);
      if (ptx::elect_one()) {
        uint8_t* sa = smem + s * C::kTileStage;
        uint8_t* sf = smem + C::kSfOff + s * C::kSfStage;
        ptx::mbar_expect_tx(&full_bar[s], C::kTxBytes);
        ptx::tma_load_3d(sa, &tmA, &full_bar[s], i * kBK, m0, 0);
        ptx::tma_load_3d(sa + C::kABytes, &tmB, &full_bar[s], i * kBK, n0, 0);
        bulk_g2s(sf, sfa_src + static_cast<long long>(i) * kSfChunk, kSfChunk, &full_bar[s]);
#pragma unroll
        for (int j = 0; j < BN / 128; ++j)
          bulk_g2s(sf + kSfChunk + j * kSfChunk,
                   sfb_src + (static_cast<long long>(j) * n_kb + i) * kSfChunk, kSfChunk, &full_bar[s]);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    const uint32_t idesc0 = make_idesc_mx8(kBM, BN);
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO 1024, version 1, SWIZZLE_128B
    const uint32_t base_lo = ptx::smem_u32(smem) >> 4;
    const uint32_t tsfa = tmem_base + BN;
    const uint32_t tsfb = tmem_base + BN + 4;
    for (int i = 0; i < n_kb; ++i) {
      const int s = i % C::kStages;
      const uint32_t ph = (i / C::kStages) & 1;
      ptx::mbar_wait(&full_bar[s], ph);
      ptx::tc_fence_after_sync();
      if (ptx::elect_one()) {
        const uint32_t sf_addr = ptx::smem_u32(smem + C::kSfOff + s * C::kSfStage);
        utccp_32x128b_warpx4(tsfa, sf_desc(sf_addr));
#pragma unroll
        for (int j = 0; j < BN / 128; ++j)
          utccp_32x128b_warpx4(tsfb + 4 * j, sf_desc(sf_addr + kSfChunk + j * kSfChunk));
        const uint32_t lo_a = (base_lo + static_cast<uint32_t>(s) * (C::kTileStage >> 4)) | (1u << 16);
        const uint32_t lo_b = lo_a + (C::kABytes >> 4);
#pragma unroll
        for (uint32_t k = 0; k < 4; ++k) {
          const uint64_t ad = (static_cast<uint64_t>(hi) << 32) | (lo_a + k * 2u);
          const uint64_t bd = (static_cast<uint64_t>(hi) << 32) | (lo_b + k * 2u);
          const uint32_t idesc = idesc0 | (k << 29) | (k << 4);
          umma_mx8(tmem_base, ad, bd, idesc, (i > 0 || k > 0) ? 1u : 0u, tsfa, tsfb);
        }
        ptx::umma_commit(&empty_bar[s]);
      }
      __syncwarp();
    }
    if (ptx::elect_one()) ptx::umma_commit(accum_bar);
    __syncwarp();
  } else {
    const int q = warp & 3;
    float* stg = stage_base + (warp - 2) * (32 * kStgLd);
    const int row_base = m0 + q * 32;
    const int cr = lane >> 3, cg = (lane & 7) * 4;
    {
      const int et = threadIdx.x - 64;
      for (int i = et; i < BN; i += 128)
        sbias[i] = (p.bias != nullptr && n0 + i < p.N) ? __ldg(p.bias + n0 + i) : 0.f;
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    ptx::mbar_wait(accum_bar, 0);
    ptx::tc_fence_after_sync();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      const int nc = n0 + c * 32;
      if (nc >= p.N) break;
      uint32_t r[32];
      ptx::tmem_ld_32x32b_x32(taddr + c * 32, r);
      ptx::tmem_ld_wait();
      float4* rowp = reinterpret_cast<float4*>(stg + lane * kStgLd);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float x = __uint_as_float(r[4 * j + k]) * p.alpha + sbias[c * 32 + 4 * j + k];
          if (p.act == 1) x = fmaxf(x, 0.f);
          else if (p.act == 2) x = gelu_f(x);
          v[k] = x;
        }
        rowp[j] = make_float4(v[0], v[1], v[2], v[3]);
      }
      __syncwarp();
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int rr = it * 4 + cr, rw = row_base + rr, col = nc + cg;
        if (rw >= p.M || col >= p.N) continue;
        const float4 x = *reinterpret_cast<const float4*>(stg + rr * kStgLd + cg);
        const long long off = static_cast<long long>(rw) * p.ldd + col;
        const bool vec = col + 3 < p.N;
        if (p.d_dtype == 0) {
          float* d = reinterpret_cast<float*>(p.d) + off;
          if (vec) *reinterpret_cast<float4*>(d) = x;
          else {
            const float xs[4] = {x.x, x.y, x.z, x.w};
            for (int k = 0; k < 4; ++k) if (col + k < p.N) d[k] = xs[k];
          }
        } else {
          __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(p.d) + off;
          if (vec) *reinterpret_cast<uint2*>(d) = make_uint2(pack2(x.x, x.y), pack2(x.z, x.w));
          else {
            const float xs[4] = {x.x, x.y, x.z, x.w};
            for (int k = 0; k < 4; ++k) if (col + k < p.N) d[k] = __float2bfloat16(xs[k]);
          }
        }
      }
      __syncwarp();
    }
    ptx::tc_fence_before_sync();
  }
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

// ------------------------------------------------------------------ quantiser
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f<uint8_t>(uint8_t v) { return static_cast<float>(v); }

// One thread per (row, 32-element K group) of the PADDED problem (rows to a multiple of 128,
// groups to a multiple of 4): amax -> UE8M0 exponent e = ceil(log2(amax / 448)) -> e4m3
// satfinite(x * 2^-e).  Padding groups / rows get scale 1.0 (0x7F; never NaN) and no data.
template <typename T>
__global__ void __launch_bounds__(256)
k_quantize_mx8(const T* __restrict__ x, long long ldx, int R, int K, float in_scale,
               uint8_t* __restrict__ q, long long ldq, uint8_t* __restrict__ sf, int k_blocks) {
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  const int groups = k_blocks * 4;
  const long long gid = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const int row = static_cast<int>(gid / groups), g = static_cast<int>(gid % groups);
  const int r_pad = (R + 255) / 256 * 256;   // 256: a BN = 256 tile reads two row blocks
  if (row >= r_pad) return;
  const int r = row & 127, rb = row >> 7;
  uint8_t* sfp = sf + (static_cast<long long>(rb) * k_blocks + (g >> 2)) * kSfChunk + (r & 31) * 16 +
                 (r >> 5) * 4 + (g & 3);
  const int k0 = g * 32;
  if (row >= R || k0 >= K) { *sfp = 127; return; }
  const int n = min(32, K - k0);
  float v[32];
  const T* xp = x + static_cast<long long>(row) * ldx + k0;
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = i < n ? to_f<T>(xp[i]) * in_scale : 0.f;
  uint32_t w[8];
  *sfp = static_cast<uint8_t>(epi::mx8_quant32(v, w));
  uint8_t* qp = q + static_cast<long long>(row) * ldq + k0;
  if (k0 + 32 <= ldq) {   // whole group inside the (16-byte padded) row pitch: two 16-byte stores
    reinterpret_cast<uint4*>(qp)[0] = make_uint4(w[0], w[1], w[2], w[3]);
    reinterpret_cast<uint4*>(qp)[1] = make_uint4(w[4], w[5], w[6], w[7]);
  } else {
    const int nb = static_cast<int>(ldq) - k0;
    for (int i = 0; i < nb; ++i) qp[i] = static_cast<uint8_t>(w[i >> 2] >> (8 * (i & 3)));
  }
}

}  // namespace

cudaError_t gemm_mx8_sm100(const Mx8Problem& p, cudaStream_t stream) {
  bind_context_once();
  if (p.M <= 0 || p.N <= 0 || p.K <= 0 || p.ldd % 4 != 0 || p.lda % 16 != 0 || p.ldb % 16 != 0)
    return cudaErrorInvalidValue;
  const int mt = (p.M + kBM - 1) / kBM;
  const int BN = (p.N > 128 && static_cast<long long>((p.N + 255) / 256) * mt >= 100) ? 256 : 128;
  CUtensorMap ta, tb;
  GemmOperand oa{p.a, p.lda, 0, false}, ob{p.b, p.ldb, 0, false};
  cudaError_t e = gemm_make_operand_map(&ta, oa, DType::FP8_E4M3, p.M, p.K, 1, kBM);
  if (e != cudaSuccess) return e;
  e = gemm_make_operand_map(&tb, ob, DType::FP8_E4M3, p.N, p.K, 1, BN);
  if (e != cudaSuccess) return e;
  PM kp{};
  kp.M = p.M; kp.N = p.N; kp.K = p.K; kp.k_blocks = (p.K + kBK - 1) / kBK;
  kp.sfa = p.sfa; kp.sfb = p.sfb;
  kp.d = p.d; kp.d_dtype = static_cast<int>(p.d_dtype); kp.ldd = p.ldd; kp.alpha = p.alpha;
  kp.bias = p.bias; kp.act = static_cast<int>(p.act);
  dim3 grid(mt, (p.N + BN - 1) / BN, 1);
  note_launch();
  if (BN == 256) {
    static bool cfg = false;
    if (!cfg) {
      e = cudaFuncSetAttribute(gemm_mx8_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<256>::kSmem);
      if (e != cudaSuccess) return e;
      cfg = true;
    }
    return launch_pdl(gemm_mx8_kernel<256>, grid, dim3(kThreads), Cfg<256>::kSmem, stream, ta, tb, kp);
  }
  static bool cfg = false;
  if (!cfg) {
    e = cudaFuncSetAttribute(gemm_mx8_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<128>::kSmem);
    if (e != cudaSuccess) return e;
    cfg = true;
  }
  return launch_pdl(gemm_mx8_kernel<128>, grid, dim3(kThreads), Cfg<128>::kSmem, stream, ta, tb, kp);
}

long long mx8_sf_bytes(int rows, int K) {
  return static_cast<long long>((rows + 255) / 256 * 2) * ((K + kBK - 1) / kBK) * kSfChunk;
}

cudaError_t quantize_mx8(const void* x, DType x_dtype, long long ldx, int R, int K, float in_scale,
                         void* q, long long ldq, void* sf, cudaStream_t stream) {
  if (R <= 0 || K <= 0 || ldq % 16 != 0 || ldq < K) return cudaErrorInvalidValue;
  const int k_blocks = (K + kBK - 1) / kBK;
  const long long total = static_cast<long long>((R + 255) / 256 * 256) * k_blocks * 4;
  const dim3 grid(static_cast<unsigned>((total + 255) / 256)), block(256);
  uint8_t* q8 = static_cast<uint8_t*>(q);
  uint8_t* sf8 = static_cast<uint8_t*>(sf);
  note_launch();
  switch (x_dtype) {
    case DType::F32:
      return launch_pdl(k_quantize_mx8<float>, grid, block, 0, stream, static_cast<const float*>(x), ldx, R,
                        K, in_scale, q8, ldq, sf8, k_blocks);
    case DType::BF16:
      return launch_pdl(k_quantize_mx8<__nv_bfloat16>, grid, block, 0, stream,
                        static_cast<const __nv_bfloat16*>(x), ldx, R, K, in_scale, q8, ldq, sf8, k_blocks);
    case DType::U8:
      return launch_pdl(k_quantize_mx8<uint8_t>, grid, block, 0, stream, static_cast<const uint8_t*>(x), ldx,
                        R, K, in_scale, q8, ldq, sf8, k_blocks);
    default:
      return cudaErrorInvalidValue;
  }
}

}  // namespace bflc
