// tcgen05 / TMEM / TMA GEMM for sm_100a with fused epilogues.
//
//   D[b] (M x N) = alpha * A[b] (M x K) . B[b]^T (N x K)   bf16 or fp8(e4m3) in, fp32 accumulate
//
// One CTA computes one 128 x BN output tile:
//   warp 0   : TMA producer  (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier tx)
//   warp 1   : TMEM allocator + single-thread tcgen05.mma issuer (accumulator in TMEM)
//   warps 2-5: epilogue (tcgen05.ld 32x32b -> registers -> fused math -> global)
//
// Operands may be K-major or MN-major (transposed views), so forward (x.W^T),
// input-gradient (dY.W) and weight-gradient (dY^T.X) GEMMs all run without a transpose
// pass.  The B operand may come from a per-batch tensor-map array in device memory, whose
// maps may point into *peer GPUs'* HBM: the committee's validation GEMM pulls each
// trainer's candidate weights over NVLink tile by tile (hot path 1, X5 in SURVEY.md 2.7b).
//
// Epilogues: bias / ReLU / GELU / activation-backward masks / column sums (bias grads) /
// split-K atomics; a row-wise softmax-cross-entropy epilogue that emits dlogits + loss +
// #correct; an argmax-accuracy epilogue (the committee score, python-sdk/main.py:182-183).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_bf16.h>
#include <cuda_fp8.h>
#include <mutex>

#include "bflc_kernels.h"
#include "epi_common.cuh"
#include "launch.cuh"
#include "sm100_ptx.cuh"

namespace bflc {

namespace {

constexpr int kBM = 128;             // UMMA M
constexpr int kStageKBytes = 128;    // one swizzle-128B span of K per stage row
constexpr int kThreads = 192;
constexpr int kABytes = kBM * kStageKBytes;  // 16 KB

struct KParams {
  int M, N, K, batch;
  int is_fp8;
  int a_mn, b_mn;           // 1 = MN-major
  int a_batched, b_batched;
  int k_blocks;             // total K blocks of BLOCK_K elements
  int split_k;
  const CUtensorMap* b_maps_dev;
  const GemmDynamic* dyn;
  const int* pred;
  long long* dbg_times;  // optional [8] clock64 stamps written by CTA (0,0,0) (bring-up only)
  // epilogue
  void* d;
  int d_dtype;
  long long ldd, d_batch_stride;
  float alpha;
  const float* bias;
  const float* const* bias_ptrs;
  int act;
  void* aux_out;
  const void* aux_in;
  int act_bwd;
  float* colsum;
  int accumulate;
  const int32_t* labels;
  long long labels_batch_stride;
  float grad_scale;
  float* loss_sum;
  unsigned int* correct;
  uint32_t lbo_a, sbo_a, lbo_b, sbo_b;
  uint32_t kstep_a, kstep_b;  // descriptor start-address advance per UMMA_K step (bytes)
  int vec_ok;                 // d / aux rows are 16-byte aligned -> vectorised global access
  // implicit-GEMM convolution (ConvView): which operand is the shifted NHWC view and its taps
  int cv_mode, cv_flip;
  int cv_cb;                  // 64-channel blocks per filter tap
  int cv_kw, cv_pad, cv_stride;
  int cv_oh, cv_ow;           // pixel grid enumerated by the rows (mode 1) / the reduction (mode 2)
  int cv_c;                   // channels of the viewed activation
  int stages;                 // smem ring depth of this launch (<= SmemLayout::kStages)
};

template <int BN>
struct SmemLayout {
  static constexpr int kBBytes = BN * kStageKBytes;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int kTileBytes = kStages * kStageBytes;
  static constexpr int kBarBytes = 256;
  static constexpr int kStgBytes = 4 * 32 * 36 * 4;  // per-epilogue-warp [32][36] fp32 staging
  static constexpr int kBiasBytes = BN * 4;
  static constexpr int kTotal = kTileBytes + kBarBytes + kStgBytes + kBiasBytes + 1024;
};

using epi::kStgLd;
using epi::col_sum32;
using epi::stage_put;
using epi::stage_get;
using epi::pack_bf16x2;

// Staged [32][32] fp32 tile -> global.  Lane (cr, cg) moves 4 consecutive columns of row
// it*4+cr, so one warp instruction covers 4 rows x 128 B.  DT: 0 fp32, 1 bf16, 2 fp8(e4m3).
template <int DT, bool RED>
__device__ __forceinline__ void tile_store(const float* stg, void* dptr, long long tile_off,
                                           long long ldd, int row_base, int nc, int M, int N,
                                           int cr, int cg, int vec_ok) {
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int rr = it * 4 + cr;
    const int row = row_base + rr;
    const int col = nc + cg;
    if (row >= M || col >= N) continue;
    const float4 x = *reinterpret_cast<const float4*>(stg + rr * kStgLd + cg);
    const float xs[4] = {x.x, x.y, x.z, x.w};
    const long long off = tile_off + static_cast<long long>(row) * ldd + col;
    const bool vec = vec_ok && (col + 3 < N);
    if constexpr (DT == 0) {
      float* d = reinterpret_cast<float*>(dptr) + off;
      if constexpr (RED) {
        if (vec) ptx::red_add_f32x4(d, x.x, x.y, x.z, x.w);
        else
          for (int k = 0; k < 4; ++k) if (col + k < N) atomicAdd(d + k, xs[k]);
      } else {
        if (vec) *reinterpret_cast<float4*>(d) = x;
        else
          for (int k = 0; k < 4; ++k) if (col + k < N) d[k] = xs[k];
      }
    } else if constexpr (DT == 1) {
      __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(dptr) + off;
      if (vec) *reinterpret_cast<uint2*>(d) = make_uint2(pack_bf16x2(x.x, x.y), pack_bf16x2(x.z, x.w));
      else
        for (int k = 0; k < 4; ++k) if (col + k < N) d[k] = __float2bfloat16(xs[k]);
    } else {
      __nv_fp8_e4m3* d = reinterpret_cast<__nv_fp8_e4m3*>(dptr) + off;
      for (int k = 0; k < 4; ++k) if (col + k < N) d[k] = __nv_fp8_e4m3(xs[k]);
    }
  }
}
// global bf16 tile -> staged fp32 tile (zeros outside the matrix)
__device__ __forceinline__ void tile_load_bf16(float* stg, const void* sptr, long long tile_off,
                                               long long ldd, int row_base, int nc, int M, int N,
                                               int cr, int cg, int vec_ok) {
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int rr = it * 4 + cr;
    const int row = row_base + rr;
    const int col = nc + cg;
    float xs[4] = {0.f, 0.f, 0.f, 0.f};
    if (row < M && col < N) {
      const __nv_bfloat16* s = reinterpret_cast<const __nv_bfloat16*>(sptr) + tile_off +
                               static_cast<long long>(row) * ldd + col;
      if (vec_ok && col + 3 < N) {
        const uint2 u = *reinterpret_cast<const uint2*>(s);
        const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
        const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
        xs[0] = a.x; xs[1] = a.y; xs[2] = b.x; xs[3] = b.y;
      } else {
        for (int k = 0; k < 4; ++k) if (col + k < N) xs[k] = __bfloat162float(s[k]);
      }
    }
    *reinterpret_cast<float4*>(stg + rr * kStgLd + cg) = make_float4(xs[0], xs[1], xs[2], xs[3]);
  }
}

__device__ __forceinline__ float gelu_f(float x) {
  return 0.5f * x * (1.f + erff(x * 0.70710678118654752f));
}
__device__ __forceinline__ float gelu_grad_f(float x) {
  const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752f));
  const float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

template <int BN, int EPI>
__global__ void __launch_bounds__(kThreads, (BN == 64 ? 2 : 1))
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmC, const KParams p) {
  using L = SmemLayout<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  // Ring depth is a launch parameter: short-K problems take a shallow ring so that two CTAs fit
  // one SM and one CTA's epilogue overlaps the other's main loop.
  const int n_stages = p.stages;
  const int tile_bytes = n_stages * L::kStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + tile_bytes);
  uint64_t* empty_bar = full_bar + L::kStages;
  uint64_t* accum_bar = empty_bar + L::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);
  float* stage_base = reinterpret_cast<float*>(smem + tile_bytes + L::kBarBytes);
  float* sbias = stage_base + 4 * 32 * kStgLd;

  // Programmatic dependent launch: let the next kernel of the stream get scheduled now; this
  // CTA's own prologue (barrier init, TMEM alloc, descriptor prefetch) runs before it waits
  // for its predecessors' memory.
  ptx::pdl_launch_dependents();
  const long long t_entry = clock64();
  const bool dbg = p.dbg_times != nullptr && (blockIdx.x | blockIdx.y | blockIdx.z) == 0;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN;
  const int m0 = blockIdx.y * kBM;
  const int z = blockIdx.z;
  const int bidx = z / p.split_k;
  const int split = z - bidx * p.split_k;
  const int kb_per = (p.k_blocks + p.split_k - 1) / p.split_k;
  const int kb_begin = split * kb_per;
  const int kb_end = min(p.k_blocks, kb_begin + kb_per);
  const int n_kb = max(0, kb_end - kb_begin);
  const int block_k = p.is_fp8 ? 128 : 64;  // elements per 128-byte span

  if (warp == 0 && lane == 0) {
    ptx::tma_prefetch_desc(&tmA);
    if (p.b_maps_dev == nullptr) ptx::tma_prefetch_desc(&tmB);
    if (p.cv_mode != 0) ptx::tma_prefetch_desc(&tmC);
    for (int s = 0; s < n_stages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    ptx::mbar_init(accum_bar, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, BN);
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  if (dbg && threadIdx.x == 0) { p.dbg_times[0] = t_entry; p.dbg_times[1] = clock64(); }

  // Everything below touches memory produced by earlier kernels of the stream.
  ptx::pdl_wait();
  // role predication / dynamic batch count (device-resident, written by the plan kernel)
  const bool inactive = (p.pred != nullptr && *p.pred == 0) ||
                        (p.dyn != nullptr && bidx >= p.dyn->active_batches);
  if (inactive) {  // CTA-uniform
    if (warp == 1) ptx::tmem_dealloc(tmem_base, BN);
    return;
  }
  const CUtensorMap* mapB =
      p.b_maps_dev ? (p.b_maps_dev + (p.dyn ? p.dyn->map_index[bidx] : bidx)) : &tmB;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    // The whole warp runs the loop (warp-uniform control flow keeps addresses / coordinates in
    // uniform registers); one elected lane issues the copies.
    const int ca2 = p.a_batched ? bidx : 0;
    const int cb2 = (p.b_batched && !p.b_maps_dev) ? bidx : 0;
    if (p.dyn != nullptr && p.dyn->wait_flag[bidx] != nullptr) {
      // B lives in a peer's upload buffer: wait until that trainer released it, then make
      // the acquired state visible to the async proxy before the first TMA pull.
      if (lane == 0) ptx::wait_flag_ge(p.dyn->wait_flag[bidx], p.dyn->wait_value);
      __syncwarp();
      ptx::fence_proxy_async_all();
    }
    for (int i = 0; i < n_kb; ++i) {
      const int s = i % n_stages;
      const uint32_t ph = (i / n_stages) & 1;
      ptx::mbar_wait(&empty_bar[s], ph ^ 1);
      uint8_t* sa = smem + s * L::kStageBytes;
      uint8_t* sb = sa + kABytes;
      const int k0 = (kb_begin + i) * block_k;
      if (p.cv_mode != 0) {
        // ---- implicit-GEMM convolution: one operand is a tap-shifted NHWC box (tmC)
        if (ptx::elect_one()) {
          ptx::mbar_expect_tx(&full_bar[s], L::kStageBytes);
          const int kb = kb_begin + i;
          const int per_img = p.cv_oh * p.cv_ow;
          if (p.cv_mode == 1) {
            // rows m0.. = 128 consecutive output pixels (whole image rows); K block = (tap, 64 ch)
            const int tap = kb / p.cv_cb, cblk = kb - tap * p.cv_cb;
            const int r = tap / p.cv_kw, t = tap - r * p.cv_kw;
            const int img0 = m0 / per_img, oh0 = (m0 - img0 * per_img) / p.cv_ow;
            const int dh = p.cv_flip ? p.cv_pad - r : r - p.cv_pad;
            const int dw = p.cv_flip ? p.cv_pad - t : t - p.cv_pad;
            ptx::tma_load_4d(sa, &tmC, &full_bar[s], cblk * 64, dw, oh0 * p.cv_stride + dh, img0);
            if (!p.cv_flip) {
              ptx::tma_load_3d(sb, mapB, &full_bar[s], k0, n0, 0);
            } else {  // w[Cout][taps*Cin] read MN-major: N offset selects the tap, K = out channels
              for (int b = 0; b < BN / 64; ++b)
                ptx::tma_load_3d(sb + b * (64 * 128), mapB, &full_bar[s],
                                 tap * p.N + n0 + b * 64, cblk * 64, 0);
            }
          } else {
            // weight gradient: K block = 64 output pixels; N tile = (tap, BN channels)
            for (int b = 0; b < 2; ++b)  // A = dy^T, MN-major: two 64-wide boxes of out channels
              ptx::tma_load_3d(sa + b * (64 * 128), &tmA, &full_bar[s], m0 + b * 64, k0, 0);
            const int tap = n0 / p.cv_c, c0 = n0 - tap * p.cv_c;
            const int r = tap / p.cv_kw, t = tap - r * p.cv_kw;
            const int pix0 = kb * 64;
            const int img0 = pix0 / per_img, oh0 = (pix0 - img0 * per_img) / p.cv_ow;
            for (int b = 0; b < BN / 64; ++b)
              ptx::tma_load_4d(sb + b * (64 * 128), &tmC, &full_bar[s], c0 + b * 64, t - p.cv_pad,
                               oh0 * p.cv_stride + r - p.cv_pad, img0);
          }
        }
        __syncwarp();
        continue;
      }
      if (ptx::elect_one()) {
        ptx::mbar_expect_tx(&full_bar[s], L::kStageBytes);
        if (!p.a_mn) {
          ptx::tma_load_3d(sa, &tmA, &full_bar[s], k0, m0, ca2);
        } else {
          // MN-major: boxes of [block_k rows of K][128 B of M]; kBM*es/128 boxes
          const int nbox = p.is_fp8 ? 1 : 2;
          for (int b = 0; b < nbox; ++b)
            ptx::tma_load_3d(sa + b * (block_k * 128), &tmA, &full_bar[s], m0 + b * block_k, k0,
                             ca2);
        }
        if (!p.b_mn) {
          ptx::tma_load_3d(sb, mapB, &full_bar[s], k0, n0, cb2);
        } else {
          const int nbox = BN / block_k;
          for (int b = 0; b < nbox; ++b)
            ptx::tma_load_3d(sb + b * (block_k * 128), mapB, &full_bar[s], n0 + b * block_k, k0,
                             cb2);
        }
        if (dbg && i == 0) p.dbg_times[2] = clock64();
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------- MMA issuer
    // Warp-uniform loop; descriptors are built from 32-bit uniform arithmetic (only the low
    // word -- start address -- changes per stage / K step) and one elected lane issues.
    const uint32_t idesc =
        ptx::make_idesc(p.is_fp8 ? 0u : 1u, p.a_mn ? 1u : 0u, p.b_mn ? 1u : 0u, kBM, BN);
    const uint32_t hi_a = ((p.sbo_a >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
    const uint32_t hi_b = ((p.sbo_b >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
    const uint32_t base_lo = (ptx::smem_u32(smem) >> 4);
    const uint32_t lo_a0 = base_lo | (((p.lbo_a >> 4) & 0x3FFFu) << 16);
    const uint32_t lo_b0 = (base_lo + (kABytes >> 4)) | (((p.lbo_b >> 4) & 0x3FFFu) << 16);
    const uint32_t ks_a = p.kstep_a >> 4, ks_b = p.kstep_b >> 4;
    const bool fp8 = p.is_fp8 != 0;
    for (int i = 0; i < n_kb; ++i) {
      const int s = i % n_stages;
      const uint32_t ph = (i / n_stages) & 1;
      ptx::mbar_wait(&full_bar[s], ph);
      ptx::tc_fence_after_sync();
      const uint32_t so = static_cast<uint32_t>(s) * (L::kStageBytes >> 4);
      if (ptx::elect_one()) {
        if (dbg && i == 0) p.dbg_times[3] = clock64();
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // 4 x UMMA_K (32 bytes of K) per 128-byte stage
          const uint64_t ad =
              (static_cast<uint64_t>(hi_a) << 32) | static_cast<uint64_t>(lo_a0 + so + k * ks_a);
          const uint64_t bd =
              (static_cast<uint64_t>(hi_b) << 32) | static_cast<uint64_t>(lo_b0 + so + k * ks_b);
          const uint32_t acc = (i > 0 || k > 0) ? 1u : 0u;
          if (fp8)
            ptx::umma_f8(tmem_base, ad, bd, idesc, acc);
          else
            ptx::umma_f16(tmem_base, ad, bd, idesc, acc);
        }
        ptx::umma_commit(&empty_bar[s]);  // frees the smem slot when these MMAs retire
      }
      __syncwarp();
    }
    if (ptx::elect_one()) {
      ptx::umma_commit(accum_bar);  // accumulator complete
      if (dbg) p.dbg_times[4] = clock64();
    }
    __syncwarp();
  } else {

    // --------------------------------------------------------------- epilogue
    // TMEM -> registers (thread = one accumulator row) -> fused math -> per-warp smem staging
    // tile [32 rows][36 floats] -> global, so that every global instruction touches full
    // 128-byte lines (4 rows x 128 B per warp instruction) instead of 32 scattered rows.
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int ew = warp - 2; // epilogue warp index 0..3 (staging buffer owner)
    const int row_base = m0 + q * 32;
    const int row = row_base + lane;
    const bool row_ok = row < p.M;
    float* stg = stage_base + ew * (32 * kStgLd);
    const float* bias = p.dyn ? p.dyn->bias[bidx] : (p.bias_ptrs ? p.bias_ptrs[bidx] : p.bias);
    {  // bias tile -> smem once per CTA (coalesced), shared by the four epilogue warps
      const int et = threadIdx.x - 64;
      for (int i = et; i < BN; i += 128)
        sbias[i] = (bias != nullptr && n0 + i < p.N && split == 0) ? __ldg(bias + n0 + i) : 0.f;
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    ptx::mbar_wait(accum_bar, 0);
    ptx::tc_fence_after_sync();
    if (dbg && warp == 2 && lane == 0) p.dbg_times[5] = clock64();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const bool have_acc = n_kb > 0;
    const long long tile_off = static_cast<long long>(bidx) * p.d_batch_stride;
    // coalesced-phase coordinates of this lane: 4 rows x 8 column groups per instruction
    const int cr = lane >> 3, cg = (lane & 7) * 4;

    if constexpr (EPI == 0) {
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        const int nc = n0 + c * 32;
        if (nc >= p.N) break;  // warp-uniform
        uint32_t r[32];
        ptx::tmem_ld_32x32b_x32(taddr + c * 32, r);
        ptx::tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j)
          v[j] = (have_acc ? __uint_as_float(r[j]) * p.alpha : 0.f) + sbias[c * 32 + j];
        if (p.aux_out != nullptr) {
          stage_put(stg, lane, v);
          __syncwarp();
          tile_store<1, false>(stg, p.aux_out, tile_off, p.ldd, row_base, nc, p.M, p.N, cr, cg,
                               p.vec_ok);
          __syncwarp();
        }
        if (p.act == 1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        } else if (p.act == 2) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = gelu_f(v[j]);
        }
        if (p.act_bwd != 0) {
          tile_load_bf16(stg, p.aux_in, tile_off, p.ldd, row_base, nc, p.M, p.N, cr, cg, p.vec_ok);
          __syncwarp();
          float a[32];
          stage_get(stg, lane, a);
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 32; ++j)
            v[j] = (p.act_bwd == 1) ? (a[j] > 0.f ? v[j] : 0.f) : v[j] * gelu_grad_f(a[j]);
        }
        stage_put(stg, lane, v);
        __syncwarp();
        if (p.split_k > 1)
          tile_store<0, true>(stg, p.d, tile_off, p.ldd, row_base, nc, p.M, p.N, cr, cg, p.vec_ok);
        else if (p.d_dtype == 0)
          // d += tile: one writer per element, so a fire-and-forget red.add is deterministic and
          // has no load round trip (the read-modify-write form spent 32 dependent L2 latencies,
          // 16 us, in a 128 x 128 weight-gradient tile)
          (p.accumulate ? tile_store<0, true>(stg, p.d, tile_off, p.ldd, row_base, nc, p.M, p.N, cr,
                                              cg, p.vec_ok)
                        : tile_store<0, false>(stg, p.d, tile_off, p.ldd, row_base, nc, p.M, p.N,
                                               cr, cg, p.vec_ok));
        else if (p.d_dtype == 1)
          tile_store<1, false>(stg, p.d, tile_off, p.ldd, row_base, nc, p.M, p.N, cr, cg, p.vec_ok);
        else
          tile_store<2, false>(stg, p.d, tile_off, p.ldd, row_base, nc, p.M, p.N, cr, cg, p.vec_ok);
        if (p.colsum != nullptr) {
          // column sums straight from the staged tile: lane j adds up column j over valid rows
          const float tot = col_sum32(stg, lane, min(32, p.M - row_base));
          if (nc + lane < p.N) atomicAdd(p.colsum + nc + lane, tot);
        }
        __syncwarp();
      }
    } else {
      // ---------------- row-wise epilogues: the whole logit row lives in this CTA's tile
      const int32_t label =
          (row_ok && p.labels)
              ? p.labels[static_cast<long long>(bidx) * p.labels_batch_stride + row]
              : -1;
      // pass 1: max / argmax (+ label logit)
      float vmax = -INFINITY, zlab = 0.f;
      int amax = -1;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        const int nc = c * 32;
        if (nc >= p.N) break;
        uint32_t r[32];
        ptx::tmem_ld_32x32b_x32(taddr + c * 32, r);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int n = nc + j;
          if (n < p.N) {
            const float x = __uint_as_float(r[j]) * p.alpha + sbias[n];
            if (x > vmax) { vmax = x; amax = n; }
            if (n == label) zlab = x;
          }
        }
      }
      const bool hit = row_ok && (amax == label);
      if constexpr (EPI == 2) {
        const unsigned cnt = __popc(__ballot_sync(0xffffffffu, hit));
        if (lane == 0 && cnt && p.correct) atomicAdd(p.correct + bidx, cnt);
      } else {
        // pass 2: sum exp
        float sum = 0.f;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          const int nc = c * 32;
          if (nc >= p.N) break;
          uint32_t r[32];
          ptx::tmem_ld_32x32b_x32(taddr + c * 32, r);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int n = nc + j;
            if (n < p.N) sum += __expf(__uint_as_float(r[j]) * p.alpha + sbias[n] - vmax);
          }
        }
        const float inv = 1.f / sum;
        float loss = row_ok ? (__logf(sum) + vmax - zlab) : 0.f;
        // pass 3: dlogits (rows padded to ldd >= round_up(N, 8); pad columns get zeros)
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          const int nc = c * 32;
          if (nc >= p.N) break;
          uint32_t r[32];
          ptx::tmem_ld_32x32b_x32(taddr + c * 32, r);
          ptx::tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int n = nc + j;
            float g = 0.f;
            if (n < p.N && row_ok) {
              const float x = __uint_as_float(r[j]) * p.alpha + sbias[n];
              g = (__expf(x - vmax) * inv - (n == label ? 1.f : 0.f)) * p.grad_scale;
            }
            v[j] = g;
          }
          stage_put(stg, lane, v);
          __syncwarp();
          if (p.d != nullptr)
            tile_store<1, false>(stg, p.d, tile_off, p.ldd, row_base, nc, p.M,
                                 static_cast<int>(p.ldd), cr, cg, p.vec_ok);
          if (p.colsum != nullptr) {
            const float tot = col_sum32(stg, lane, 32);
            if (nc + lane < p.N) atomicAdd(p.colsum + nc + lane, tot);
          }
          __syncwarp();
        }
        // loss / correct reductions
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, off);
        const unsigned cnt = __popc(__ballot_sync(0xffffffffu, hit));
        if (lane == 0) {
          if (p.loss_sum) atomicAdd(p.loss_sum, loss);
          if (p.correct && cnt) atomicAdd(p.correct + bidx, cnt);
        }
      }
    }
    if (dbg && warp == 2 && lane == 0) p.dbg_times[6] = clock64();
    ptx::tc_fence_before_sync();
  }
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, BN);
    if (dbg && lane == 0) p.dbg_times[7] = clock64();
  }
}

// ------------------------------------------------------------------ host side
using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                              const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                              const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn get_encode() {
  static EncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(sym);
  });
  return fn;
}

// rows_tile: number of M|N rows the CTA tile covers (128 for A, BN for B)
cudaError_t make_map(CUtensorMap* out, const GemmOperand& op, DType dt, int rows_extent, int K,
                     int batch, int rows_tile) {
  EncodeFn enc = get_encode();
  if (!enc) return cudaErrorNotSupported;
  const int es = (dt == DType::FP8_E4M3) ? 1 : 2;
  const int epb = 128 / es;  // elements per 128-byte swizzle span
  const CUtensorMapDataType cdt =
      (dt == DType::FP8_E4M3) ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  cuuint64_t dims[3], strides[2];
  cuuint32_t box[3], estr[3] = {1, 1, 1};
  const bool batched = op.batch_stride != 0 && batch > 1;
  if (!op.mn_major) {
    dims[0] = static_cast<cuuint64_t>(K);
    dims[1] = static_cast<cuuint64_t>(rows_extent);
    box[0] = epb;
    box[1] = static_cast<cuuint32_t>(rows_tile);
  } else {
    dims[0] = static_cast<cuuint64_t>(rows_extent);
    dims[1] = static_cast<cuuint64_t>(K);
    box[0] = epb;
    box[1] = epb;  // BLOCK_K rows of K
  }
  dims[2] = batched ? static_cast<cuuint64_t>(batch) : 1;
  box[2] = 1;
  strides[0] = static_cast<cuuint64_t>(op.ld) * es;
  strides[1] = batched ? static_cast<cuuint64_t>(op.batch_stride) * es
                       : strides[0] * dims[1];
  if ((strides[0] & 15) || (strides[1] & 15) || (reinterpret_cast<uintptr_t>(op.ptr) & 15))
    return cudaErrorMisalignedAddress;
  CUresult r = enc(out, cdt, 3, const_cast<void*>(op.ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    std::fprintf(stderr,
                 "[bflc] cuTensorMapEncodeTiled failed (CUresult %d): ptr=%p mn=%d dims={%llu,%llu,%llu} "
                 "strides={%llu,%llu} box={%u,%u,%u} es=%d\n",
                 static_cast<int>(r), op.ptr, op.mn_major ? 1 : 0,
                 static_cast<unsigned long long>(dims[0]), static_cast<unsigned long long>(dims[1]),
                 static_cast<unsigned long long>(dims[2]), static_cast<unsigned long long>(strides[0]),
                 static_cast<unsigned long long>(strides[1]), box[0], box[1], box[2], es);
    return cudaErrorInvalidValue;
  }
  return cudaSuccess;
}

// Pixel tile of `pix` output pixels as whole image rows: (pw, ph, pn) with pw*ph*pn == pix.
bool conv_pixel_tile(int pix, int OH, int OW, int* pw, int* ph, int* pn) {
  if (OW <= 0 || OH <= 0 || OW > pix || pix % OW != 0) return false;
  const int rows = pix / OW;
  if (rows <= OH) {
    if (OH % rows != 0) return false;
    *pw = OW; *ph = rows; *pn = 1;
  } else {
    if (rows % OH != 0) return false;
    *pw = OW; *ph = OH; *pn = rows / OH;
  }
  return true;
}

// 4-D map over an NHWC bf16 activation whose box is `pix` pixels x 64 channels, traversed with
// the convolution stride (elementStrides) so consecutive box rows are consecutive output pixels.
cudaError_t make_conv_map(CUtensorMap* out, const ConvView& cv, int pix) {
  EncodeFn enc = get_encode();
  if (!enc) return cudaErrorNotSupported;
  int pw, ph, pn;
  if (!conv_pixel_tile(pix, cv.OH, cv.OW, &pw, &ph, &pn)) return cudaErrorInvalidValue;
  if (cv.C % 64 != 0 || cv.stride < 1 || pw * cv.stride > 256 || ph * cv.stride > 256)
    return cudaErrorInvalidValue;
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(cv.C), static_cast<cuuint64_t>(cv.W),
                        static_cast<cuuint64_t>(cv.H), static_cast<cuuint64_t>(cv.N)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(cv.C) * 2,
                           static_cast<cuuint64_t>(cv.W) * cv.C * 2,
                           static_cast<cuuint64_t>(cv.H) * cv.W * cv.C * 2};
  cuuint32_t box[4] = {64, static_cast<cuuint32_t>(pw * cv.stride),
                       static_cast<cuuint32_t>(ph * cv.stride), static_cast<cuuint32_t>(pn)};
  cuuint32_t estr[4] = {1, static_cast<cuuint32_t>(cv.stride), static_cast<cuuint32_t>(cv.stride), 1};
  if (reinterpret_cast<uintptr_t>(cv.x) & 15) return cudaErrorMisalignedAddress;
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(cv.x), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    std::fprintf(stderr, "[bflc] conv tensor map failed (CUresult %d): NHWC=%d,%d,%d,%d box=%u,%u,%u,%u stride=%d\n",
                 static_cast<int>(r), cv.N, cv.H, cv.W, cv.C, box[0], box[1], box[2], box[3], cv.stride);
    return cudaErrorInvalidValue;
  }
  return cudaSuccess;
}

template <int BN, int EPI>
cudaError_t launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc,
                   const KParams& kp, dim3 grid, cudaStream_t stream) {
  using L = SmemLayout<BN>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_kernel<BN, EPI>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  note_launch();
  const int smem = L::kTotal - (L::kStages - kp.stages) * L::kStageBytes;
  return launch_pdl(gemm_kernel<BN, EPI>, grid, dim3(kThreads), smem, stream, ta, tb, tc, kp);
}

}  // namespace

static unsigned long long g_launches = 0;
unsigned long long launch_count() { return g_launches; }
void note_launch() { ++g_launches; }
static thread_local long long* g_dbg_times = nullptr;
void set_debug_times(long long* p) { g_dbg_times = p; }
static bool g_pdl = true;
static unsigned long long g_pdl_fallbacks = 0;
void note_pdl_fallback() { ++g_pdl_fallbacks; }
unsigned long long pdl_fallbacks() { return g_pdl_fallbacks; }
void set_pdl(bool on) { g_pdl = on; }
bool pdl_enabled() { return g_pdl; }
static thread_local const int* g_pred = nullptr;
void set_predicate(const int* pred) { g_pred = pred; }
const int* current_predicate() { return g_pred; }

cudaError_t gemm_make_operand_map(CUtensorMap* out, const GemmOperand& op, DType dt,
                                  int rows_extent, int K, int batch, int rows_tile) {
  return make_map(out, op, dt, rows_extent, K, batch, rows_tile);
}

int gemm_pick_bn(int N, EpiKind kind, int M, int z) {
  if (kind != EpiKind::GENERIC) return N <= 64 ? 64 : (N <= 128 ? 128 : 256);
  // Largest tile that still yields ~a wave of CTAs; tiny problems take the narrowest tile so
  // the fixed per-CTA latency (setup + first TMA + epilogue) is spread over more SMs.
  const int mt = (M + kBM - 1) / kBM;
  const int cand[3] = {256, 128, 64};
  for (int i = 0; i < 3; ++i) {
    const int bn = cand[i];
    if (bn > 64 && N <= bn / 2) continue;
    const long long ctas = static_cast<long long>((N + bn - 1) / bn) * mt * (z < 1 ? 1 : z);
    if (ctas >= 120 || bn == 64) return bn;
  }
  return 64;
}

cudaError_t gemm_make_b_map(const GemmProblem& p, CUtensorMap* out_host) {
  bind_context_once();
  const int BN = p.force_bn ? p.force_bn : gemm_pick_bn(p.N, p.epi.kind, p.M, p.batch);
  return make_map(out_host, p.b, p.ab_dtype, p.N, p.K, p.batch, BN);
}

cudaError_t gemm_sm100(const GemmProblem& p, cudaStream_t stream) {
  bind_context_once();
  if (p.M <= 0 || p.N <= 0 || p.K <= 0 || p.batch <= 0) return cudaErrorInvalidValue;
  const bool fp8 = p.ab_dtype == DType::FP8_E4M3;
  if (p.ab_dtype == DType::F32) return cudaErrorInvalidValue;
  int BN = p.force_bn ? p.force_bn
                      : gemm_pick_bn(p.N, p.epi.kind, p.M,
                                     p.batch * (p.epi.split_k < 1 ? 1 : p.epi.split_k));
  if (BN != 64 && BN != 128 && BN != 256) return cudaErrorInvalidValue;
  if (p.epi.kind != EpiKind::GENERIC && (p.N > 256 || BN < p.N)) return cudaErrorInvalidValue;
  if (fp8 && p.b.mn_major && BN < 128) BN = 128;
  const int block_k = fp8 ? 128 : 64;

  CUtensorMap ta, tb, tc;
  std::memset(&tc, 0, sizeof(tc));
  cudaError_t e = cudaSuccess;
  const ConvView& cv = p.conv;
  const int taps = cv.KH * cv.KW;
  if (cv.mode != 0) {
    if (fp8 || p.batch != 1 || p.b_maps_dev || p.dyn || p.epi.kind != EpiKind::GENERIC)
      return cudaErrorInvalidValue;
    if (cv.mode == 1) {
      // rows = pixels of the (OH, OW) grid; K = taps x C
      if (p.a.mn_major || p.K != taps * cv.C || p.M != cv.N * cv.OH * cv.OW) return cudaErrorInvalidValue;
      if (cv.flip && (cv.stride != 1 || !p.b.mn_major || p.N % 64 != 0)) return cudaErrorInvalidValue;
      if (!cv.flip && p.b.mn_major) return cudaErrorInvalidValue;
      if (cv.flip && BN > p.N) BN = (p.N % 128 == 0) ? 128 : 64;
      if (cv.flip && p.N % BN != 0) BN = 64;
      e = make_conv_map(&tc, cv, kBM);
      if (e != cudaSuccess) return e;
      std::memset(&ta, 0, sizeof(ta));
      e = cv.flip ? make_map(&tb, p.b, p.ab_dtype, taps * p.N, cv.C, 1, BN)
                  : make_map(&tb, p.b, p.ab_dtype, p.N, p.K, 1, BN);
      if (e != cudaSuccess) return e;
    } else {
      // D = dW [Cout][taps x C]; reduction over the pixels
      if (!p.a.mn_major || p.N != taps * cv.C || p.K != cv.N * cv.OH * cv.OW) return cudaErrorInvalidValue;
      BN = (cv.C % 256 == 0 && BN == 256) ? 256 : ((cv.C % 128 == 0 && BN >= 128) ? 128 : 64);
      e = make_conv_map(&tc, cv, 64);
      if (e != cudaSuccess) return e;
      e = make_map(&ta, p.a, p.ab_dtype, p.M, p.K, 1, kBM);
      if (e != cudaSuccess) return e;
      std::memset(&tb, 0, sizeof(tb));
    }
  } else {
    e = make_map(&ta, p.a, p.ab_dtype, p.M, p.K, p.batch, kBM);
    if (e != cudaSuccess) return e;
    if (p.b_maps_dev == nullptr) {
      e = make_map(&tb, p.b, p.ab_dtype, p.N, p.K, p.batch, BN);
      if (e != cudaSuccess) return e;
    } else {
      std::memset(&tb, 0, sizeof(tb));
    }
  }

  KParams kp{};
  kp.M = p.M; kp.N = p.N; kp.K = p.K; kp.batch = p.batch;
  kp.is_fp8 = fp8 ? 1 : 0;
  kp.a_mn = p.a.mn_major ? 1 : 0;
  kp.b_mn = p.b.mn_major ? 1 : 0;
  kp.a_batched = (p.a.batch_stride != 0 && p.batch > 1) ? 1 : 0;
  kp.b_batched = (p.b.batch_stride != 0 && p.batch > 1) ? 1 : 0;
  kp.k_blocks = (p.K + block_k - 1) / block_k;
  kp.split_k = p.epi.split_k < 1 ? 1 : p.epi.split_k;
  if (kp.split_k > kp.k_blocks) kp.split_k = kp.k_blocks;
  if (kp.split_k > 1 && (p.epi.kind != EpiKind::GENERIC || p.epi.d_dtype != DType::F32))
    return cudaErrorInvalidValue;
  kp.b_maps_dev = p.b_maps_dev;
  kp.dyn = p.dyn;
  kp.pred = current_predicate();
  kp.dbg_times = g_dbg_times;
  kp.d = p.epi.d;
  kp.d_dtype = static_cast<int>(p.epi.d_dtype);
  kp.ldd = p.epi.ldd;
  kp.d_batch_stride = p.epi.d_batch_stride;
  kp.alpha = p.epi.alpha;
  kp.bias = p.epi.bias;
  kp.bias_ptrs = p.epi.bias_ptrs;
  kp.act = static_cast<int>(p.epi.act);
  kp.aux_out = p.epi.aux_out;
  kp.aux_in = p.epi.aux_in;
  kp.act_bwd = p.epi.act_bwd;
  kp.colsum = p.epi.colsum;
  kp.accumulate = p.epi.accumulate;
  kp.labels = p.epi.labels;
  kp.labels_batch_stride = p.epi.labels_batch_stride;
  kp.grad_scale = p.epi.grad_scale;
  kp.loss_sum = p.epi.loss_sum;
  kp.correct = p.epi.correct;
  {
    // vector (16 B fp32 / 8 B bf16) global access needs 4-element aligned rows everywhere
    auto al = [](const void* q, int bytes) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) % bytes) == 0; };
    const int eb = p.epi.d_dtype == DType::F32 ? 16 : (p.epi.d_dtype == DType::BF16 ? 8 : 4);
    kp.vec_ok = (p.epi.ldd % 4 == 0) && (p.epi.d_batch_stride % 4 == 0) && al(p.epi.d, eb) &&
                al(p.epi.aux_out, 8) && al(p.epi.aux_in, 8);
  }
  // Canonical SWIZZLE_128B descriptors.
  //  K-major : rows of 128 B; 8-row groups every 1024 B (SBO); LBO unused (1 unit).
  //            one UMMA_K step = 32 B inside the swizzle span.
  //  MN-major: each TMA box is [block_k K-rows][128 B of MN]; 8-row K groups every 1024 B
  //            (SBO); the next 128-B MN chunk is the next box, block_k*128 B away (LBO);
  //            one UMMA_K step = (32 / es) K-rows = 32/es * 128 B.
  const uint32_t mn_kstep = (fp8 ? 32u : 16u) * 128u;
  kp.lbo_a = p.dbg_lbo_a ? p.dbg_lbo_a : (p.a.mn_major ? static_cast<uint32_t>(block_k) * 128u : 16u);
  kp.sbo_a = p.dbg_sbo_a ? p.dbg_sbo_a : 1024u;
  kp.lbo_b = p.dbg_lbo_b ? p.dbg_lbo_b : (p.b.mn_major ? static_cast<uint32_t>(block_k) * 128u : 16u);
  kp.sbo_b = p.dbg_sbo_b ? p.dbg_sbo_b : 1024u;
  kp.kstep_a = p.a.mn_major ? mn_kstep : 32u;
  kp.kstep_b = p.b.mn_major ? mn_kstep : 32u;
  kp.cv_mode = cv.mode; kp.cv_flip = cv.flip;
  kp.cv_cb = cv.C / 64; kp.cv_kw = cv.KW; kp.cv_pad = cv.pad; kp.cv_stride = cv.stride;
  kp.cv_oh = cv.OH; kp.cv_ow = cv.OW; kp.cv_c = cv.C;
  if (cv.mode == 2) kp.b_mn = 1;  // shifted activation boxes are [pixels][channels]: MN-major B
  if (cv.mode == 2) { kp.lbo_b = 64u * 128u; kp.kstep_b = mn_kstep; }

  dim3 grid((p.N + BN - 1) / BN, (p.M + kBM - 1) / kBM, p.batch * kp.split_k);
  {
    // Ring depth: the full ring for long reductions; 3 stages (2 CTAs per SM with the 64-wide
    // tile) when every CTA only runs a few K blocks and there is more than a wave of tiles.
    const int full = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
    const int kb_cta = (kp.k_blocks + kp.split_k - 1) / kp.split_k;
    const long long ctas = static_cast<long long>(grid.x) * grid.y * grid.z;
    static const int force = [] { const char* e = std::getenv("BFLC_GEMM_STAGES"); return e ? std::atoi(e) : 0; }();
    kp.stages = full;
    if (BN == 64 && kb_cta <= 40 && ctas > 148) kp.stages = 3;
    if (force >= 2 && force <= full) kp.stages = force;
  }
#define BFLC_LAUNCH(BN_, EPI_) return launch<BN_, EPI_>(ta, tb, tc, kp, grid, stream)
  const int epi = static_cast<int>(p.epi.kind);
  if (epi == 0) {
    if (BN == 64) BFLC_LAUNCH(64, 0);
    if (BN == 128) BFLC_LAUNCH(128, 0);
    BFLC_LAUNCH(256, 0);
  } else if (epi == 1) {
    if (BN == 64) BFLC_LAUNCH(64, 1);
    if (BN == 128) BFLC_LAUNCH(128, 1);
    BFLC_LAUNCH(256, 1);
  } else {
    if (BN == 64) BFLC_LAUNCH(64, 2);
    if (BN == 128) BFLC_LAUNCH(128, 2);
    BFLC_LAUNCH(256, 2);
  }
#undef BFLC_LAUNCH
}

}  // namespace bflc
