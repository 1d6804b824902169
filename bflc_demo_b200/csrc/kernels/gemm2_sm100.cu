// 2-CTA tcgen05 GEMM (cta_group::2) for large K-major problems -- persistent, with the TMEM
// accumulator double-buffered so the epilogue of tile i drains under the mainloop of tile i+1.
//
// 74 CTA pairs (one per TPC, 148 SMs) walk the tile list with a static stride; the TMA ring, its
// mbarriers and the 512-column TMEM allocation (2 x 256 accumulator columns) live for the whole
// kernel.  Hand-offs: `tmem_full[b]` (MMA commit, multicast to both CTAs) tells the epilogue warps
// that accumulator b is complete; `tmem_empty[b]` on the LEADER collects one arrive per epilogue
// warp of BOTH CTAs (the peer's warps arrive remotely through the cluster address) before the
// leader's MMA thread may overwrite buffer b.  Before this change every tile paid a launch slot,
// a TMEM allocation, barrier init and an exposed epilogue: 577 TFLOP/s at 16384 x 1024 x 1024
// against cuBLAS' 1078 (16 K-blocks per tile cannot hide any of it).
//
// A CTA pair (cluster 2x1, both SMs of one TPC) computes one 256 x 256 output tile:
// each CTA stages its own 128 rows of A and HALF of the B tile (128 of the 256 N-rows) per
// K-block, the leader CTA's elected thread issues `tcgen05.mma.cta_group::2` (UMMA M = 256)
// which reads both CTAs' shared memory, and each CTA ends up with its 128 x 256 slice of the
// accumulator in its own TMEM.  Per SM and K-block this moves 32 KB instead of the 48 KB of the
// 1-CTA 128 x 256 tile (gemm_sm100.cu), which is what matters: that kernel is pinned at the
// L2 -> SM feed rate (~1.15 PFLOP/s, profiles/), not at the tensor pipe.
//
//   warp 0 (both CTAs)  TMA producer: cp.async.bulk.tensor ... cta_group::2, completing on the
//                        LEADER's full barrier (peer bit masked off the mbarrier address)
//   warp 1 (leader)      waits full, issues 4 x UMMA per K-block, tcgen05.commit multicast to
//                        both CTAs' empty barriers; (both) TMEM alloc/dealloc with cta_group::2
//   warps 2-5 (both)     epilogue of the CTA's own 128 rows (staged, coalesced stores)
#include <cuda_bf16.h>

#include <cstring>

#include "bflc_kernels.h"
#include "epi_common.cuh"
#include "sm100_ptx.cuh"

namespace bflc {

namespace {

constexpr int kBM = 128;           // rows per CTA (UMMA M = 256 across the pair)
constexpr int kBN = 256;           // tile N (each CTA stages kBN/2 rows of B)
constexpr int kStages = 6;
constexpr int kABytes = kBM * 128, kBBytes = (kBN / 2) * 128, kStageBytes = kABytes + kBBytes;
constexpr int kTileBytes = kStages * kStageBytes;
constexpr int kBarBytes = 256;
using epi::kStgLd;
using epi::kStgBytes;
constexpr int kSmemTotal = kTileBytes + kBarBytes + kStgBytes + 2 * kBN * 4 + 1024;
constexpr int kAccBufs = 2;
constexpr int kThreads = 192;
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address

struct P2 {
  int M, N, K, k_blocks;
  int m_pairs, n_tiles;      // tile grid: 256-row pair tiles x 256-column tiles (M fastest)
  void* d; int d_dtype; long long ldd; float alpha;
  const float* bias; int act;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(void* smem_dst, const void* tmap, uint32_t mbar_addr,
                                                int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5}], [%2];"
      :
      : "r"(ptx::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(mbar_addr), "r"(c0),
        "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      :
      : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(ptx::smem_u32(bar)), "h"(static_cast<uint16_t>(3))
      : "memory");
}
__device__ __forceinline__ uint32_t pack2(float a, float b) { return epi::pack_bf16x2(a, b); }
__device__ __forceinline__ float gelu_f(float x) {
  return 0.5f * x * (1.f + erff(x * 0.70710678118654752f));
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const P2 p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kTileBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;           // [kAccBufs] accumulator b complete (both CTAs)
  uint64_t* tmem_empty = tmem_full + kAccBufs;         // [kAccBufs] leader only: accumulator b drained by all 8 warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + kAccBufs);
  float* stage_base = reinterpret_cast<float*>(smem + kTileBytes + kBarBytes);
  float* sbias = stage_base + 4 * 32 * kStgLd;         // [2][kBN]: double-buffered with the accumulator

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int n_clusters = gridDim.x >> 1;
  const int cluster = blockIdx.x >> 1;
  const int n_tiles_total = p.m_pairs * p.n_tiles;

  if (warp == 0 && lane == 0) {
    ptx::tma_prefetch_desc(&tmA);
    ptx::tma_prefetch_desc(&tmB);
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(&full_bar[s], 2);   // leader's expect_tx arrive + the peer's remote arrive
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < kAccBufs; ++b) {
      ptx::mbar_init(&tmem_full[b], 1);
      ptx::mbar_init(&tmem_empty[b], 8);   // 4 epilogue warps x 2 CTAs
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     ptx::smem_u32(tmem_slot)), "r"(kAccBufs * kBN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  ptx::tc_fence_before_sync();
  cluster_sync_all();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const int n_kb = p.k_blocks;

  if (warp == 0) {
    uint32_t it = 0;
    for (int t = cluster; t < n_tiles_total; t += n_clusters) {
      const int m0 = (t % p.m_pairs) * (2 * kBM) + static_cast<int>(rank) * kBM;
      const int n0 = (t / p.m_pairs) * kBN;
      for (int i = 0; i < n_kb; ++i, ++it) {
        const int s = it % kStages;
        const uint32_t ph = (it / kStages) & 1;
        ptx::mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* sa = smem + s * kStageBytes;
        uint8_t* sb = sa + kABytes;
        const uint32_t full_leader = ptx::smem_u32(&full_bar[s]) & kPeerMask;
        if (ptx::elect_one()) {
          if (leader) ptx::mbar_expect_tx(&full_bar[s], 2 * kStageBytes);
          else mbar_arrive_cluster(full_leader);
          tma_load_3d_2sm(sa, &tmA, full_leader, i * 64, m0, 0);
          tma_load_3d_2sm(sb, &tmB, full_leader, i * 64, n0 + static_cast<int>(rank) * (kBN / 2), 0);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // instruction descriptor: bf16 x bf16 -> f32, K-major A and B, M = 256 (pair), N = 256
      const uint32_t idesc = ptx::make_idesc(1u, 0u, 0u, 2 * kBM, kBN);
      const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
      const uint32_t base_lo = ptx::smem_u32(smem) >> 4;
      const uint32_t lo_a0 = base_lo | (1u << 16);
      const uint32_t lo_b0 = (base_lo + (kABytes >> 4)) | (1u << 16);
      uint32_t it = 0, tile = 0;
      for (int t = cluster; t < n_tiles_total; t += n_clusters, ++tile) {
        const uint32_t b = tile & 1u, use = tile >> 1;
        // buffer b was drained by the epilogue of tile - 2 (first two tiles: free)
        ptx::mbar_wait(&tmem_empty[b], (use & 1u) ^ 1u);
        ptx::tc_fence_after_sync();
        const uint32_t tacc = tmem_base + b * kBN;
        for (int i = 0; i < n_kb; ++i, ++it) {
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          ptx::mbar_wait(&full_bar[s], ph);
          ptx::tc_fence_after_sync();
          const uint32_t so = static_cast<uint32_t>(s) * (kStageBytes >> 4);
          if (ptx::elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t ad = (static_cast<uint64_t>(hi) << 32) | (lo_a0 + so + k * 2u);
              const uint64_t bd = (static_cast<uint64_t>(hi) << 32) | (lo_b0 + so + k * 2u);
              umma2_f16(tacc, ad, bd, idesc, (i > 0 || k > 0) ? 1u : 0u);
            }
            umma2_commit_mc(&empty_bar[s]);
          }
          __syncwarp();
        }
        if (ptx::elect_one()) umma2_commit_mc(&tmem_full[b]);
        __syncwarp();
      }
    }
  } else {
    const int q = warp & 3;
    float* stg = stage_base + (warp - 2) * (32 * kStgLd);
    const int cr = lane >> 3, cg = (lane & 7) * 4;
    const uint32_t empty_leader = ptx::smem_u32(&tmem_empty[0]) & kPeerMask;
    uint32_t tile = 0;
    for (int t = cluster; t < n_tiles_total; t += n_clusters, ++tile) {
      const uint32_t b = tile & 1u, use = tile >> 1;
      const int m0 = (t % p.m_pairs) * (2 * kBM) + static_cast<int>(rank) * kBM;
      const int n0 = (t / p.m_pairs) * kBN;
      const int row_base = m0 + q * 32;
      float* sb = sbias + b * kBN;
      {
        // bias of this tile's columns (buffer b of sbias was last read two tiles ago: every
        // epilogue warp has since passed a bar.sync of tile - 1)
        const int et = threadIdx.x - 64;
        for (int i = et; i < kBN; i += 128)
          sb[i] = (p.bias != nullptr && n0 + i < p.N) ? __ldg(p.bias + n0 + i) : 0.f;
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      ptx::mbar_wait(&tmem_full[b], use & 1u);
      ptx::tc_fence_after_sync();
      const uint32_t taddr = tmem_base + b * kBN + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < kBN / 32; ++c) {
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
            float x = __uint_as_float(r[4 * j + k]) * p.alpha + sb[c * 32 + 4 * j + k];
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
      // this warp's quarter of accumulator b is in registers / memory: hand the buffer back to
      // the leader's MMA thread (one arrive per warp, remote for the peer CTA)
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(empty_leader + b * 8u);
    }
  }
  // both CTAs must be done with the pair's TMEM / smem before either tears down
  ptx::tc_fence_before_sync();
  cluster_sync_all();
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kAccBufs * kBN)
                 : "memory");
  }
}

}  // namespace

// D (M x N) = alpha * A (M x K, K-major) . B^T (N x K, K-major) [+ bias] [act]; ld % 4 == 0.
cudaError_t gemm2_sm100(const GemmProblem& p, cudaStream_t stream) {
  bind_context_once();
  if (p.ab_dtype != DType::BF16 || p.a.mn_major || p.b.mn_major || p.batch != 1 ||
      p.epi.kind != EpiKind::GENERIC || p.epi.split_k > 1 || p.epi.aux_in || p.epi.aux_out ||
      p.epi.colsum || p.epi.accumulate || p.epi.ldd % 4 != 0 || p.epi.d_dtype == DType::FP8_E4M3)
    return cudaErrorNotSupported;
  CUtensorMap ta, tb;
  cudaError_t e = gemm_make_operand_map(&ta, p.a, DType::BF16, p.M, p.K, 1, kBM);
  if (e != cudaSuccess) return e;
  e = gemm_make_operand_map(&tb, p.b, DType::BF16, p.N, p.K, 1, kBN / 2);
  if (e != cudaSuccess) return e;
  P2 kp{};
  kp.M = p.M; kp.N = p.N; kp.K = p.K; kp.k_blocks = (p.K + 63) / 64;
  kp.m_pairs = (p.M + 2 * kBM - 1) / (2 * kBM); kp.n_tiles = (p.N + kBN - 1) / kBN;
  kp.d = p.epi.d; kp.d_dtype = static_cast<int>(p.epi.d_dtype); kp.ldd = p.epi.ldd;
  kp.alpha = p.epi.alpha; kp.bias = p.epi.bias; kp.act = static_cast<int>(p.epi.act);
  static bool configured = false;
  if (!configured) {
    e = cudaFuncSetAttribute(gemm2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  // persistent: one CTA pair per TPC (74 on a B200), M fastest so that pairs running at the same
  // time share B tiles in L2
  const int tiles = kp.m_pairs * kp.n_tiles;
  const int pairs = tiles < 74 ? tiles : 74;
  dim3 grid(2 * pairs, 1, 1);
  (void)cudaGetLastError();
  gemm2_kernel<<<grid, kThreads, kSmemTotal, stream>>>(ta, tb, kp);
  note_launch();
  return cudaGetLastError();
}

}  // namespace bflc
