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
#include <cstring>
#include <cuda_bf16.h>
#include <cuda_fp8.h>
#include <mutex>

#include "bflc_kernels.h"
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
};

template <int BN>
struct SmemLayout {
  static constexpr int kBBytes = BN * kStageKBytes;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int kTileBytes = kStages * kStageBytes;
  static constexpr int kBarBytes = 256;
  static constexpr int kTotal = kTileBytes + kBarBytes + 1024;  // + alignment slack
};

__device__ __forceinline__ float gelu_f(float x) {
  return 0.5f * x * (1.f + erff(x * 0.70710678118654752f));
}
__device__ __forceinline__ float gelu_grad_f(float x) {
  const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752f));
  const float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// Sum v[j] over the 32 lanes of the warp for each j; lane j returns column j's total.
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int j = 0; j < off; ++j) {
      // keep half of the columns: lanes with bit `off` set keep [off, 2*off), others [0, off)
      const float send = upper ? v[j] : v[j + off];
      const float recv = __shfl_xor_sync(0xffffffffu, send, off);
      v[j] = (upper ? v[j + off] : v[j]) + recv;
    }
  }
  return v[0];
}

template <int BN, int EPI>
__global__ void __launch_bounds__(kThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const KParams p) {
  using L = SmemLayout<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kTileBytes);
  uint64_t* empty_bar = full_bar + L::kStages;
  uint64_t* accum_bar = empty_bar + L::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  // role predication / dynamic batch count: CTA-uniform early exits before any barrier
  if (p.pred != nullptr && *p.pred == 0) return;
  if (p.dyn != nullptr && static_cast<int>(blockIdx.z) / p.split_k >= p.dyn->active_batches)
    return;

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

  const CUtensorMap* mapB =
      p.b_maps_dev ? (p.b_maps_dev + (p.dyn ? p.dyn->map_index[bidx] : bidx)) : &tmB;

  if (warp == 0 && lane == 0) {
    ptx::tma_prefetch_desc(&tmA);
    ptx::tma_prefetch_desc(mapB);
    for (int s = 0; s < L::kStages; ++s) {
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

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      const int ca2 = p.a_batched ? bidx : 0;
      const int cb2 = (p.b_batched && !p.b_maps_dev) ? bidx : 0;
      const uint32_t stage_bytes = L::kStageBytes;
      if (p.dyn != nullptr && p.dyn->wait_flag[bidx] != nullptr) {
        // B lives in a peer's upload buffer: wait until that trainer released it, then make
        // the acquired state visible to the async proxy before the first TMA pull.
        ptx::wait_flag_ge(p.dyn->wait_flag[bidx], p.dyn->wait_value);
        ptx::fence_proxy_async_all();
      }
      for (int i = 0; i < n_kb; ++i) {
        const int s = i % L::kStages;
        const uint32_t ph = (i / L::kStages) & 1;
        ptx::mbar_wait(&empty_bar[s], ph ^ 1);
        ptx::mbar_expect_tx(&full_bar[s], stage_bytes);
        uint8_t* sa = smem + s * L::kStageBytes;
        uint8_t* sb = sa + kABytes;
        const int k0 = (kb_begin + i) * block_k;
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
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      const uint32_t idesc =
          ptx::make_idesc(p.is_fp8 ? 0u : 1u, p.a_mn ? 1u : 0u, p.b_mn ? 1u : 0u, kBM, BN);
      for (int i = 0; i < n_kb; ++i) {
        const int s = i % L::kStages;
        const uint32_t ph = (i / L::kStages) & 1;
        ptx::mbar_wait(&full_bar[s], ph);
        ptx::tc_fence_after_sync();
        if (dbg && i == 0) p.dbg_times[3] = clock64();
        const uint32_t sa = ptx::smem_u32(smem + s * L::kStageBytes);
        const uint32_t sb = sa + kABytes;
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // 4 x UMMA_K (32 bytes of K) per 128-byte stage
          const uint64_t ad = ptx::make_smem_desc_sw128(sa + k * p.kstep_a, p.lbo_a, p.sbo_a);
          const uint64_t bd = ptx::make_smem_desc_sw128(sb + k * p.kstep_b, p.lbo_b, p.sbo_b);
          const uint32_t acc = (i > 0 || k > 0) ? 1u : 0u;
          if (p.is_fp8)
            ptx::umma_f8(tmem_base, ad, bd, idesc, acc);
          else
            ptx::umma_f16(tmem_base, ad, bd, idesc, acc);
        }
        ptx::umma_commit(&empty_bar[s]);  // frees the smem slot when these MMAs retire
      }
      ptx::umma_commit(accum_bar);  // accumulator complete
      if (dbg) p.dbg_times[4] = clock64();
    }
  } else {
    // --------------------------------------------------------------- epilogue
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int row = m0 + q * 32 + lane;
    const bool row_ok = row < p.M;
    ptx::mbar_wait(accum_bar, 0);
    ptx::tc_fence_after_sync();
    if (dbg && warp == 2 && lane == 0) p.dbg_times[5] = clock64();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const bool have_acc = n_kb > 0;

    if constexpr (EPI == 0) {
      const float* bias =
          p.dyn ? p.dyn->bias[bidx] : (p.bias_ptrs ? p.bias_ptrs[bidx] : p.bias);
      const long long row_off = static_cast<long long>(bidx) * p.d_batch_stride +
                                static_cast<long long>(row) * p.ldd;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        const int nc = n0 + c * 32;
        if (nc >= p.N) break;  // warp-uniform
        uint32_t r[32];
        ptx::tmem_ld_32x32b_x32(taddr + c * 32, r);
        ptx::tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float x = have_acc ? __uint_as_float(r[j]) * p.alpha : 0.f;
          const int n = nc + j;
          if (bias != nullptr && n < p.N && split == 0) x += __ldg(bias + n);
          v[j] = x;
        }
        const bool full = (nc + 32 <= p.N);
        if (p.aux_out != nullptr && row_ok) {
          __nv_bfloat16* ao = reinterpret_cast<__nv_bfloat16*>(p.aux_out) + row_off + nc;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (full || nc + j < p.N) ao[j] = __float2bfloat16(v[j]);
        }
        if (p.act == 1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        } else if (p.act == 2) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = gelu_f(v[j]);
        }
        if (p.act_bwd != 0) {
          const __nv_bfloat16* ai =
              reinterpret_cast<const __nv_bfloat16*>(p.aux_in) + row_off + nc;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float a = 0.f;
            if (row_ok && (full || nc + j < p.N)) a = __bfloat162float(ai[j]);
            v[j] = (p.act_bwd == 1) ? (a > 0.f ? v[j] : 0.f) : v[j] * gelu_grad_f(a);
          }
        }
        if (row_ok) {
          if (p.split_k > 1) {
            float* d = reinterpret_cast<float*>(p.d) + row_off + nc;
            if (full && ((reinterpret_cast<uintptr_t>(d) & 15) == 0)) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                ptx::red_add_f32x4(d + j, v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (full || nc + j < p.N) atomicAdd(d + j, v[j]);
            }
          } else if (p.d_dtype == 0) {
            float* d = reinterpret_cast<float*>(p.d) + row_off + nc;
            const bool vec = full && ((reinterpret_cast<uintptr_t>(d) & 15) == 0);
            if (vec && !p.accumulate) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(d + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (full || nc + j < p.N) d[j] = p.accumulate ? d[j] + v[j] : v[j];
            }
          } else if (p.d_dtype == 1) {
            __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(p.d) + row_off + nc;
            const bool vec = full && ((reinterpret_cast<uintptr_t>(d) & 15) == 0);
            if (vec) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                uint4 pk;
                __nv_bfloat162 t0 = __floats2bfloat162_rn(v[j], v[j + 1]);
                __nv_bfloat162 t1 = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
                __nv_bfloat162 t2 = __floats2bfloat162_rn(v[j + 4], v[j + 5]);
                __nv_bfloat162 t3 = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
                pk.x = *reinterpret_cast<uint32_t*>(&t0);
                pk.y = *reinterpret_cast<uint32_t*>(&t1);
                pk.z = *reinterpret_cast<uint32_t*>(&t2);
                pk.w = *reinterpret_cast<uint32_t*>(&t3);
                *reinterpret_cast<uint4*>(d + j) = pk;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (full || nc + j < p.N) d[j] = __float2bfloat16(v[j]);
            }
          } else {
            // fp8 e4m3 output (already scaled by alpha)
            __nv_fp8_e4m3* d = reinterpret_cast<__nv_fp8_e4m3*>(p.d) + row_off + nc;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (full || nc + j < p.N) d[j] = __nv_fp8_e4m3(v[j]);
          }
        }
        if (p.colsum != nullptr) {
          if (!row_ok) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.f;
          }
          const float tot = warp_colsum32(v, lane);
          if (nc + lane < p.N) atomicAdd(p.colsum + nc + lane, tot);
        }
      }
    } else {
      // ---------------- row-wise epilogues: the whole logit row lives in this CTA's tile
      static_assert(BN <= 256, "row epilogue needs N <= BN");
      const int32_t label =
          (row_ok && p.labels)
              ? p.labels[static_cast<long long>(bidx) * p.labels_batch_stride + row]
              : -1;
      const float* bias =
          p.dyn ? p.dyn->bias[bidx] : (p.bias_ptrs ? p.bias_ptrs[bidx] : p.bias);
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
            float x = __uint_as_float(r[j]) * p.alpha + (bias ? __ldg(bias + n) : 0.f);
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
            if (n < p.N) {
              float x = __uint_as_float(r[j]) * p.alpha + (bias ? __ldg(bias + n) : 0.f);
              sum += __expf(x - vmax);
            }
          }
        }
        const float inv = 1.f / sum;
        float loss = row_ok ? (__logf(sum) + vmax - zlab) : 0.f;
        // pass 3: dlogits
        const long long row_off = static_cast<long long>(bidx) * p.d_batch_stride +
                                  static_cast<long long>(row) * p.ldd;
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
              float x = __uint_as_float(r[j]) * p.alpha + (bias ? __ldg(bias + n) : 0.f);
              g = (__expf(x - vmax) * inv - (n == label ? 1.f : 0.f)) * p.grad_scale;
            }
            v[j] = g;
          }
          if (row_ok && p.d != nullptr) {
            // dlogits rows are padded to ldd (>= round_up(N, 8)); pad columns get zeros
            __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(p.d) + row_off + nc;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (nc + j < p.ldd) d[j] = __float2bfloat16(v[j]);
          }
          if (p.colsum != nullptr) {
            const float tot = warp_colsum32(v, lane);
            if (nc + lane < p.N) atomicAdd(p.colsum + nc + lane, tot);
          }
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
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

template <int BN, int EPI>
cudaError_t launch(const CUtensorMap& ta, const CUtensorMap& tb, const KParams& kp, dim3 grid,
                   cudaStream_t stream) {
  using L = SmemLayout<BN>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_kernel<BN, EPI>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  gemm_kernel<BN, EPI><<<grid, kThreads, L::kTotal, stream>>>(ta, tb, kp);
  note_launch();
  return cudaGetLastError();
}

}  // namespace

static unsigned long long g_launches = 0;
unsigned long long launch_count() { return g_launches; }
void note_launch() { ++g_launches; }
static thread_local long long* g_dbg_times = nullptr;
void set_debug_times(long long* p) { g_dbg_times = p; }
static thread_local const int* g_pred = nullptr;
void set_predicate(const int* pred) { g_pred = pred; }
const int* current_predicate() { return g_pred; }

int gemm_pick_bn(int N, EpiKind kind) {
  if (kind != EpiKind::GENERIC) return N <= 64 ? 64 : (N <= 128 ? 128 : 256);
  if (N <= 64) return 64;
  if (N <= 128) return 128;
  return 256;
}

cudaError_t gemm_make_b_map(const GemmProblem& p, CUtensorMap* out_host) {
  const int BN = gemm_pick_bn(p.N, p.epi.kind);
  return make_map(out_host, p.b, p.ab_dtype, p.N, p.K, p.batch, BN);
}

cudaError_t gemm_sm100(const GemmProblem& p, cudaStream_t stream) {
  if (p.M <= 0 || p.N <= 0 || p.K <= 0 || p.batch <= 0) return cudaErrorInvalidValue;
  const bool fp8 = p.ab_dtype == DType::FP8_E4M3;
  if (p.ab_dtype == DType::F32) return cudaErrorInvalidValue;
  const int BN = gemm_pick_bn(p.N, p.epi.kind);
  if (p.epi.kind != EpiKind::GENERIC && p.N > 256) return cudaErrorInvalidValue;
  if (fp8 && p.b.mn_major && BN < 128) return cudaErrorInvalidValue;
  const int block_k = fp8 ? 128 : 64;

  CUtensorMap ta, tb;
  cudaError_t e = make_map(&ta, p.a, p.ab_dtype, p.M, p.K, p.batch, kBM);
  if (e != cudaSuccess) return e;
  if (p.b_maps_dev == nullptr) {
    e = make_map(&tb, p.b, p.ab_dtype, p.N, p.K, p.batch, BN);
    if (e != cudaSuccess) return e;
  } else {
    std::memset(&tb, 0, sizeof(tb));
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

  dim3 grid((p.N + BN - 1) / BN, (p.M + kBM - 1) / kBM, p.batch * kp.split_k);
#define BFLC_LAUNCH(BN_, EPI_) return launch<BN_, EPI_>(ta, tb, kp, grid, stream)
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
