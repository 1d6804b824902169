// One persistent kernel for a trainer's whole local-training pass of the 2-layer MLP:
// every mini-batch step (forward, softmax-xent, both weight gradients, the hidden gradient
// and the optimizer) runs inside ONE launch; phases are separated by a device-wide barrier
// instead of kernel boundaries.
//
//   per step:  P1  h  = relu(x W1^T + b1)                       16 tiles   (K = 784)
//              P2  dlogits = softmax(h W2^T + b2) - onehot       4 tiles   (+loss, #correct, db2)
//              P3  dW2 = dlogits^T h   ||   dh = (dlogits W2) * relu'(h) (+db1)   4 + 16 tiles
//              P4  dW1 = dh^T x                                  26 tiles
//              P5  SGD / Adam over the flat buffer (+bf16 shadow refresh, grad zeroing)
//
// Each GEMM tile is the same tcgen05 / TMEM / TMA pipeline as gemm_sm100.cu (128 x 64 tiles,
// 8-stage 128B-swizzled ring, one elected MMA thread, staged coalesced epilogue); the smem
// ring, its mbarriers and the TMEM allocation persist across tiles, phases and steps.
//
// Why: at this problem size every stand-alone GEMM launch costs 6-12 us of which only a
// fraction is math (launch, prologue, first-TMA latency, drain) -- six launches per step,
// 48 per round.  Inside one kernel the fixed costs are paid once and a phase boundary is a
// ~1 us grid barrier.  (Reference step: python-sdk/main.py:141-148, three sess.run calls.)
#include <cuda_bf16.h>

#include <algorithm>
#include <cstring>

#include "bflc_kernels.h"
#include "launch.cuh"
#include "sm100_ptx.cuh"

namespace bflc {

namespace {

constexpr int kBM = 128, kBN = 64, kStages = 8;
constexpr int kABytes = kBM * 128, kBBytes = kBN * 128, kStageBytes = kABytes + kBBytes;
constexpr int kTileBytes = kStages * kStageBytes;
constexpr int kBarBytes = 256;
constexpr int kStgLd = 36;
constexpr int kStgBytes = 4 * 32 * kStgLd * 4;
constexpr int kSmemTotal = kTileBytes + kBarBytes + kStgBytes + kBN * 4 + 1024;
constexpr int kThreads = 192;
constexpr int kGrid = 32;

enum EpiMode : int { E_BIAS_RELU_BF16 = 0, E_XENT = 1, E_F32 = 2, E_MASK_COLSUM_BF16 = 3 };

struct Maps {  // 10 TMA descriptors, all bf16, SWIZZLE_128B
  CUtensorMap x_k, w1_k, h_k, w2_k, dl_mn, h_mn, dl_k, w2_mn, dh_mn, x_mn;
};

struct Args {
  int B, steps, in_dim, hidden, n_classes, ncp;  // ncp = dlogits row stride (padded classes)
  long long n_params;
  const int* pred;               // whole kernel is a no-op when *pred == 0 (non-trainer rank)
  unsigned int* barrier;         // device-wide phase barrier counter (zeroed before launch)
  // parameters / optimizer state
  float* master; const float* b1; const float* b2;
  float* grad; float* gw1; float* gb1; float* gw2; float* gb2;
  __nv_bfloat16* shadow;
  float* adam_m; float* adam_v;
  int adam; float lr, beta1, beta2, eps; const int* step_base;
  // activations
  __nv_bfloat16* h; __nv_bfloat16* dlogits; __nv_bfloat16* dh;
  const int32_t* labels;
  float* loss_sum; unsigned int* correct;
};

struct Job {  // one 128 x 64 output tile
  const CUtensorMap* ta; const CUtensorMap* tb;
  int a_mn, b_mn;
  int a_c0, a_c1, b_c0, b_c1;   // TMA coordinates of K-block 0 (c0 = innermost)
  int n_kb;
  int m0, n0, M, N;             // output tile origin / logical extent
  int mode;
  long long ldd;
  void* d;                      // output
  const float* bias;            // E_BIAS_RELU_BF16 / E_XENT
  const __nv_bfloat16* aux;     // E_MASK_COLSUM_BF16: relu mask source, same shape as d
  float* colsum;
  const int32_t* labels;        // E_XENT (already offset to this step's rows)
  float grad_scale;
};

struct Pipe {  // persistent pipeline state of one role
  uint32_t it;    // K-blocks processed so far (ring slot / parity)
  uint32_t tile;  // tiles processed so far (accumulator barrier parity)
};

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ void stage_put(float* stg, int lane, const float (&v)[32]) {
  float4* rowp = reinterpret_cast<float4*>(stg + lane * kStgLd);
#pragma unroll
  for (int j = 0; j < 8; ++j) rowp[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}
__device__ __forceinline__ void stage_get(const float* stg, int lane, float (&v)[32]) {
  const float4* rowp = reinterpret_cast<const float4*>(stg + lane * kStgLd);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 t = rowp[j];
    v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
  }
}

// ---------------------------------------------------------------- producer / MMA / epilogue
__device__ __forceinline__ void produce_tile(const Job& j, uint8_t* smem, uint64_t* full_bar,
                                             uint64_t* empty_bar, Pipe& pp) {
  for (int i = 0; i < j.n_kb; ++i, ++pp.it) {
    const int s = pp.it % kStages;
    const uint32_t ph = (pp.it / kStages) & 1;
    ptx::mbar_wait(&empty_bar[s], ph ^ 1);
    uint8_t* sa = smem + s * kStageBytes;
    uint8_t* sb = sa + kABytes;
    if (ptx::elect_one()) {
      ptx::mbar_expect_tx(&full_bar[s], kStageBytes);
      if (!j.a_mn) {
        ptx::tma_load_3d(sa, j.ta, &full_bar[s], j.a_c0 + i * 64, j.a_c1, 0);
      } else {
        ptx::tma_load_3d(sa, j.ta, &full_bar[s], j.a_c0, j.a_c1 + i * 64, 0);
        ptx::tma_load_3d(sa + 64 * 128, j.ta, &full_bar[s], j.a_c0 + 64, j.a_c1 + i * 64, 0);
      }
      if (!j.b_mn)
        ptx::tma_load_3d(sb, j.tb, &full_bar[s], j.b_c0 + i * 64, j.b_c1, 0);
      else
        ptx::tma_load_3d(sb, j.tb, &full_bar[s], j.b_c0, j.b_c1 + i * 64, 0);
    }
    __syncwarp();
  }
}

__device__ __forceinline__ void mma_tile(const Job& j, uint8_t* smem, uint64_t* full_bar,
                                         uint64_t* empty_bar, uint64_t* accum_bar,
                                         uint32_t tmem_base, Pipe& pp) {
  const uint32_t idesc = ptx::make_idesc(1u, j.a_mn ? 1u : 0u, j.b_mn ? 1u : 0u, kBM, kBN);
  const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO = 1024, v1, SWIZZLE_128B
  const uint32_t base_lo = ptx::smem_u32(smem) >> 4;
  const uint32_t lbo_a = j.a_mn ? (8192u >> 4) : 1u, lbo_b = j.b_mn ? (8192u >> 4) : 1u;
  const uint32_t lo_a0 = base_lo | (lbo_a << 16);
  const uint32_t lo_b0 = (base_lo + (kABytes >> 4)) | (lbo_b << 16);
  const uint32_t ks_a = (j.a_mn ? 2048u : 32u) >> 4, ks_b = (j.b_mn ? 2048u : 32u) >> 4;
  for (int i = 0; i < j.n_kb; ++i, ++pp.it) {
    const int s = pp.it % kStages;
    const uint32_t ph = (pp.it / kStages) & 1;
    ptx::mbar_wait(&full_bar[s], ph);
    ptx::tc_fence_after_sync();
    const uint32_t so = static_cast<uint32_t>(s) * (kStageBytes >> 4);
    if (ptx::elect_one()) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t ad = (static_cast<uint64_t>(hi) << 32) | (lo_a0 + so + k * ks_a);
        const uint64_t bd = (static_cast<uint64_t>(hi) << 32) | (lo_b0 + so + k * ks_b);
        ptx::umma_f16(tmem_base, ad, bd, idesc, (i > 0 || k > 0) ? 1u : 0u);
      }
      ptx::umma_commit(&empty_bar[s]);
    }
    __syncwarp();
  }
  if (ptx::elect_one()) ptx::umma_commit(accum_bar);
  __syncwarp();
  ++pp.tile;
}

// epilogue warps 2..5; `warp` is the hardware warp index
__device__ __forceinline__ void epilogue_tile(const Job& j, const Args& a, int warp, int lane,
                                              uint64_t* accum_bar, uint32_t tmem_base,
                                              float* stage_base, float* sbias, Pipe& pp) {
  const int q = warp & 3;
  float* stg = stage_base + (warp - 2) * (32 * kStgLd);
  const int row_base = j.m0 + q * 32;
  const int row = row_base + lane;
  const bool row_ok = row < j.M;
  const int cr = lane >> 3, cg = (lane & 7) * 4;
  {
    const int et = threadIdx.x - 64;
    // coherent (L2) loads: the biases are rewritten by the optimizer phase of this same kernel
    for (int i = et; i < kBN; i += 128)
      sbias[i] = (j.bias != nullptr && j.n0 + i < j.N) ? __ldcg(j.bias + j.n0 + i) : 0.f;
    asm volatile("bar.sync 1, 128;" ::: "memory");
  }
  ptx::mbar_wait(accum_bar, pp.tile & 1);
  ptx::tc_fence_after_sync();
  ++pp.tile;
  const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);

  if (j.mode != E_XENT) {
#pragma unroll 1
    for (int c = 0; c < kBN / 32; ++c) {
      const int nc = j.n0 + c * 32;
      if (nc >= j.N) break;
      uint32_t r[32];
      ptx::tmem_ld_32x32b_x32(taddr + c * 32, r);
      ptx::tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) v[k] = __uint_as_float(r[k]) + sbias[c * 32 + k];
      if (j.mode == E_BIAS_RELU_BF16) {
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = fmaxf(v[k], 0.f);
      } else if (j.mode == E_MASK_COLSUM_BF16) {
        // coalesced (L2-coherent) load of the mask tile through the staging buffer
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int rr = it * 4 + cr, rw = row_base + rr, col = nc + cg;
          float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
          if (rw < j.M && col + 3 < j.N) {
            const uint2 u = __ldcg(reinterpret_cast<const uint2*>(j.aux + static_cast<long long>(rw) * j.ldd + col));
            const float2 lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
            const float2 hi2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
            x = make_float4(lo.x, lo.y, hi2.x, hi2.y);
          }
          *reinterpret_cast<float4*>(stg + rr * kStgLd + cg) = x;
        }
        __syncwarp();
        float m[32];
        stage_get(stg, lane, m);
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = m[k] > 0.f ? v[k] : 0.f;
      }
      stage_put(stg, lane, v);
      __syncwarp();
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int rr = it * 4 + cr, rw = row_base + rr, col = nc + cg;
        if (rw >= j.M || col >= j.N) continue;
        const float4 x = *reinterpret_cast<const float4*>(stg + rr * kStgLd + cg);
        const long long off = static_cast<long long>(rw) * j.ldd + col;
        if (j.mode == E_F32) {
          float* d = reinterpret_cast<float*>(j.d) + off;
          if (col + 3 < j.N) *reinterpret_cast<float4*>(d) = x;
          else {
            const float xs[4] = {x.x, x.y, x.z, x.w};
            for (int k = 0; k < 4; ++k) if (col + k < j.N) d[k] = xs[k];
          }
        } else {
          __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(j.d) + off;
          *reinterpret_cast<uint2*>(d) = make_uint2(pack2(x.x, x.y), pack2(x.z, x.w));
        }
      }
      if (j.colsum != nullptr) {
        float tot = 0.f;
        const int rmax = min(32, j.M - row_base);
        for (int rr = 0; rr < rmax; ++rr) tot += stg[rr * kStgLd + lane];
        if (nc + lane < j.N) atomicAdd(j.colsum + nc + lane, tot);
      }
      __syncwarp();
    }
  } else {
    // softmax cross-entropy over the N (<= 64) logits of each row
    const int32_t label = row_ok ? j.labels[row] : -1;
    float vmax = -INFINITY, zlab = 0.f;
    int amax = -1;
    float z[64];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t r[32];
      ptx::tmem_ld_32x32b_x32(taddr + c * 32, r);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const int n = c * 32 + k;
        const float x = __uint_as_float(r[k]) + sbias[n];
        z[n] = x;
        if (n < j.N) {
          if (x > vmax) { vmax = x; amax = n; }
          if (n == label) zlab = x;
        }
      }
    }
    float sum = 0.f;
#pragma unroll
    for (int n = 0; n < 64; ++n)
      if (n < j.N) sum += __expf(z[n] - vmax);
    const float inv = 1.f / sum;
    float loss = row_ok ? (__logf(sum) + vmax - zlab) : 0.f;
    const bool hit = row_ok && (amax == label);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      float v[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const int n = c * 32 + k;
        v[k] = (n < j.N && row_ok)
                   ? (__expf(z[n] - vmax) * inv - (n == label ? 1.f : 0.f)) * j.grad_scale
                   : 0.f;
      }
      stage_put(stg, lane, v);
      __syncwarp();
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int rr = it * 4 + cr, rw = row_base + rr, col = c * 32 + cg;
        if (rw >= j.M || col >= j.ldd) continue;
        const float4 x = *reinterpret_cast<const float4*>(stg + rr * kStgLd + cg);
        __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(j.d) + static_cast<long long>(rw) * j.ldd + col;
        *reinterpret_cast<uint2*>(d) = make_uint2(pack2(x.x, x.y), pack2(x.z, x.w));
      }
      if (j.colsum != nullptr) {
        float tot = 0.f;
        for (int rr = 0; rr < 32; ++rr) tot += stg[rr * kStgLd + lane];
        if (c * 32 + lane < j.N) atomicAdd(j.colsum + c * 32 + lane, tot);
      }
      __syncwarp();
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, off);
    const unsigned cnt = __popc(__ballot_sync(0xffffffffu, hit));
    if (lane == 0) {
      atomicAdd(a.loss_sum, loss);
      if (cnt) atomicAdd(a.correct, cnt);
    }
  }
  ptx::tc_fence_before_sync();
}

// Device-wide barrier between phases.  All kGrid CTAs are co-resident (one per SM), the counter
// only grows.  Writers: bar.sync orders every thread's writes before thread 0's gpu-scope
// fence (cumulative release); readers: acquire, then a proxy fence so the next phase's TMA
// (async proxy) observes what other CTAs stored with ordinary instructions.
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int& epoch) {
  ++epoch;
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x == 0) {
    ptx::fence_proxy_async_all();
    __threadfence();
    atomicAdd(counter, 1u);
    const unsigned int target = epoch * gridDim.x;
    unsigned long long spins = 0;
    while (true) {
      unsigned int v;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
      if (v >= target) break;
      if (++spins > (1ull << 27)) __trap();  // a lost CTA traps within seconds instead of hanging
    }
    ptx::fence_proxy_async_all();
  }
  __syncthreads();
  ptx::tc_fence_after_sync();
}

__global__ void __launch_bounds__(kThreads, 1)
mlp_round_kernel(const __grid_constant__ Maps maps, const Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kTileBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* accum_bar = empty_bar + kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);
  float* stage_base = reinterpret_cast<float*>(smem + kTileBytes + kBarBytes);
  float* sbias = stage_base + 4 * 32 * kStgLd;

  ptx::pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    ptx::mbar_init(accum_bar, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, kBN);
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  ptx::pdl_wait();
  if (a.pred != nullptr && *a.pred == 0) {
    if (warp == 1) ptx::tmem_dealloc(tmem_base, kBN);
    return;
  }

  Pipe pp{0u, 0u};
  unsigned int bar_epoch = 0;
  const int t = blockIdx.x;
  const int B = a.B, H = a.hidden, C = a.n_classes, D = a.in_dim;
  const int mt_b = (B + kBM - 1) / kBM;                 // M-tiles over the batch
  const int nt_h = (H + kBN - 1) / kBN;                 // N-tiles over hidden
  const int nt_d = (D + kBN - 1) / kBN;                 // N-tiles over in_dim
  const int mt_h = (H + kBM - 1) / kBM;                 // M-tiles over hidden (dW1)
  const int kb_d = (D + 63) / 64, kb_h = (H + 63) / 64, kb_b = (B + 63) / 64, kb_c = (C + 63) / 64;

  auto run = [&](const Job& j) {
    if (warp == 0) produce_tile(j, smem, full_bar, empty_bar, pp);
    else if (warp == 1) mma_tile(j, smem, full_bar, empty_bar, accum_bar, tmem_base, pp);
    else epilogue_tile(j, a, warp, lane, accum_bar, tmem_base, stage_base, sbias, pp);
  };

  for (int step = 0; step < a.steps; ++step) {
    const int r0 = step * B;
    // ---- P1: h = relu(x W1^T + b1)
    if (t < mt_b * nt_h) {
      Job j{};
      j.ta = &maps.x_k; j.tb = &maps.w1_k; j.a_mn = 0; j.b_mn = 0;
      j.m0 = (t / nt_h) * kBM; j.n0 = (t % nt_h) * kBN; j.M = B; j.N = H;
      j.a_c0 = 0; j.a_c1 = r0 + j.m0; j.b_c0 = 0; j.b_c1 = j.n0; j.n_kb = kb_d;
      j.mode = E_BIAS_RELU_BF16; j.d = a.h; j.ldd = H; j.bias = a.b1;
      run(j);
    }
    grid_barrier(a.barrier, bar_epoch);
    // ---- P2: logits -> dlogits / loss / db2
    if (t < mt_b) {
      Job j{};
      j.ta = &maps.h_k; j.tb = &maps.w2_k; j.a_mn = 0; j.b_mn = 0;
      j.m0 = t * kBM; j.n0 = 0; j.M = B; j.N = C;
      j.a_c0 = 0; j.a_c1 = j.m0; j.b_c0 = 0; j.b_c1 = 0; j.n_kb = kb_h;
      j.mode = E_XENT; j.d = a.dlogits; j.ldd = a.ncp; j.bias = a.b2; j.colsum = a.gb2;
      j.labels = a.labels + r0; j.grad_scale = 1.f / static_cast<float>(B);
      run(j);
    }
    grid_barrier(a.barrier, bar_epoch);
    // ---- P3: dW2 (tiles [0, nt_h))  ||  dh (tiles [nt_h, nt_h + mt_b*nt_h))
    if (t < nt_h) {
      Job j{};
      j.ta = &maps.dl_mn; j.tb = &maps.h_mn; j.a_mn = 1; j.b_mn = 1;
      j.m0 = 0; j.n0 = t * kBN; j.M = C; j.N = H;
      j.a_c0 = 0; j.a_c1 = 0; j.b_c0 = j.n0; j.b_c1 = 0; j.n_kb = kb_b;
      j.mode = E_F32; j.d = a.gw2; j.ldd = H;
      run(j);
    } else if (t < nt_h + mt_b * nt_h) {
      const int u = t - nt_h;
      Job j{};
      j.ta = &maps.dl_k; j.tb = &maps.w2_mn; j.a_mn = 0; j.b_mn = 1;
      j.m0 = (u / nt_h) * kBM; j.n0 = (u % nt_h) * kBN; j.M = B; j.N = H;
      j.a_c0 = 0; j.a_c1 = j.m0; j.b_c0 = j.n0; j.b_c1 = 0; j.n_kb = kb_c;
      j.mode = E_MASK_COLSUM_BF16; j.d = a.dh; j.ldd = H; j.aux = a.h; j.colsum = a.gb1;
      run(j);
    }
    grid_barrier(a.barrier, bar_epoch);
    // ---- P4: dW1 = dh^T x
    if (t < mt_h * nt_d) {
      Job j{};
      j.ta = &maps.dh_mn; j.tb = &maps.x_mn; j.a_mn = 1; j.b_mn = 1;
      j.m0 = (t / nt_d) * kBM; j.n0 = (t % nt_d) * kBN; j.M = H; j.N = D;
      j.a_c0 = j.m0; j.a_c1 = 0; j.b_c0 = j.n0; j.b_c1 = r0; j.n_kb = kb_b;
      j.mode = E_F32; j.d = a.gw1; j.ldd = D;
      run(j);
    }
    grid_barrier(a.barrier, bar_epoch);
    // ---- P5: optimizer over the flat buffer (all threads of all CTAs)
    {
      const long long nv = a.n_params / 4;
      const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
      float bc1 = 1.f, bc2 = 1.f;
      if (a.adam) {
        const int tt = (a.step_base ? *a.step_base : 0) + step + 1;
        bc1 = 1.f - powf(a.beta1, static_cast<float>(tt));
        bc2 = 1.f - powf(a.beta2, static_cast<float>(tt));
      }
      for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nv;
           i += stride) {
        const float4 w4 = __ldcg(reinterpret_cast<const float4*>(a.master) + i);
        const float4 g4 = __ldcg(reinterpret_cast<const float4*>(a.grad) + i);
        float w[4] = {w4.x, w4.y, w4.z, w4.w};
        const float g[4] = {g4.x, g4.y, g4.z, g4.w};
        if (a.adam) {
          const float4 m4 = __ldcg(reinterpret_cast<const float4*>(a.adam_m) + i);
          const float4 v4 = __ldcg(reinterpret_cast<const float4*>(a.adam_v) + i);
          float m[4] = {m4.x, m4.y, m4.z, m4.w}, v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            m[k] = a.beta1 * m[k] + (1.f - a.beta1) * g[k];
            v[k] = a.beta2 * v[k] + (1.f - a.beta2) * g[k] * g[k];
            w[k] -= a.lr * (m[k] / bc1) / (sqrtf(v[k] / bc2) + a.eps);
          }
          reinterpret_cast<float4*>(a.adam_m)[i] = make_float4(m[0], m[1], m[2], m[3]);
          reinterpret_cast<float4*>(a.adam_v)[i] = make_float4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k) w[k] -= a.lr * g[k];
        }
        reinterpret_cast<float4*>(a.master)[i] = make_float4(w[0], w[1], w[2], w[3]);
        reinterpret_cast<uint2*>(a.shadow)[i] = make_uint2(pack2(w[0], w[1]), pack2(w[2], w[3]));
        reinterpret_cast<float4*>(a.grad)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    grid_barrier(a.barrier, bar_epoch);
  }

  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, kBN);
  }
}

}  // namespace

cudaError_t mlp_round_sm100(const MlpRoundArgs& r, cudaStream_t stream) {
  bind_context_once();
  if (r.hidden % 8 || r.in_dim % 8 || r.n_params % 4 || r.batch % 8 || r.ncp % 8 || r.n_classes > 64)
    return cudaErrorInvalidValue;
  const int mt_b = (r.batch + kBM - 1) / kBM, nt_h = (r.hidden + kBN - 1) / kBN;
  const int nt_d = (r.in_dim + kBN - 1) / kBN, mt_h = (r.hidden + kBM - 1) / kBM;
  const int need = std::max(std::max(mt_b * nt_h + nt_h, mt_h * nt_d), mt_b);
  if (need > kGrid * 4) return cudaErrorInvalidValue;
  const int grid = need > kGrid ? need : kGrid;
  if (grid > 148) return cudaErrorInvalidValue;

  Maps m;
  const long long rows_x = static_cast<long long>(r.steps) * r.batch;
  auto mk = [&](CUtensorMap* out, const void* ptr, long long ld, bool mn, int rows_extent, int K,
                int rows_tile) {
    GemmOperand op{ptr, ld, 0, mn};
    return gemm_make_operand_map(out, op, DType::BF16, rows_extent, K, 1, rows_tile);
  };
  cudaError_t e;
  // K-major: (rows_extent = M|N, K);  MN-major: memory [K][M|N]
  if ((e = mk(&m.x_k, r.x, r.in_dim, false, (int)rows_x, r.in_dim, kBM)) != cudaSuccess) return e;
  if ((e = mk(&m.w1_k, r.w1_shadow, r.in_dim, false, r.hidden, r.in_dim, kBN)) != cudaSuccess) return e;
  if ((e = mk(&m.h_k, r.h, r.hidden, false, r.batch, r.hidden, kBM)) != cudaSuccess) return e;
  if ((e = mk(&m.w2_k, r.w2_shadow, r.hidden, false, r.n_classes, r.hidden, kBN)) != cudaSuccess) return e;
  if ((e = mk(&m.dl_mn, r.dlogits, r.ncp, true, r.n_classes, r.batch, kBM)) != cudaSuccess) return e;
  if ((e = mk(&m.h_mn, r.h, r.hidden, true, r.hidden, r.batch, kBN)) != cudaSuccess) return e;
  if ((e = mk(&m.dl_k, r.dlogits, r.ncp, false, r.batch, r.n_classes, kBM)) != cudaSuccess) return e;
  if ((e = mk(&m.w2_mn, r.w2_shadow, r.hidden, true, r.hidden, r.n_classes, kBN)) != cudaSuccess) return e;
  if ((e = mk(&m.dh_mn, r.dh, r.hidden, true, r.hidden, r.batch, kBM)) != cudaSuccess) return e;
  if ((e = mk(&m.x_mn, r.x, r.in_dim, true, r.in_dim, (int)rows_x, kBN)) != cudaSuccess) return e;

  Args a{};
  a.B = r.batch; a.steps = r.steps; a.in_dim = r.in_dim; a.hidden = r.hidden;
  a.n_classes = r.n_classes; a.ncp = r.ncp; a.n_params = r.n_params;
  a.pred = r.pred ? r.pred : current_predicate();
  a.barrier = r.barrier;
  a.master = r.master; a.b1 = r.b1; a.b2 = r.b2;
  a.grad = r.grad; a.gw1 = r.gw1; a.gb1 = r.gb1; a.gw2 = r.gw2; a.gb2 = r.gb2;
  a.shadow = reinterpret_cast<__nv_bfloat16*>(r.shadow);
  a.adam_m = r.adam_m; a.adam_v = r.adam_v; a.adam = r.adam ? 1 : 0;
  a.lr = r.lr; a.beta1 = r.beta1; a.beta2 = r.beta2; a.eps = r.eps; a.step_base = r.step_base;
  a.h = reinterpret_cast<__nv_bfloat16*>(r.h);
  a.dlogits = reinterpret_cast<__nv_bfloat16*>(r.dlogits);
  a.dh = reinterpret_cast<__nv_bfloat16*>(r.dh);
  a.labels = r.labels; a.loss_sum = r.loss_sum; a.correct = r.correct;

  static bool configured = false;
  if (!configured) {
    e = cudaFuncSetAttribute(mlp_round_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  note_launch();
  return launch_pdl(mlp_round_kernel, dim3(grid), dim3(kThreads), kSmemTotal, stream, m, a);
}

}  // namespace bflc
