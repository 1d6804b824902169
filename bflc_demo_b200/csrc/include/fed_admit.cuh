// Device side of first-K-wins admission (struct AdmitPage, bflc_kernels.h): the ticket counter,
// the slot publication by an admitted trainer and the slot wait of its consumers.  Shared by the
// stand-alone upload kernel (fed_kernels.cu) and the persistent trainer whose last optimizer
// epilogue is the upload (mlp_round_sm100.cu).
#pragma once
#include "bflc_kernels.h"
#include "sm100_ptx.cuh"

namespace bflc {
namespace admit {

__device__ __forceinline__ AdmitPage* page(char* heap_base, const HeapLayout& lay, uint32_t parity) {
  return reinterpret_cast<AdmitPage*>(heap_base + lay.admit_off) + parity;
}
__device__ __forceinline__ bool first_k(const RoundState* st) {
  uint32_t trainers = 0;
  for (uint32_t r = 0; r < st->n_ranks; ++r) trainers += (st->role[r] & 1u) ? 1u : 0u;
  return st->n_needed != 0u && st->n_needed < trainers;
}
// Ticket of this trainer for round `epoch` (0-based), or -1 when the counter already belongs to a
// later round (a straggler more than a round behind).  The counter only ever grows.
__device__ __forceinline__ int take_ticket(uint32_t* counter, uint32_t epoch) {
  const uint32_t tag = epoch + 1u;
  uint32_t cur = ptx::ld_relaxed_sys(counter);
  while (true) {
    uint32_t want;
    if ((cur >> 8) < tag) want = (tag << 8) | 1u;
    else if ((cur >> 8) == tag) want = cur + 1u;
    else return -1;
    uint32_t old;
    asm volatile("atom.cas.acq_rel.sys.global.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(counter), "r"(cur), "r"(want)
                 : "memory");
    if (old == cur) return static_cast<int>(want & 0xffu) - 1;
    cur = old;
  }
}
// Wait until candidate slot z of this rank's replica holds an admission of round `epoch`;
// returns the admitted trainer's rank.  (acquire: the trainer's upload is readable afterwards)
__device__ __forceinline__ int wait_slot(const AdmitPage* pg, int z, uint32_t epoch) {
  const uint32_t tag = epoch + 1u;
  unsigned long long spins = 0;
  while (true) {
    const uint32_t v = ptx::ld_acquire_sys(&pg->slot[z]);
    if ((v >> 8) == tag) return static_cast<int>(v & 0xffu);
    if (++spins > BFLC_SPIN_LIMIT) __trap();
    __nanosleep(20);
  }
}
__device__ __forceinline__ void straggle(int us) {
  if (us <= 0) return;
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  do {
    __nanosleep(1000);
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  } while (t - t0 < static_cast<unsigned long long>(us) * 1000ull);
}

}  // namespace admit
}  // namespace bflc
