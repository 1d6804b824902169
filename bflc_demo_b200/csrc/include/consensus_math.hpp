// The committee-consensus decision procedure, written once and compiled for both the
// host C++ ledger and the device consensus kernel, so the HBM ledger replicas and the
// host chain can never disagree on an election.
//
// Spec = reference `CommitteePrecompiled::Aggregate`
// (FISCO-BCOS/libprecompiled/extension/CommitteePrecompiled.cpp:349-456):
//   0. per-trainer score = median over the committee's scores          (C:351-362)
//   1. rank trainers by score, descending                               (C:365-366)
//   2-3. aggregate the top AGGREGATE_COUNT, weighted by n_samples       (C:373-400)
//   5. next committee = top COMM_COUNT scorers; old committee -> trainer (C:444-455)
// Deliberate deviations:
//   * true median (mean of the two middle values for even counts); the reference's
//     quickselect `GetMid` (C:81-115) is input-order dependent and is NOT emulated;
//   * ties are broken by ascending rank id (the reference: unstable sort over
//     unordered_map order);
//   * optional score-weighted aggregation (`weight_by_score`), default off = reference.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define BFLC_HD __host__ __device__ __forceinline__
#else
#define BFLC_HD inline
#endif

namespace bflc {

constexpr int kCMaxRanks = 64;  // host ledger supports up to 64 clients; device path uses <= 8

enum RoleBits : uint32_t { ROLE_TRAINER = 1u, ROLE_COMM = 2u };

template <int MAXR>
struct ConsensusIn {
  int n_ranks;
  int n_comm;        // COMM_COUNT
  int n_aggregate;   // AGGREGATE_COUNT
  int weight_by_score;
  uint32_t role[MAXR];          // RoleBits
  uint8_t admitted[MAXR];       // 1 = this rank's update was admitted this round
  uint8_t scored[MAXR][MAXR];   // scored[c][t] = committee c supplied a score for trainer t
  float score[MAXR][MAXR];      // score[c][t]
  uint32_t n_samples[MAXR];
  float avg_cost[MAXR];
};

template <int MAXR>
struct ConsensusOut {
  float median[MAXR];
  int order[MAXR];      // admitted trainers sorted by (median desc, rank asc)
  int n_ranked;
  int n_selected;
  uint8_t selected[MAXR];
  float weight[MAXR];   // aggregation weights, sum to 1 over the selected trainers
  uint32_t role_after[MAXR];
  float global_loss;
};

BFLC_HD float median_of(float* v, int n) {
  // insertion sort: n <= committee size (tiny)
  for (int i = 1; i < n; ++i) {
    float x = v[i];
    int j = i - 1;
    while (j >= 0 && v[j] > x) { v[j + 1] = v[j]; --j; }
    v[j + 1] = x;
  }
  if (n <= 0) return 0.f;
  return (n & 1) ? v[n / 2] : 0.5f * (v[n / 2 - 1] + v[n / 2]);
}

template <int MAXR>
BFLC_HD void run_consensus(const ConsensusIn<MAXR>& in, ConsensusOut<MAXR>& out) {
  const int n = in.n_ranks;
  // 0. median committee score per admitted trainer
  out.n_ranked = 0;
  for (int t = 0; t < n; ++t) {
    out.median[t] = 0.f;
    out.selected[t] = 0;
    out.weight[t] = 0.f;
    if (!in.admitted[t]) continue;
    float tmp[MAXR];
    int m = 0;
    for (int c = 0; c < n; ++c)
      if ((in.role[c] & ROLE_COMM) && in.scored[c][t]) tmp[m++] = in.score[c][t];
    out.median[t] = median_of(tmp, m);
    out.order[out.n_ranked++] = t;
  }
  // 1. sort by (median desc, rank asc) -- insertion sort keeps it stable and tiny
  for (int i = 1; i < out.n_ranked; ++i) {
    const int x = out.order[i];
    int j = i - 1;
    while (j >= 0 && (out.median[out.order[j]] < out.median[x])) {
      out.order[j + 1] = out.order[j];
      --j;
    }
    out.order[j + 1] = x;
  }
  // 2-3. top-K, weights
  const int k = in.n_aggregate < out.n_ranked ? in.n_aggregate : out.n_ranked;
  out.n_selected = k;
  double wsum = 0.0;
  float cost = 0.f;
  for (int i = 0; i < k; ++i) {
    const int t = out.order[i];
    out.selected[t] = 1;
    double w = static_cast<double>(in.n_samples[t]);
    if (in.weight_by_score) w *= static_cast<double>(out.median[t]);
    out.weight[t] = static_cast<float>(w);
    wsum += w;
    cost += in.avg_cost[t];
  }
  if (k > 0 && wsum <= 0.0) {  // degenerate: all-zero scores -> fall back to uniform
    for (int i = 0; i < k; ++i) out.weight[out.order[i]] = 1.f;
    wsum = static_cast<double>(k);
  }
  for (int i = 0; i < k; ++i) {
    const int t = out.order[i];
    out.weight[t] = static_cast<float>(static_cast<double>(out.weight[t]) / wsum);
  }
  out.global_loss = k > 0 ? cost / static_cast<float>(k) : 0.f;
  // 5. re-election. Solo mode (a rank that is both trainer and committee) keeps its roles.
  bool solo = false;
  for (int r = 0; r < n; ++r)
    if ((in.role[r] & ROLE_TRAINER) && (in.role[r] & ROLE_COMM)) solo = true;
  for (int r = 0; r < n; ++r) out.role_after[r] = solo ? in.role[r] : ROLE_TRAINER;
  if (!solo) {
    int elected = 0;
    for (int i = 0; i < out.n_ranked && elected < in.n_comm; ++i) {
      out.role_after[out.order[i]] = ROLE_COMM;
      ++elected;
    }
    // not enough scored trainers to fill the committee: keep the lowest-ranked old members
    for (int r = 0; r < n && elected < in.n_comm; ++r) {
      if ((in.role[r] & ROLE_COMM) && out.role_after[r] != ROLE_COMM) {
        out.role_after[r] = ROLE_COMM;
        ++elected;
      }
    }
  }
}

}  // namespace bflc
