// Stand-alone exerciser of the C++ ledger runtime for sanitizer builds:
//   g++ -std=c++17 -O1 -g -fsanitize=address,undefined ledger.cpp ledger_selftest.cpp -lpthread
//   g++ -std=c++17 -O1 -g -fsanitize=thread            ledger.cpp ledger_selftest.cpp -lpthread
// Drives full protocol rounds (register -> upload with quota -> scores -> aggregate -> re-election)
// from several client threads at once while reader threads poll the views, then checks the chain,
// snapshot/restore and replica determinism.  Exit code 0 = pass.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <thread>
#include <vector>

#include "ledger.hpp"

using namespace bflc;

namespace {

int fails = 0;
#define CHECK(cond)                                                     \
  do {                                                                  \
    if (!(cond)) {                                                      \
      std::fprintf(stderr, "CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #cond); \
      ++fails;                                                          \
    }                                                                   \
  } while (0)

constexpr uint32_t kTrainer = 1u, kComm = 2u;

// one client's role loop for `rounds` rounds (M:103-169, 196-228 in the reference client)
void client_loop(Ledger* lg, int id, int rounds, std::atomic<int>* uploads_rejected) {
  const LedgerConfig& cfg = lg->config();
  CHECK(lg->RegisterNode(id) == Status::OK);
  int trained = -1, scored = -1;
  while (true) {
    auto [role, ep] = lg->QueryState(id);
    if (ep == kEpochNotStarted) { std::this_thread::yield(); continue; }
    if (ep >= rounds) break;
    if ((role & kTrainer) && !(role & kComm) && trained < ep) {
      auto [w, mep] = lg->QueryGlobalModel();
      if (mep != ep) continue;
      std::vector<float> delta(w.size());
      for (size_t i = 0; i < w.size(); ++i) delta[i] = 0.01f * static_cast<float>((id * 7 + i + ep) % 11) - 0.05f;
      UpdateMeta meta{static_cast<uint32_t>(100 + id), 1.0f / static_cast<float>(1 + id + ep)};
      Status s = lg->UploadLocalUpdate(id, delta, meta, ep);
      if (s == Status::QUOTA_FULL) uploads_rejected->fetch_add(1);
      CHECK(s == Status::OK || s == Status::QUOTA_FULL || s == Status::STALE_EPOCH);
      trained = ep;
    } else if ((role & kComm) && scored < ep) {
      auto ups = lg->QueryAllUpdates();
      if (ups.empty()) { std::this_thread::yield(); continue; }
      if (static_cast<int>(ups.size()) != cfg.needed_update_count) continue;
      std::map<int, float> scores;
      for (const auto& u : ups) scores[u.sender] = 0.5f + 0.01f * static_cast<float>((u.sender * 13 + id + ep) % 17);
      Status s = lg->UploadScores(id, ep, scores);
      CHECK(s == Status::OK || s == Status::AGGREGATED || s == Status::STALE_EPOCH || s == Status::NOT_READY);
      if (s != Status::NOT_READY) scored = ep;
    } else {
      std::this_thread::yield();
    }
  }
}

}  // namespace

int main() {
  LedgerConfig cfg;  // the reference configuration: 20 clients, committee 4, top-6 of 10
  CHECK(cfg.validate().empty());
  const int rounds = 4;
  Ledger lg(cfg);
  std::atomic<int> rejected{0};
  std::atomic<bool> stop{false};
  std::vector<std::thread> readers;
  for (int r = 0; r < 3; ++r)
    readers.emplace_back([&] {
      while (!stop.load()) {
        (void)lg.QueryGlobalModel();
        (void)lg.counters();
        (void)lg.roles();
        (void)lg.state_hash();
        std::this_thread::yield();
      }
    });
  std::vector<std::thread> clients;
  for (int id = 0; id < cfg.client_num; ++id) clients.emplace_back(client_loop, &lg, id, rounds, &rejected);
  for (auto& t : clients) t.join();
  stop.store(true);
  for (auto& t : readers) t.join();

  CHECK(lg.epoch() == rounds);
  CHECK(lg.verify_chain());
  CHECK(lg.n_blocks() >= static_cast<size_t>(rounds));
  // 6 of the 16 trainers lose the first-10-wins race each round; a very late one sees STALE_EPOCH
  // instead of QUOTA_FULL, so the count is an upper bound under scheduling noise
  CHECK(rejected.load() <= rounds * (cfg.client_num - cfg.comm_count - cfg.needed_update_count));
  auto roles = lg.roles();
  int n_comm = 0;
  for (uint32_t r : roles) n_comm += (r & kComm) ? 1 : 0;
  CHECK(n_comm == cfg.comm_count);

  // snapshot / restore reproduces the state hash and keeps working
  const std::string blob = lg.snapshot();
  auto copy = Ledger::restore(blob);
  CHECK(copy != nullptr);
  if (copy) {
    CHECK(copy->state_hash() == lg.state_hash());
    CHECK(copy->verify_chain());
    CHECK(copy->epoch() == lg.epoch());
  }
  (void)lg.drain_log();
  if (fails == 0) std::puts("ledger_selftest OK");
  return fails == 0 ? 0 : 1;
}
