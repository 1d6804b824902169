// pybind11 module `_ledger`: the C++ ledger runtime without any CUDA / torch dependency,
// so the protocol tests and the gloo plumbing path run on a CPU-only box.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <cstring>

#include "consensus_math.hpp"
#include "ledger.hpp"

namespace py = pybind11;
using namespace bflc;

namespace {
py::bytes h2b(const Hash256& h) { return py::bytes(reinterpret_cast<const char*>(h.data()), 32); }

py::dict block_to_dict(const Block& b) {
  py::dict d;
  d["index"] = b.index; d["epoch"] = b.epoch;
  d["prev_hash"] = hex(b.prev_hash); d["hash"] = hex(b.hash);
  d["role_before"] = b.role_before; d["role_after"] = b.role_after;
  d["admitted"] = b.admitted; d["committee"] = b.committee; d["scores"] = b.scores;
  d["median"] = b.median; d["selected"] = b.selected; d["weight"] = b.weight;
  d["global_loss"] = b.global_loss; d["model_hash"] = hex(b.model_hash);
  d["device_digest"] = b.device_digest; d["from_device"] = (bool)b.from_device;
  return d;
}

std::vector<float> arr_to_vec(const py::array_t<float, py::array::c_style | py::array::forcecast>& a) {
  return std::vector<float>(a.data(), a.data() + a.size());
}
}  // namespace

PYBIND11_MODULE(_ledger, m) {
  m.doc() = "bflc_demo_b200 C++ ledger runtime";
  m.attr("EPOCH_NOT_STARTED") = kEpochNotStarted;
  m.attr("ROLE_TRAINER") = (uint32_t)ROLE_TRAINER;
  m.attr("ROLE_COMM") = (uint32_t)ROLE_COMM;

  py::enum_<Status>(m, "Status")
      .value("OK", Status::OK).value("NOT_STARTED", Status::NOT_STARTED)
      .value("STALE_EPOCH", Status::STALE_EPOCH).value("DUPLICATE", Status::DUPLICATE)
      .value("QUOTA_FULL", Status::QUOTA_FULL).value("NOT_COMMITTEE", Status::NOT_COMMITTEE)
      .value("UNKNOWN_CLIENT", Status::UNKNOWN_CLIENT).value("BAD_PAYLOAD", Status::BAD_PAYLOAD)
      .value("AGGREGATED", Status::AGGREGATED).value("NOT_TRAINER", Status::NOT_TRAINER)
      .value("NOT_READY", Status::NOT_READY);

  py::class_<LedgerConfig>(m, "LedgerConfig")
      .def(py::init<>())
      .def_readwrite("client_num", &LedgerConfig::client_num)
      .def_readwrite("comm_count", &LedgerConfig::comm_count)
      .def_readwrite("aggregate_count", &LedgerConfig::aggregate_count)
      .def_readwrite("needed_update_count", &LedgerConfig::needed_update_count)
      .def_readwrite("learning_rate", &LedgerConfig::learning_rate)
      .def_readwrite("model_size", &LedgerConfig::model_size)
      .def_readwrite("weight_by_score", &LedgerConfig::weight_by_score)
      .def_readwrite("solo", &LedgerConfig::solo)
      .def_readwrite("seed", &LedgerConfig::seed)
      .def("validate", &LedgerConfig::validate);

  py::class_<Ledger>(m, "Ledger")
      .def(py::init<const LedgerConfig&>())
      .def("RegisterNode", &Ledger::RegisterNode)
      .def("QueryState", &Ledger::QueryState)
      .def("QueryGlobalModel",
           [](Ledger& L) {
             auto r = L.QueryGlobalModel();
             py::array_t<float> a(r.first.size());
             std::memcpy(a.mutable_data(), r.first.data(), r.first.size() * sizeof(float));
             return py::make_tuple(a, r.second);
           })
      .def("UploadLocalUpdate",
           [](Ledger& L, int client,
              py::array_t<float, py::array::c_style | py::array::forcecast> delta,
              uint32_t n_samples, float avg_cost, int ep) {
             UpdateMeta meta;
             meta.n_samples = n_samples; meta.avg_cost = avg_cost;
             std::vector<float> d = arr_to_vec(delta);
             py::gil_scoped_release rel;
             return L.UploadLocalUpdate(client, d, meta, ep);
           })
      .def("UploadScores",
           [](Ledger& L, int client, int ep, const std::map<int, float>& scores) {
             return L.UploadScores(client, ep, scores);
           })
      .def("QueryAllUpdates",
           [](Ledger& L) {
             py::list out;
             for (auto& u : L.QueryAllUpdates()) {
               py::dict d;
               py::array_t<float> a(u.delta.size());
               std::memcpy(a.mutable_data(), u.delta.data(), u.delta.size() * sizeof(float));
               d["sender"] = u.sender; d["delta"] = a; d["n_samples"] = u.meta.n_samples;
               d["avg_cost"] = u.meta.avg_cost; d["arrival"] = u.arrival;
               out.append(d);
             }
             return out;
           })
      .def("Bootstrap", &Ledger::Bootstrap)
      .def("AppendDeviceRound",
           [](Ledger& L, const py::dict& d) {
             Ledger::DeviceRound r;
             r.epoch = d["epoch"].cast<int>();
             r.role_before = d["role_before"].cast<std::vector<uint32_t>>();
             r.role_after = d["role_after"].cast<std::vector<uint32_t>>();
             r.score_rows = d["score_rows"].cast<std::vector<std::vector<float>>>();
             r.scored_mask = d["scored_mask"].cast<std::vector<uint32_t>>();
             r.n_samples = d["n_samples"].cast<std::vector<uint32_t>>();
             r.avg_cost = d["avg_cost"].cast<std::vector<float>>();
             r.admitted_mask = d["admitted_mask"].cast<uint32_t>();
             r.selected_mask = d["selected_mask"].cast<uint32_t>();
             r.global_loss = d["global_loss"].cast<float>();
             r.model_digest = d["model_digest"].cast<uint64_t>();
             r.weight_by_score = d["weight_by_score"].cast<int>();
             return L.AppendDeviceRound(r);
           })
      .def("epoch", &Ledger::epoch)
      .def("roles", &Ledger::roles)
      .def("update_count", &Ledger::update_count)
      .def("score_count", &Ledger::score_count)
      .def("n_blocks", &Ledger::n_blocks)
      .def("last_global_loss", &Ledger::last_global_loss)
      .def("blocks",
           [](Ledger& L) {
             py::list out;
             for (auto& b : L.blocks()) out.append(block_to_dict(b));
             return out;
           })
      .def("counters",
           [](Ledger& L) {
             OpCounters c = L.counters();
             py::dict d;
             d["calls"] = c.calls; d["register_ok"] = c.register_ok; d["uploads_ok"] = c.uploads_ok;
             d["uploads_rejected"] = c.uploads_rejected; d["scores_ok"] = c.scores_ok;
             d["scores_rejected"] = c.scores_rejected; d["aggregations"] = c.aggregations;
             d["queries"] = c.queries;
             return d;
           })
      .def("drain_log", &Ledger::drain_log)
      .def("state_hash", [](Ledger& L) { return hex(L.state_hash()); })
      .def("verify_chain", &Ledger::verify_chain)
      .def("snapshot", [](Ledger& L) { return py::bytes(L.snapshot()); })
      .def_static("restore", [](const py::bytes& b) { return Ledger::restore(std::string(b)); })
      .def("config", [](Ledger& L) { return L.config(); });

  // ABI-style method table + dispatcher-by-signature (reference C:46-52, C:132-167, C:312-318)
  m.def("method_table", [] {
    int n = 0;
    const MethodInfo* t = method_table(&n);
    py::list out;
    for (int i = 0; i < n; ++i)
      out.append(py::make_tuple((int)t[i].id, std::string(t[i].signature), t[i].is_view));
    return out;
  });
  m.def("method_from_signature", [](const std::string& s) { return (int)method_from_signature(s); });

  m.def("sha256_hex", [](const py::bytes& b) {
    std::string s = b;
    return hex(sha256(s.data(), s.size()));
  });
  m.def("status_name", [](Status s) { return std::string(status_name(s)); });

  // Stand-alone access to the shared decision procedure (differential tests vs the oracle
  // and vs the device kernel).
  m.def("run_consensus", [](int n_ranks, int n_comm, int n_aggregate, bool weight_by_score,
                            std::vector<uint32_t> role, std::vector<int> admitted,
                            std::vector<std::vector<float>> score,
                            std::vector<std::vector<int>> scored, std::vector<uint32_t> n_samples,
                            std::vector<float> avg_cost) {
    ConsensusIn<kCMaxRanks> in;
    std::memset(&in, 0, sizeof(in));
    ConsensusOut<kCMaxRanks> out;
    std::memset(&out, 0, sizeof(out));
    in.n_ranks = n_ranks; in.n_comm = n_comm; in.n_aggregate = n_aggregate;
    in.weight_by_score = weight_by_score ? 1 : 0;
    for (int r = 0; r < n_ranks; ++r) {
      in.role[r] = role.at(r); in.admitted[r] = admitted.at(r) ? 1 : 0;
      in.n_samples[r] = n_samples.at(r); in.avg_cost[r] = avg_cost.at(r);
      for (int t = 0; t < n_ranks; ++t) {
        in.scored[r][t] = scored.at(r).at(t) ? 1 : 0;
        in.score[r][t] = score.at(r).at(t);
      }
    }
    run_consensus<kCMaxRanks>(in, out);
    py::dict d;
    std::vector<float> med(out.median, out.median + n_ranks), w(out.weight, out.weight + n_ranks);
    std::vector<int> order(out.order, out.order + out.n_ranked);
    std::vector<int> sel;
    for (int r = 0; r < n_ranks; ++r) if (out.selected[r]) sel.push_back(r);
    std::vector<uint32_t> ra(out.role_after, out.role_after + n_ranks);
    d["median"] = med; d["weight"] = w; d["order"] = order; d["selected"] = sel;
    d["role_after"] = ra; d["global_loss"] = out.global_loss;
    return d;
  });
}
