// Host-side circuit container and fusion planner (no CUDA dependency, so the
// same sources also build into the host emulator used by the CPU tests).
#pragma once
#include <complex>
#include <cstdint>
#include <cstdlib>
#include <memory>
#include <string>
#include <vector>

#include "../../include/qsim_b200.h"
#include "plan.h"

namespace qs {

typedef std::complex<double> cplx;

enum OpKind { OP_DENSE = 0, OP_SIGN = 1 };

// One gate in library coordinates: bits[f] is the index bit that matrix factor
// f (f = 0 most significant) acts on.
struct Op {
  int kind = OP_DENSE;
  int k = 0;
  std::vector<int> bits;
  bool diag = false;
  std::vector<cplx> mat;     // 2^k x 2^k row-major (empty for OP_SIGN)
  // k == 1 unitaries emitted by the merge pass as  [[c,-s],[s,c]] . diag(r0, r1)
  bool rot = false;
  double c = 1.0, s = 0.0;
  cplx r0 = cplx(1.0, 0.0), r1 = cplx(1.0, 0.0);
  uint64_t mask() const {
    uint64_t m = 0;
    for (int b : bits) m |= 1ull << b;
    return m;
  }
};

struct PlanItem {
  bool generic = false;      // a dense block on more than QS_MAX_R qubits (own kernel)
  QsPass pass;               // valid when !generic
  Op op;                     // valid when generic
  // device copy of op.mat, made at the first execution and owned by the plan
  // (libqsim_b200.so only; the deleter is cudaFree)
  mutable std::shared_ptr<void> dev_mat;
  mutable int dev_index = -1;
};

}  // namespace qs

struct qsim_circuit {
  int n = 0;
  std::vector<qs::Op> ops;
};

struct qsim_plan {
  int n = 0;
  std::vector<qs::PlanItem> items;
  std::vector<double> residual;          // n x 8 doubles (index bit b first), empty unless defer_tail
  qsim_plan_stats_t stats{};
};

namespace qs {

// Development knobs (A/B switches behind the measurements in DESIGN.md section 5) read the
// environment only in builds with -DQSIM_DEV_KNOBS; the shipped library ignores it.  Every knob
// selects between CORRECT variants; none changes results.
inline int dev_knob(const char* name, int fallback) {
#if defined(QSIM_DEV_KNOBS)
  const char* e = getenv(name);
  return e ? atoi(e) : fallback;
#else
  (void)name;
  return fallback;
#endif
}

void set_error(const std::string& msg);
int fail(int code, const std::string& msg);

// Single-qubit merging pre-pass (returns the reduced op list).
// With `residual` the trailing diagonal / antidiagonal product of every index bit is not
// emitted but written there (8 doubles per bit, identity where nothing is pending).
// Index bits in `apply_mask` are exempt.
std::vector<Op> merge_single_qubit(int n, const std::vector<Op>& in, std::vector<double>* residual = nullptr,
                                   uint64_t apply_mask = 0);

// Greedy tile/pass construction.
int build_plan(int n, const std::vector<Op>& ops, const qsim_plan_options_t& opt, qsim_plan* out);

qsim_plan_options_t resolve_options(const qsim_plan_options_t* opt);

}  // namespace qs
