// Circuit container, single-qubit merging and the greedy fusion planner.
//
// Replaces the per-gate Python loop of the reference executor
// (DV/simulator.py:40-52): instead of one full-state operation per gate, gates
// are packed into *passes* (plan.h).  Three facts do the work:
//   * runs of single-qubit gates on one qubit collapse into one 2x2 matrix, and
//     diagonal ones slide through CZ/Z gates to reach a neighbour to merge with;
//   * CZ/Z are signs that depend only on index bits, so they cost no memory
//     traffic and may involve qubits outside the tile;
//   * non-diagonal gates need their qubits inside the tile, so each pass picks
//     the T tile bits that let it absorb the most pending gates.
#include "planner.h"
#include "tile_exec.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace qs {

static thread_local std::string g_error;

void set_error(const std::string& msg) { g_error = msg; }
int fail(int code, const std::string& msg) {
  g_error = msg;
  return code;
}
const char* last_error() { return g_error.c_str(); }

qsim_plan_options_t resolve_options(const qsim_plan_options_t* opt) {
  qsim_plan_options_t o{};
  if (opt) o = *opt;
  if (o.tile_bits <= 0) o.tile_bits = 12;
  if (o.low_bits <= 0) o.low_bits = 4;
  if (o.max_group <= 0) o.max_group = 4;
  if (o.max_dense_ops <= 0) o.max_dense_ops = 20;    // 20 -> 0.54 of the HBM roof at 14.4k gates/s; 28 -> 0.50 at 14.6k
  if (o.lookahead <= 0) o.lookahead = 600;
  if (o.merge_1q <= 0) o.merge_1q = 1;
  if (o.max_layers <= 0) o.max_layers = 3;
  if (o.max_layers > 8) o.max_layers = 8;
  if (o.cta_log2 != QS_THREADS_LOG2_MIN && o.cta_log2 != QS_THREADS_LOG2) o.cta_log2 = 0;     // 0: per pass
  if (o.defer_tail != 1 || o.merge_1q != 1) o.defer_tail = 0;
  if (o.tile_bits > QS_MAX_T) o.tile_bits = QS_MAX_T;
  if (o.max_group > QS_MAX_R) o.max_group = QS_MAX_R;
  if (o.low_bits + QS_MAX_R > o.tile_bits) o.low_bits = o.tile_bits - QS_MAX_R;
  if (o.low_bits < 0) o.low_bits = 0;
  return o;
}

// ---- small dense helpers -------------------------------------------------------

static bool is_diagonal(const std::vector<cplx>& m, int dim) {
  for (int r = 0; r < dim; ++r)
    for (int c = 0; c < dim; ++c)
      if (r != c && (m[r * dim + c].real() != 0.0 || m[r * dim + c].imag() != 0.0)) return false;
  return true;
}

static bool is_exact(const cplx& v, double re) { return v.real() == re && v.imag() == 0.0; }

static bool is_cz(const std::vector<cplx>& m) {
  if (!is_diagonal(m, 4)) return false;
  return is_exact(m[0], 1.0) && is_exact(m[5], 1.0) && is_exact(m[10], 1.0) && is_exact(m[15], -1.0);
}

static bool is_z(const std::vector<cplx>& m) {
  return is_diagonal(m, 2) && is_exact(m[0], 1.0) && is_exact(m[3], -1.0);
}

static bool is_antidiagonal(const std::vector<cplx>& m) {   // 2x2 only
  return m[0].real() == 0.0 && m[0].imag() == 0.0 && m[3].real() == 0.0 && m[3].imag() == 0.0;
}

static bool is_identity(const std::vector<cplx>& m, int dim) {
  if (!is_diagonal(m, dim)) return false;
  for (int r = 0; r < dim; ++r)
    if (!is_exact(m[r * dim + r], 1.0)) return false;
  return true;
}

// U = diag(l0, l1) . [[c, -s], [s, c]] . diag(1, r1)  for a 2x2 unitary with no zero
// entry; returns false (and leaves the outputs alone) when U is not of that form to
// within rounding, e.g. a non-unitary user matrix.
static bool factor_rotation(const std::vector<cplx>& u, cplx& l0, cplx& l1, double& c, double& s, cplx& r1) {
  const double c0 = std::abs(u[0]), s0 = std::abs(u[2]);
  if (!(c0 > 1e-8) || !(s0 > 1e-8)) return false;
  const cplx tl0 = u[0] / c0, tl1 = u[2] / s0;
  const cplx tr1 = u[3] / (c0 * tl1);
  // reconstruct and compare
  const cplx v01 = -tl0 * s0 * tr1, v11 = tl1 * c0 * tr1;
  const double scale = std::max(std::max(std::abs(u[0]), std::abs(u[1])), std::max(std::abs(u[2]), std::abs(u[3])));
  const double err = std::max(std::abs(v01 - u[1]), std::abs(v11 - u[3]));
  if (!(err <= 4e-15 * scale)) return false;
  if (std::abs(std::abs(tr1) - 1.0) > 1e-13 || std::abs(std::abs(tl0) - 1.0) > 1e-13) return false;
  l0 = tl0; l1 = tl1; c = c0; s = s0; r1 = tr1;
  return true;
}

// 2x2 product a*b
static std::vector<cplx> mul2(const std::vector<cplx>& a, const std::vector<cplx>& b) {
  std::vector<cplx> o(4);
  o[0] = a[0] * b[0] + a[1] * b[2];
  o[1] = a[0] * b[1] + a[1] * b[3];
  o[2] = a[2] * b[0] + a[3] * b[2];
  o[3] = a[2] * b[1] + a[3] * b[3];
  return o;
}

// (g on factor f) * m   -- g applied AFTER m
static void fold_left(Op& op, int f, const std::vector<cplx>& g) {
  const int k = op.k, dim = 1 << k, bit = 1 << (k - 1 - f);
  std::vector<cplx> o(op.mat.size());
  for (int r = 0; r < dim; ++r) {
    const int rf = (r & bit) ? 1 : 0;
    const int r0 = r & ~bit, r1 = r | bit;
    for (int c = 0; c < dim; ++c)
      o[r * dim + c] = g[rf * 2 + 0] * op.mat[r0 * dim + c] + g[rf * 2 + 1] * op.mat[r1 * dim + c];
  }
  op.mat.swap(o);
  op.diag = is_diagonal(op.mat, dim);
  op.rot = false;
}

// m * (g on factor f)   -- g applied BEFORE m
static void fold_right(Op& op, int f, const std::vector<cplx>& g) {
  const int k = op.k, dim = 1 << k, bit = 1 << (k - 1 - f);
  std::vector<cplx> o(op.mat.size());
  for (int r = 0; r < dim; ++r)
    for (int c = 0; c < dim; ++c) {
      const int cf = (c & bit) ? 1 : 0;
      const int c0 = c & ~bit, c1 = c | bit;
      o[r * dim + c] = op.mat[r * dim + c0] * g[0 * 2 + cf] + op.mat[r * dim + c1] * g[1 * 2 + cf];
    }
  op.mat.swap(o);
  op.diag = is_diagonal(op.mat, dim);
}

// ---- single-qubit merging ---------------------------------------------------------

namespace {
enum Since { SINCE_NOTHING = 0, SINCE_DIAG = 1, SINCE_BLOCKED = 2 };
struct BitTrack {
  bool has_pending = false;
  std::vector<cplx> pending;   // product of not-yet-emitted 1q gates on this bit
  bool pending_diag = true;
  int last_op = -1;            // output op a later 1q gate may still fold into
  int last_factor = 0;
  int since = SINCE_BLOCKED;   // what has been emitted on this bit after last_op
};
}  // namespace

std::vector<Op> merge_single_qubit(int n, const std::vector<Op>& in, std::vector<double>* residual,
                                   uint64_t apply_mask) {
  std::vector<Op> out;
  out.reserve(in.size());
  std::vector<BitTrack> tr(n);
  static const bool quarter_turns = dev_knob("QSIM_NO_QUARTER_TURNS", 0) == 0;

  auto emit_pending = [&](int b) {
    BitTrack& t = tr[b];
    if (!t.has_pending) return;
    t.has_pending = false;
    if (is_identity(t.pending, 2)) return;
    Op o;
    if (is_z(t.pending)) {
      o.kind = OP_SIGN;
      o.k = 2;
      o.bits = {b, b};
      o.diag = true;
      out.push_back(o);
      if (t.since == SINCE_NOTHING) t.since = SINCE_DIAG;
      return;
    }
    o.kind = OP_DENSE;
    o.k = 1;
    o.bits = {b};
    o.mat = t.pending;
    o.diag = t.pending_diag;
    // A general unitary leaves as (rotation . right phases); its left phases stay
    // behind as a new pending diagonal that slides on to the next gate of the qubit.
    cplx l0, l1, r1;
    double c, s;
    if (!o.diag && !is_antidiagonal(o.mat) && factor_rotation(o.mat, l0, l1, c, s, r1)) {
      o.rot = true;
      o.r0 = cplx(1.0, 0.0); o.r1 = r1;
      t.has_pending = true;
      if (s <= c || !quarter_turns) {
        o.c = c; o.s = s;
        o.mat = {cplx(c, 0.0), -s * r1, cplx(s, 0.0), c * r1};
        t.pending = {l0, cplx(0.0, 0.0), cplx(0.0, 0.0), l1};
        t.pending_diag = true;
      } else {
        // More than 45 degrees: split off a quarter turn,
        //   [[c, -s], [s, c]] = [[0, -1], [1, 0]] . [[s, c], [-c, s]],
        // which stays behind with the left phases as an ANTIDIAGONAL pending product (it slides
        // through CZ/Z like any bit flip, see below).  What is emitted turns by at most 45
        // degrees, the two-shear form of the kernel (plan.h).
        o.c = s; o.s = -c;
        o.mat = {cplx(s, 0.0), c * r1, cplx(-c, 0.0), s * r1};
        t.pending = {cplx(0.0, 0.0), -l0, l1, cplx(0.0, 0.0)};
        t.pending_diag = false;
      }
    }
    out.push_back(o);
    t.last_op = (int)out.size() - 1;
    t.last_factor = 0;
    t.since = SINCE_NOTHING;
  };

  for (const Op& g : in) {
    if (g.kind == OP_DENSE && g.k == 1) {
      const int b = g.bits[0];
      BitTrack& t = tr[b];
      if (t.has_pending) {
        t.pending = mul2(g.mat, t.pending);
        t.pending_diag = is_diagonal(t.pending, 2);
      } else if (t.last_op >= 0 && (t.since == SINCE_NOTHING || (g.diag && t.since == SINCE_DIAG))) {
        fold_left(out[t.last_op], t.last_factor, g.mat);
      } else {
        t.has_pending = true;
        t.pending = g.mat;
        t.pending_diag = g.diag;
      }
      continue;
    }
    if (g.kind == OP_SIGN) {
      const int a = g.bits[0], b = g.bits[1];
      // Pending products that are diagonal slide through a CZ/Z unchanged.  Pending
      // ANTIDIAGONAL products (X times a diagonal) slide through as well, leaving
      // Pauli-frame debris behind:  CZ_ab P_a = P_a CZ_ab Z_b  and  Z_a P_a = -P_a Z_a.
      // So bit flips never cost a pass of their own: they ride along until the next
      // non-diagonal gate on the qubit absorbs them.
      for (int x : g.bits) {
        BitTrack& t = tr[x];
        if (t.has_pending && !t.pending_diag && !is_antidiagonal(t.pending)) emit_pending(x);
      }
      const bool anti_a = tr[a].has_pending && !tr[a].pending_diag;
      const bool anti_b = tr[b].has_pending && !tr[b].pending_diag;
      for (int x : g.bits)
        if (tr[x].since == SINCE_NOTHING) tr[x].since = SINCE_DIAG;
      out.push_back(g);
      auto emit_z = [&](int x) {
        Op z;
        z.kind = OP_SIGN;
        z.k = 2;
        z.bits = {x, x};
        z.diag = true;
        out.push_back(z);
      };
      if (a == b) {
        if (anti_a)
          for (cplx& v : tr[a].pending) v = -v;
      } else {
        if (anti_a) emit_z(b);
        if (anti_b) emit_z(a);
        if (anti_a && anti_b)
          for (cplx& v : tr[a].pending) v = -v;
      }
      continue;
    }
    // multi-qubit dense gate: swallow pending single-qubit gates on its qubits
    Op d = g;
    for (int f = 0; f < d.k; ++f) {
      BitTrack& t = tr[d.bits[f]];
      if (t.has_pending) {
        fold_right(d, f, t.pending);
        t.has_pending = false;
      }
    }
    out.push_back(d);
    const int idx = (int)out.size() - 1;
    for (int f = 0; f < d.k; ++f) {
      BitTrack& t = tr[d.bits[f]];
      if (d.diag) {
        // a later diagonal 1q gate commutes with d and may still reach last_op
        if (t.since == SINCE_NOTHING) t.since = SINCE_DIAG;
        if (t.last_op < 0) { t.last_op = idx; t.last_factor = f; t.since = SINCE_NOTHING; }
      } else {
        t.last_op = idx;
        t.last_factor = f;
        t.since = SINCE_NOTHING;
      }
    }
  }
  if (residual) residual->assign((size_t)8 * n, 0.0);
  for (int b = 0; b < n; ++b) {
    if (!residual || (apply_mask >> b & 1)) {
      emit_pending(b);      // a general gate leaves its left phases pending ...
      emit_pending(b);      // ... which go out as a diagonal gate
      if (residual) { (*residual)[(size_t)8 * b] = 1.0; (*residual)[(size_t)8 * b + 6] = 1.0; }
      continue;
    }
    BitTrack& t = tr[b];
    if (t.has_pending && !t.pending_diag && !is_antidiagonal(t.pending)) emit_pending(b);
    double* r = residual->data() + (size_t)8 * b;
    if (t.has_pending) {
      for (int e = 0; e < 4; ++e) { r[2 * e] = t.pending[e].real(); r[2 * e + 1] = t.pending[e].imag(); }
      t.has_pending = false;
    } else {
      r[0] = 1.0; r[6] = 1.0;
    }
  }
  return out;
}

// ---- greedy pass construction --------------------------------------------------------

namespace {

constexpr int kMaxLayerPairs = 96;    // sign pairs attached to one layer
constexpr int kMaxLo = 48;            // of which (local, outer) pairs
constexpr int kMaxStepLayers = 8;     // hard cap on layers per step (options.max_layers <= this)

// A layer / step while the walk is still adding gates to them.
struct WLayer {
  uint8_t kind = QS_LAYER_ROT;
  uint8_t form[QS_MAX_R] = {0, 0, 0, 0};
  double coef[QS_MAX_R][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};   // (u, t) for TAN, (p, q) for SHEAR3
  cplx ph[QS_MAX_R][2];                         // phases (and scale) on the factor's 0 / 1 amplitudes
  bool has_ph = false;
  int op_idx[QS_MAX_R] = {-1, -1, -1, -1};      // QS_FORM_FULL members / [0]: the dense op
  int npairs = 0, nlo = 0;
  uint8_t pairs[kMaxLayerPairs][2];             // global bits, a <= b (a == b: a Z)
};
struct WStep {
  int r = 0;
  int bits[QS_MAX_R] = {0, 0, 0, 0};            // global bit of factor f
  bool dense = false;
  int nlayers = 0;
  int layers[kMaxStepLayers];
};

// How a single-qubit op enters a layer.
struct OneQ {
  uint8_t form = QS_FORM_NONE;
  double c0 = 0.0, c1 = 0.0;                    // (u, t) for TAN, (p, q) for SHEAR3
  cplx ph0 = cplx(1.0, 0.0), ph1 = cplx(1.0, 0.0);
  bool full = false;                            // needs a QS_LAYER_GENERAL layer
};

// [[c, -s], [s, c]] . diag(r0, r1) with c >= 0, c^2 + s^2 = 1 (plan.h); s < 0 only with c >= |s|
OneQ rotation_form(double c, double s, cplx r0, cplx r1) {
  OneQ q;
  if (c >= std::abs(s)) {
    const double t = s / c, d = 1.0 + t * t;
    q.form = QS_FORM_TAN; q.c0 = t / d; q.c1 = t;
    q.ph0 = c * r0; q.ph1 = (c * d) * r1;
  } else {
    q.form = QS_FORM_SHEAR3; q.c0 = s / (1.0 + c); q.c1 = s;
    q.ph0 = r0; q.ph1 = r1;
  }
  return q;
}

OneQ classify_1q(const Op& op) {
  auto zero = [](const cplx& v) { return v.real() == 0.0 && v.imag() == 0.0; };
  if (op.rot) return rotation_form(op.c, op.s, op.r0, op.r1);
  OneQ q;
  if (zero(op.mat[1]) && zero(op.mat[2])) {
    q.form = QS_FORM_NONE; q.ph0 = op.mat[0]; q.ph1 = op.mat[3];           // diagonal: phases only
  } else if (zero(op.mat[0]) && zero(op.mat[3])) {
    // [[0, b], [c, 0]] = [[0, -1], [1, 0]] . diag(c, -b): a quarter turn
    q = rotation_form(0.0, 1.0, op.mat[2], -op.mat[1]);
  } else {
    q.form = QS_FORM_FULL; q.full = true;
  }
  return q;
}

struct Walker {
  int n;
  const std::vector<Op>& ops;
  const std::vector<uint64_t>& masks;
  std::vector<char>& done;
  qsim_plan_options_t opt;

  // Walk the pending ops in order and take every op that can run in a pass whose
  // tile is the bit set S.  With `pass` == nullptr only the score is computed
  // (and at most opt.lookahead pending ops are visited).
  // Returns dense_taken * 4096 + min(sign_taken, 4095).
  //
  // Positions inside a pass are (step, layer) pairs, ordered lexicographically and
  // encoded as step * 64 + layer.
  long walk(size_t first, uint64_t S, QsPass* pass, std::vector<size_t>* taken_idx) const {
    const uint64_t all = (n >= 64) ? ~0ull : ((1ull << n) - 1ull);
    uint64_t blocked_full = 0, blocked_diag = 0;
    int dense_taken = 0, sign_taken = 0, visited = 0;
    const int max_layers = std::min(opt.max_layers, kMaxStepLayers);
    static const bool pad_steps = dev_knob("QSIM_NO_STEP_PADDING", 0) == 0;
    static const bool pair_dense = dev_knob("QSIM_NO_DENSE_PAIRS", 0) == 0;

    int lpos[64];
    {
      int t = 0;
      for (int b = 0; b < n; ++b) lpos[b] = (S >> b & 1) ? t++ : -1;
      for (int b = n; b < 64; ++b) lpos[b] = -1;
    }

    WStep steps[QS_MAX_STEPS];
    WLayer layers[QS_MAX_LAYERS];
    int nsteps = 0, nlayers = 0, ncoef = 0, total_lo = 0;
    // one layer is held back for the final sign layer
    const int layer_cap = QS_MAX_LAYERS - 1;
    const int coef_cap = QS_MAX_COEF - 2 * (1 << QS_MAX_R);      // ... and room for its phase table
    int last_dense[64], last_sign[64];
    for (int b = 0; b < 64; ++b) { last_dense[b] = -1; last_sign[b] = 0; }
    // Sign pairs taken but not yet attached: each waits for the first later layer
    // whose group contains one of its bits, else for the pass's final layer.
    uint64_t pend_mask = 0;
    uint8_t pend[256][2];
    int npend = 0;

    // Earliest position that a matrix on bit c may take, given the pending pairs that
    // touch c (they would be attached to that layer's sign block, so the layer must come
    // after every earlier matrix on their partner bits).
    auto earliest_pos = [&](int c) {
      int e = std::max(last_dense[c] + 1, last_sign[c]);
      if (pend_mask >> c & 1)
        for (int p = 0; p < npend; ++p) {
          const int a = pend[p][0], b = pend[p][1];
          if (a == c && b != c) e = std::max(e, last_dense[b] + 1);
          else if (b == c && a != c) e = std::max(e, last_dense[a] + 1);
        }
      return e;
    };
    auto pending_count = [&](int c, int* lo) {
      int cnt = 0;
      *lo = 0;
      if (pend_mask >> c & 1)
        for (int p = 0; p < npend; ++p) {
          const int a = pend[p][0], b = pend[p][1];
          if (a != c && b != c) continue;
          ++cnt;
          if (a != b && lpos[a == c ? b : a] < 0) ++*lo;
        }
      return cnt;
    };
    // Move every pending pair that touches bit c into layer `li` (position `pos`).
    auto attach = [&](int c, int li, int pos) {
      if (!(pend_mask >> c & 1)) return;
      WLayer& L = layers[li];
      int kept = 0;
      uint64_t new_mask = 0;
      for (int p = 0; p < npend; ++p) {
        const int a = pend[p][0], b = pend[p][1];
        if (a != c && b != c) {
          pend[kept][0] = (uint8_t)a; pend[kept][1] = (uint8_t)b; ++kept;
          new_mask |= (1ull << a) | (1ull << b);
          continue;
        }
        const int e = (a == c) ? b : a;              // partner (== c for a Z)
        last_sign[c] = std::max(last_sign[c], pos);
        last_sign[e] = std::max(last_sign[e], pos);
        L.pairs[L.npairs][0] = (uint8_t)a;
        L.pairs[L.npairs][1] = (uint8_t)b;
        L.npairs++;
        if (e != c && lpos[e] < 0) { L.nlo++; ++total_lo; }
      }
      npend = kept;
      pend_mask = new_mask;
    };
    auto layer_coef_need = [&](bool full) {
      return (full ? 8 * opt.max_group : 2 * opt.max_group) + 2 * (1 << opt.max_group);
    };
    auto new_layer = [&](int sidx, uint8_t kind) {
      WLayer& L = layers[nlayers];
      L = WLayer();
      L.kind = kind;
      for (int f = 0; f < QS_MAX_R; ++f) { L.ph[f][0] = cplx(1.0, 0.0); L.ph[f][1] = cplx(1.0, 0.0); }
      steps[sidx].layers[steps[sidx].nlayers++] = nlayers;
      return nlayers++;
    };

    for (size_t i = first; i < ops.size(); ++i) {
      if (done[i]) continue;
      if (!pass && ++visited > opt.lookahead) break;
      const Op& op = ops[i];
      const uint64_t m = masks[i];
      if (op.kind == OP_SIGN) {
        const bool room = total_lo + npend + 1 <= QS_MAX_PAIRS && npend < 250;
        if ((m & blocked_full) == 0 && room) {
          // CZ is an involution: a repeated pending pair cancels
          const int a = std::min(op.bits[0], op.bits[1]), b = std::max(op.bits[0], op.bits[1]);
          int hit = -1;
          for (int p = 0; p < npend; ++p)
            if (pend[p][0] == a && pend[p][1] == b) { hit = p; break; }
          if (hit >= 0) {
            pend[hit][0] = pend[npend - 1][0]; pend[hit][1] = pend[npend - 1][1];
            --npend;
          } else {
            pend[npend][0] = (uint8_t)a; pend[npend][1] = (uint8_t)b;
            ++npend;
          }
          pend_mask |= m;      // conservative after a cancellation
          ++sign_taken;
          if (taken_idx) taken_idx->push_back(i);
        } else {
          blocked_diag |= m;
        }
        if ((blocked_full & all) == all) break;
        continue;
      }
      const bool in_tile = (m & ~S) == 0;
      const bool free_bits = (m & blocked_full) == 0 && (op.diag || (m & blocked_diag) == 0);
      bool take = in_tile && free_bits && op.k <= QS_MAX_R && dense_taken < opt.max_dense_ops;
      if (take && op.k == 1) {
        const int c = op.bits[0];
        const OneQ q = classify_1q(op);
        const uint8_t want = q.full ? QS_LAYER_GENERAL : QS_LAYER_ROT;
        int nlo = 0;
        const int cnt = pending_count(c, &nlo);
        const int e = earliest_pos(c);
        int js = -1, jl = -1;             // chosen step / layer-in-step
        bool fresh_layer = false;
        // earliest fit: a round trip costs about four layers, and an early position keeps the
        // qubit (and its CZ partners) free for what follows
        for (int sidx = e >> 6; sidx < nsteps && js < 0; ++sidx) {
          const WStep& st = steps[sidx];
          if (st.dense) continue;
          bool member = false;
          for (int f = 0; f < st.r; ++f) member |= st.bits[f] == c;
          if (!member && st.r >= opt.max_group) continue;
          for (int l = (sidx == (e >> 6)) ? (e & 63) : 0; l <= st.nlayers; ++l) {
            if (l < st.nlayers) {
              const WLayer& L = layers[st.layers[l]];
              if (L.kind != want) continue;
              if (L.npairs + cnt > kMaxLayerPairs || L.nlo + nlo > kMaxLo) continue;
              js = sidx; jl = l;
              break;
            }
            // a new layer at the end of this step
            if (st.nlayers >= max_layers || nlayers >= layer_cap) break;
            if (ncoef + layer_coef_need(q.full) > coef_cap) break;
            if (cnt > kMaxLayerPairs || nlo > kMaxLo) break;
            js = sidx; jl = l; fresh_layer = true;
            break;
          }
        }
        if (js < 0) {
          // a new step
          if (nsteps >= QS_MAX_STEPS || nlayers >= layer_cap || ncoef + layer_coef_need(q.full) > coef_cap ||
              cnt > kMaxLayerPairs || nlo > kMaxLo || total_lo + nlo + npend > QS_MAX_PAIRS) {
            take = false;
          } else {
            steps[nsteps] = WStep();
            js = nsteps++;
            jl = 0;
            fresh_layer = true;
          }
        }
        if (take) {
          WStep& st = steps[js];
          int li;
          if (fresh_layer) {
            li = new_layer(js, want);
            ncoef += layer_coef_need(q.full);
          } else {
            li = st.layers[jl];
          }
          int f = -1;
          for (int g = 0; g < st.r; ++g)
            if (st.bits[g] == c) f = g;
          if (f < 0) { f = st.r; st.bits[st.r++] = c; }
          WLayer& L = layers[li];
          L.form[f] = q.form;
          L.coef[f][0] = q.c0;
          L.coef[f][1] = q.c1;
          if (q.full) {
            L.op_idx[f] = (int)i;
          } else {
            L.ph[f][0] = q.ph0; L.ph[f][1] = q.ph1;
            if (q.ph0 != cplx(1.0, 0.0) || q.ph1 != cplx(1.0, 0.0)) L.has_ph = true;
          }
          const int pos = js * 64 + jl;
          attach(c, li, pos);
          last_dense[c] = pos;
        }
      } else if (take) {
        // dense k >= 2 gate: a step of its own with a single layer
        const int dim = 1 << op.k;
        int nlo = 0, cnt = 0;
        for (int b : op.bits) { int lo1 = 0; cnt += pending_count(b, &lo1); nlo += lo1; }
        // A 4x4 block may share the round trip of an earlier 4x4 block on two other bits (paired dense
        // layer, plan.h): the step must come after everything this op depends on, i.e. after the last
        // matrix on its bits and on the partners of its pending pairs.
        int join = -1;
        if (op.k == 2 && opt.max_group >= 4 && pair_dense) {
          const int e = std::max(earliest_pos(op.bits[0]), earliest_pos(op.bits[1]));
          for (int sidx = (e + 63) >> 6; sidx < nsteps; ++sidx) {
            const WStep& st = steps[sidx];
            if (!st.dense || st.r != 2 || st.nlayers != 1) continue;
            if (st.bits[0] == op.bits[0] || st.bits[0] == op.bits[1] || st.bits[1] == op.bits[0] ||
                st.bits[1] == op.bits[1])
              continue;
            const WLayer& L = layers[st.layers[0]];
            if (L.npairs + cnt > kMaxLayerPairs || L.nlo + nlo > kMaxLo) continue;
            join = sidx;
            break;
          }
        }
        if (join >= 0 && ncoef + 2 * dim * dim <= coef_cap && total_lo + nlo + npend <= QS_MAX_PAIRS) {
          WStep& st = steps[join];
          st.bits[2] = op.bits[0];
          st.bits[3] = op.bits[1];
          st.r = 4;
          const int li = st.layers[0];
          layers[li].op_idx[1] = (int)i;
          ncoef += 2 * dim * dim;
          const int pos = join * 64;
          for (int f = 0; f < op.k; ++f) attach(op.bits[f], li, pos);
          for (int b : op.bits) last_dense[b] = pos;
        } else if (nsteps >= QS_MAX_STEPS || nlayers >= layer_cap || ncoef + 2 * dim * dim > coef_cap ||
            cnt > kMaxLayerPairs || nlo > kMaxLo || total_lo + nlo + npend > QS_MAX_PAIRS) {
          take = false;
        } else {
          steps[nsteps] = WStep();
          WStep& st = steps[nsteps];
          st.dense = true;
          st.r = op.k;
          for (int f = 0; f < op.k; ++f) st.bits[f] = op.bits[f];
          const int li = new_layer(nsteps, QS_LAYER_DENSE);
          layers[li].op_idx[0] = (int)i;
          ncoef += 2 * dim * dim;
          const int pos = nsteps * 64;
          for (int f = 0; f < op.k; ++f) attach(op.bits[f], li, pos);
          for (int b : op.bits) last_dense[b] = pos;
          ++nsteps;
        }
      }
      if (take) {
        ++dense_taken;
        if (taken_idx) taken_idx->push_back(i);
      } else if (op.diag) {
        blocked_diag |= m;
      } else {
        blocked_full |= m;
      }
      if ((blocked_full & all) == all) break;
    }

    if (pass && pad_steps) {
      // A step with fewer group bits than the widest one runs a smaller instantiation of the step
      // body, and ptxas keeps only the widest one's layer loop on the uniform datapath (DESIGN.md
      // section 5).  An idle group bit costs nothing (no rotation, phase 1, one more popcount), so
      // narrow rotation steps are widened with tile bits they do not use, highest positions first
      // (positions below 6 are the ones the conflict-free thread maps need).
      int widest = 0;
      for (int sidx = 0; sidx < nsteps; ++sidx) widest = std::max(widest, steps[sidx].r);
      for (int sidx = 0; sidx < nsteps; ++sidx) {
        WStep& st = steps[sidx];
        if (st.dense) continue;
        for (int b = n - 1; b >= 0 && st.r < widest; --b) {
          if (!(S >> b & 1) || lpos[b] < 6) continue;
          bool member = false;
          for (int f = 0; f < st.r; ++f) member |= st.bits[f] == b;
          if (!member) st.bits[st.r++] = b;
        }
      }
    }
    if (pass) finalize(*pass, S, lpos, steps, nsteps, layers, pend, npend);
    return (long)dense_taken * 4096 + std::min(sign_taken, 4095);
  }

  // Turn the walk's steps / layers / leftover pairs into the pass descriptor.
  void finalize(QsPass& P, uint64_t S, const int* lpos, WStep* steps, int nsteps, WLayer* layers,
                const uint8_t (*pend)[2], int npend) const {
    int nlayers_used = 0;
    for (int s = 0; s < nsteps; ++s) nlayers_used += steps[s].nlayers;
    // leftover pairs: a sign-only final layer at the end of the last step
    WLayer fin;
    bool has_final = npend > 0;
    if (has_final) {
      if (nsteps == 0) {
        steps[0] = WStep();
        steps[0].r = 1;
        for (int b = n - 1; b >= 0; --b)
          if (S >> b & 1) { steps[0].bits[0] = b; break; }
        nsteps = 1;
      }
      fin = WLayer();
      fin.kind = QS_LAYER_ROT;
      for (int f = 0; f < QS_MAX_R; ++f) { fin.ph[f][0] = cplx(1.0, 0.0); fin.ph[f][1] = cplx(1.0, 0.0); }
    }
    P.nsteps = (uint32_t)nsteps;
    int ncoef = 0, npairs = 0, nl = 0;
    bool pair_of[QS_MAX_LAYERS + 1] = {};          // layer holds two 4x4 blocks
    for (int s = 0; s < nsteps; ++s) {
      WStep& ws = steps[s];
      QsStep& st = P.steps[s];
      st = QsStep{};
      st.r = (uint8_t)ws.r;
      st.layer0 = (uint8_t)nl;
      uint32_t gmask = 0;
      for (int f = 0; f < ws.r; ++f) {
        st.gpos[f] = (uint8_t)lpos[ws.bits[f]];
        gmask |= 1u << st.gpos[f];
      }
      const bool last = s == nsteps - 1;
      const int count = ws.nlayers + ((last && has_final) ? 1 : 0);
      for (int k = 0; k < count; ++k) {
        const bool is_fin = k == ws.nlayers;
        const WLayer& wl = is_fin ? fin : layers[ws.layers[k]];
        QsLayer& L = P.layers[nl++];
        L = QsLayer{};
        L.kind = wl.kind;
        L.step = (uint8_t)s;
        const int r = ws.r, na = 1 << r;
        // coefficients
        const bool paired = wl.kind == QS_LAYER_DENSE && wl.op_idx[1] >= 0;
        pair_of[nl - 1] = paired;
        if (paired) {
          L.coef_off = (uint16_t)ncoef;
          for (int blk = 0; blk < 2; ++blk) {
            const Op& op = ops[wl.op_idx[blk]];
            for (int e = 0; e < 16; ++e) {
              P.coef[ncoef++] = op.mat[e].real();
              P.coef[ncoef++] = op.mat[e].imag();
            }
          }
        } else if (wl.kind == QS_LAYER_DENSE) {
          const Op& op = ops[wl.op_idx[0]];
          L.coef_off = (uint16_t)ncoef;
          for (int e = 0; e < na * na; ++e) {
            P.coef[ncoef++] = op.mat[e].real();
            P.coef[ncoef++] = op.mat[e].imag();
          }
        } else {
          L.coef_off = (uint16_t)ncoef;
          for (int f = 0; f < r; ++f) L.form[f] = wl.form[f];
          if (wl.kind == QS_LAYER_GENERAL) {
            for (int f = 0; f < r; ++f) {
              double* dst = P.coef + ncoef + 8 * f;
              for (int e = 0; e < 8; ++e) dst[e] = 0.0;
              if (wl.form[f] != QS_FORM_FULL) continue;
              const Op& op = ops[wl.op_idx[f]];
              for (int e = 0; e < 4; ++e) { dst[2 * e] = op.mat[e].real(); dst[2 * e + 1] = op.mat[e].imag(); }
            }
            ncoef += 8 * r;
          } else {
            for (int f = 0; f < r; ++f) { P.coef[ncoef + 2 * f] = wl.coef[f][0]; P.coef[ncoef + 2 * f + 1] = wl.coef[f][1]; }
            ncoef += 2 * r;
          }
        }
        // sign block
        L.pair_off = (uint16_t)npairs;
        bool coupled[QS_MAX_R][QS_MAX_R] = {};
        auto add_pair = [&](int a, int b) {
          if (a == b) { L.zconst ^= (uint16_t)(1u << lpos[a]); return; }
          const int pa = lpos[a], pb = lpos[b];
          const bool ga = pa >= 0 && (gmask >> pa & 1), gb = pb >= 0 && (gmask >> pb & 1);
          auto factor_of = [&](int p) { for (int f = 0; f < r; ++f) if (st.gpos[f] == p) return f; return -1; };
          if (ga && gb) {
            const int fa = factor_of(pa), fb = factor_of(pb);
            coupled[fa][fb] ^= true; coupled[fb][fa] ^= true;
          } else if (ga || gb) {
            const int pg = ga ? pa : pb, po = ga ? pb : pa, bo = ga ? b : a;
            if (po >= 0) {
              const int fg = factor_of(pg);
              L.ngp[fg >> 1] ^= (1u << po) << (16 * (fg & 1));
            } else {
              P.pairs[2 * npairs] = (uint8_t)pg; P.pairs[2 * npairs + 1] = (uint8_t)bo; ++npairs; L.n_lo++;
            }
          } else if (pa >= 0 && pb >= 0) {         // final layer only: no group bit involved
            P.fin_nsym[pa] ^= (uint16_t)(1u << pb);
            P.fin_nsym[pb] ^= (uint16_t)(1u << pa);
          } else {                                  // final layer only: (local, outer); z is kept by position
            const int pl = pa >= 0 ? pa : pb, bo = pa >= 0 ? b : a;
            P.pairs[2 * npairs] = (uint8_t)pl; P.pairs[2 * npairs + 1] = (uint8_t)bo; ++npairs; L.n_lo++;
          }
        };
        if (is_fin) {
          // (outer, outer) pairs first: they only feed the tile-uniform bit g
          L.flags |= QS_LF_SIGN | QS_LF_FINAL;
          P.has_final = 1;
          P.fin_pair_off = (uint16_t)npairs;
          for (int p = 0; p < npend; ++p) {           // a Z on an outer bit is the pair (a, a)
            const int a = pend[p][0], b = pend[p][1];
            if (lpos[a] < 0 && lpos[b] < 0) {
              P.pairs[2 * npairs] = (uint8_t)a; P.pairs[2 * npairs + 1] = (uint8_t)b; ++npairs; P.fin_n_oo++;
            }
          }
          L.pair_off = (uint16_t)npairs;
          for (int p = 0; p < npend; ++p) {
            const int a = pend[p][0], b = pend[p][1];
            if (lpos[a] < 0 && lpos[b] < 0) continue;
            add_pair(a, b);
          }
        } else {
          if (wl.npairs > 0) L.flags |= QS_LF_SIGN;
          for (int p = 0; p < wl.npairs; ++p) add_pair(wl.pairs[p][0], wl.pairs[p][1]);
        }
        // sign pairs inside the group depend on m only: fold them into the phase table (into the
        // columns of a dense matrix; for a final layer behind a dense one, into its rows)
        uint32_t qg = 0;
        for (int m = 0; m < na; ++m) {
          uint32_t q = 0;
          for (int f = 0; f < r; ++f)
            for (int f2 = f + 1; f2 < r; ++f2)
              if (coupled[f][f2] && ((m >> (r - 1 - f)) & 1) && ((m >> (r - 1 - f2)) & 1)) q ^= 1u;
          qg |= q << m;
        }
        const bool behind_pair = is_fin && ws.nlayers > 0 && layers[ws.layers[ws.nlayers - 1]].kind == QS_LAYER_DENSE &&
                                 layers[ws.layers[ws.nlayers - 1]].op_idx[1] >= 0;
        if (paired || behind_pair) {
          // Pairs inside block A (factors 0,1) or inside block B (factors 2,3) go into the columns of
          // that 4x4 (rows, for the final layer behind it); pairs BETWEEN the blocks depend on both
          // block indices and travel as the 16-bit sign pattern `cross`.
          QsLayer& D = paired ? L : P.layers[nl - 2];
          const bool inA = coupled[0][1], inB = coupled[2][3];
          for (int blk = 0; blk < 2; ++blk) {
            if (!(blk == 0 ? inA : inB)) continue;
            double* m4 = P.coef + D.coef_off + 32 * blk;
            for (int row = 0; row < 4; ++row)
              for (int c = 0; c < 4; ++c) {
                const int idx = paired ? c : row;              // column (before) / row (after) index 3 = both bits set
                if (idx != 3) continue;
                m4[2 * (row * 4 + c)] *= -1.0;
                m4[2 * (row * 4 + c) + 1] *= -1.0;
              }
          }
          uint32_t cross = 0;
          for (int m = 0; m < na; ++m) {
            uint32_t q = 0;
            for (int f = 0; f < 2; ++f)
              for (int f2 = 2; f2 < 4; ++f2)
                if (coupled[f][f2] && ((m >> (r - 1 - f)) & 1) && ((m >> (r - 1 - f2)) & 1)) q ^= 1u;
            cross |= q << m;
          }
          L.cross = (uint16_t)cross;
        } else if (wl.kind == QS_LAYER_DENSE) {
          for (int row = 0; row < na; ++row)
            for (int c = 0; c < na; ++c)
              if (qg >> c & 1) {
                P.coef[L.coef_off + 2 * (row * na + c)] *= -1.0;
                P.coef[L.coef_off + 2 * (row * na + c) + 1] *= -1.0;
              }
        } else if (is_fin && ws.nlayers > 0 && layers[ws.layers[ws.nlayers - 1]].kind == QS_LAYER_DENSE) {
          const QsLayer& D = P.layers[nl - 2];
          for (int row = 0; row < na; ++row)
            if (qg >> row & 1)
              for (int c = 0; c < na; ++c) {
                P.coef[D.coef_off + 2 * (row * na + c)] *= -1.0;
                P.coef[D.coef_off + 2 * (row * na + c) + 1] *= -1.0;
              }
        } else if (wl.has_ph || qg) {
          L.flags |= QS_LF_PHASE;
          L.ph_off = (uint16_t)ncoef;
          for (int m = 0; m < na; ++m) {
            cplx ph((qg >> m & 1) ? -1.0 : 1.0, 0.0);
            for (int f = 0; f < r; ++f) ph *= wl.ph[f][(m >> (r - 1 - f)) & 1];
            P.coef[ncoef++] = ph.real();
            P.coef[ncoef++] = ph.imag();
          }
        }
      }
      st.nlayers = (uint8_t)count;
    }
    for (int l = 0; l < nl; ++l) {               // the head words the kernel branches on
      QsLayer& L = P.layers[l];
      uint32_t h = 0;
      if (L.flags & QS_LF_SIGN) h |= QS_LH_SIGN;
      if (L.flags & QS_LF_PHASE) h |= QS_LH_PHASE;
      if (L.flags & QS_LF_FINAL) h |= QS_LH_FINAL;
      if (L.kind == QS_LAYER_GENERAL) h |= QS_LH_GENERAL;
      if (L.kind == QS_LAYER_DENSE) h |= QS_LH_DENSE;
      if (L.kind == QS_LAYER_DENSE && P.steps[L.step].r == 4 && pair_of[l]) h |= QS_LH_PAIR;
      for (int f = 0; f < QS_MAX_R; ++f) {
        if (L.kind == QS_LAYER_ROT && L.form[f] == QS_FORM_TAN) h |= QS_LH_TAN(f);
        if (L.kind == QS_LAYER_ROT && L.form[f] == QS_FORM_SHEAR3) h |= QS_LH_SHEAR3(f) | QS_LH_SHEAR3ANY;
        if (L.kind == QS_LAYER_GENERAL && L.form[f] == QS_FORM_FULL) h |= QS_LH_FULL(f);
      }
      L.head = h;
    }
    P.nlayers = (uint32_t)nl;
    P.ncoef = (uint32_t)ncoef;
    P.npairs = (uint32_t)npairs;
  }
};

// Order the free local positions of a step: the three fastest thread bits go to
// positions below 6 that differ mod 3 (conflict-free under the 128-byte swizzle,
// tile_exec.h), lanes 3-4 next, then the `nwarp` warp-owned positions (thread-id bits
// 5..7), then the per-thread iteration bits.
void order_free_positions(QsStep& st, int T, const int* warp_pos, int nwarp) {
  bool used[QS_MAX_T + 1] = {false};
  for (int f = 0; f < st.r; ++f) used[st.gpos[f]] = true;
  for (int w = 0; w < nwarp; ++w) used[warp_pos[w]] = true;
  std::vector<int> freep;
  for (int p = 0; p < T; ++p)
    if (!used[p]) freep.push_back(p);
  std::vector<int> order;
  bool have[3] = {false, false, false};
  for (int p : freep)
    if (p < 6 && !have[p % 3]) { have[p % 3] = true; order.push_back(p); }
  for (int p : freep)
    if (std::find(order.begin(), order.end(), p) == order.end()) order.push_back(p);
  if (nwarp > 0) order.insert(order.begin() + 5, warp_pos, warp_pos + nwarp);   // needs >= 5 lane positions
  for (size_t i = 0; i < order.size() && i < QS_MAX_T; ++i) st.fpos[i] = (uint8_t)order[i];
}

// Split the steps of a pass into runs inside which every warp keeps working on
// the same 2^(T-3) amplitudes: three tile positions that no step of the run uses
// as a group bit are tied to the warp number, so steps of a run are separated by
// __syncwarp() only and the eight warps of a CTA drift apart -- shared-memory
// traffic of one warp then overlaps the FP64 work of another.
void assign_runs(QsPass& P) {
  const int T = (int)P.T;
  const int nsteps = (int)P.nsteps;
  const int thr_log2 = (int)P.cta_log2;
  const int warp_bits = thr_log2 - 5;            // thread-id bits that select the warp
  auto group_mask = [&](int i) {
    uint32_t g = 0;
    for (int f = 0; f < P.steps[i].r; ++f) g |= 1u << P.steps[i].gpos[f];
    return g;
  };
  // is step i left with three lane positions below 6 with different residues mod 3?
  auto lanes_ok = [&](int i, uint32_t wmask) {
    const uint32_t g = group_mask(i) | wmask;
    int have = 0;
    for (int p = 0; p < T && p < 6; ++p)
      if (!(g >> p & 1)) have |= 1 << (p % 3);
    return have == 7;
  };
  // Best warp-owned positions for the run [s, e): QS_WARP_BITS positions that no step of the
  // run uses as a group bit, chosen to leave as many steps as possible conflict-free
  // (highest positions first among equals).  Returns the number of conflict-free steps, -1
  // if there are not enough unused positions.
  auto best_warp_positions = [&](int s, int e, uint32_t used, int* wp) {
    int cand[QS_MAX_T], nc = 0;
    for (int p = T - 1; p >= 0; --p)
      if (!(used >> p & 1)) cand[nc++] = p;
    if (nc < warp_bits) return -1;
    int best = -1;
    // every choice of warp_bits (2 or 3) candidates
    for (int a = 0; a < nc; ++a)
      for (int b = a + 1; b < nc; ++b)
        for (int c = (warp_bits == 3 ? b + 1 : nc - 1); c < nc; ++c) {
          uint32_t wmask = (1u << cand[a]) | (1u << cand[b]);
          if (warp_bits == 3) wmask |= 1u << cand[c];
          int ok = 0;
          for (int i = s; i < e; ++i) ok += lanes_ok(i, wmask) ? 1 : 0;
          if (ok > best) {
            best = ok; wp[0] = cand[a]; wp[1] = cand[b];
            if (warp_bits == 3) wp[2] = cand[c];
          }
        }
    return best;
  };
  int s = 0;
  while (s < nsteps) {
    // Grow the run while warp-owned positions exist that cost no step its conflict-free
    // thread map: a bank conflict doubles a step's shared-memory time, a block-level
    // barrier instead of a warp-level one costs far less.
    uint32_t used = 0;
    int e = s, maxr = 0, can_ok = 0;
    int wp[4] = {0, 0, 0, 0};
    while (e < nsteps) {
      const uint32_t u = used | group_mask(e);
      const int r = std::max(maxr, (int)P.steps[e].r);
      if (T - __builtin_popcount(u) < warp_bits || T - r < thr_log2) break;
      int trial[4];
      const int ok = best_warp_positions(s, e + 1, u, trial);
      const int want = can_ok + (lanes_ok(e, 0) ? 1 : 0);
      if (ok < 0 || (e > s && ok < want)) break;
      used = u;
      maxr = r;
      can_ok = want;
      for (int i = 0; i < warp_bits; ++i) wp[i] = trial[i];
      ++e;
    }
    if (e == s) {                       // no room for warp-owned positions: plain block-synchronised step
      order_free_positions(P.steps[s], T, nullptr, 0);
      P.steps[s].block_sync = 1;
      ++s;
      continue;
    }
    uint32_t wp_mask = 0;
    for (int i = 0; i < warp_bits; ++i) wp_mask |= 1u << wp[i];
    if (e == s + 1 && !lanes_ok(s, wp_mask) && lanes_ok(s, 0)) {
      // a run of one step gains nothing from warp-owned positions
      order_free_positions(P.steps[s], T, nullptr, 0);
      P.steps[s].block_sync = 1;
      ++s;
      continue;
    }
    for (int i = s; i < e; ++i) {
      order_free_positions(P.steps[i], T, wp, warp_bits);
      P.steps[i].block_sync = (i == e - 1) ? 1 : 0;
    }
    s = e;
  }
}

// Thread-independent tables of the steps (they depend on the thread maps chosen by
// assign_runs): where the amplitudes of a work item and the per-thread iterations sit.
void finish_tables(QsPass& P) {
  const int T = (int)P.T;
  const int thr_log2 = (int)P.cta_log2;
  for (uint32_t s = 0; s < P.nsteps; ++s) {
    QsStep& st = P.steps[s];
    const int r = st.r;
    for (int m = 0; m < (1 << QS_MAX_R); ++m) {
      uint32_t d = 0;
      if (m < (1 << r))
        for (int f = 0; f < r; ++f) d |= (uint32_t)((m >> (r - 1 - f)) & 1) << st.gpos[f];
      st.sdepb[m] = qs_swz(d) << 4;
    }
    const int nfree = T - r;
    const int lo_bits = std::min(nfree, thr_log2);
    for (int i = 0; i < QS_MAX_WORK; ++i) {
      const uint32_t jhi =
          i < (1 << (nfree - lo_bits)) ? qs_scatter8((uint32_t)i, st.fpos + thr_log2, nfree - lo_bits) : 0u;
      st.hi[i] = jhi | (qs_swz(jhi) << 16);
    }
  }
  if (P.has_final) {
    const QsStep& st = P.steps[P.nsteps - 1];
    uint32_t qhi = 0;
    for (int i = 0; i < QS_MAX_WORK; ++i) {
      const uint32_t jhi = st.hi[i] & 0xffffu;
      qhi |= qs_fin_quad(P, jhi) << i;
      P.fin_neigh[i] = (uint16_t)qs_fin_neigh(P, jhi);
    }
    P.fin_qhi = (uint16_t)qhi;
  }
}

}  // namespace

int build_plan(int n, const std::vector<Op>& ops, const qsim_plan_options_t& opt, qsim_plan* out) {
  out->n = n;
  out->items.clear();
  qsim_plan_stats_t& stats = out->stats;
  stats.n_merged_ops = (int64_t)ops.size();
  std::vector<char> done(ops.size(), 0);
  std::vector<uint64_t> masks(ops.size());
  for (size_t i = 0; i < ops.size(); ++i) masks[i] = ops[i].mask();
  Walker wk{n, ops, masks, done, opt};
  const int T = std::min(opt.tile_bits, n);
  const int L = std::min(opt.low_bits, T);

  // first pending dense op per bit (tie-break when growing the tile)
  size_t first = 0;
  while (true) {
    while (first < ops.size() && done[first]) ++first;
    if (first >= ops.size()) break;

    const Op& head = ops[first];
    if (head.kind == OP_DENSE && head.k > QS_MAX_R) {
      PlanItem it;
      it.generic = true;
      it.op = head;
      out->items.push_back(std::move(it));
      done[first] = 1;
      stats.n_generic++;
      continue;
    }

    uint64_t S = 0;
    if (n <= T) {
      S = (n >= 64) ? ~0ull : ((1ull << n) - 1ull);
    } else {
      S = (1ull << L) - 1ull;
      if (head.kind == OP_DENSE) S |= head.mask();
      int size = __builtin_popcountll(S);
      // distance (in pending ops) to the next dense gate on each bit: tie-break
      std::vector<int> next_dense(n, 1 << 30);
      {
        int seen = 0;
        for (size_t i = first; i < ops.size() && seen < opt.lookahead; ++i) {
          if (done[i]) continue;
          ++seen;
          if (ops[i].kind != OP_DENSE) continue;
          for (int b : ops[i].bits)
            if (next_dense[b] == (1 << 30)) next_dense[b] = seen;
        }
      }
      while (size < T) {
        long best_score = -1;
        int best_bit = -1;
        for (int b = 0; b < n; ++b) {
          if (S >> b & 1) continue;
          const long sc = wk.walk(first, S | (1ull << b), nullptr, nullptr);
          if (sc > best_score ||
              (sc == best_score && best_bit >= 0 && next_dense[b] < next_dense[best_bit])) {
            best_score = sc;
            best_bit = b;
          }
        }
        S |= 1ull << best_bit;
        ++size;
      }
    }

    PlanItem it;
    it.generic = false;
    QsPass& P = it.pass;
    memset(&P, 0, sizeof(P));
    P.T = (uint32_t)__builtin_popcountll(S);
    {
      int l = 0;
      for (int b = 0; b < n; ++b)
        if (S >> b & 1) P.tile_bits[l++] = (uint8_t)b;
    }
    std::vector<size_t> taken;
    wk.walk(first, S, &P, &taken);
    if (taken.empty())
      return fail(QSIM_ERR_UNSUPPORTED, "planner made no progress (tile too small for the next gate)");
    for (size_t idx : taken) done[idx] = 1;
    {
      // Threads per CTA.  A pass with 16-amplitude steps needs ~120 registers per thread: 256-thread
      // CTAs then fit two to an SM, 128-thread CTAs three (the 64 KiB tiles allow no more), and three
      // independent load / compute / store phases on an SM beat two although the warps are fewer
      // (measured).  Passes with 8-amplitude steps fit three 256-thread CTAs and
      // keep them.  (Every thread must find its work items in st.hi[]: 2^(T-1) / threads <= QS_MAX_WORK.)
      int maxr = 0;
      for (uint32_t st = 0; st < P.nsteps; ++st) maxr = std::max(maxr, (int)P.steps[st].r);
      int cta = opt.cta_log2;
      if (cta == 0) cta = maxr >= 4 ? QS_THREADS_LOG2_MIN : QS_THREADS_LOG2;
      if ((int)P.T - 1 - cta > 4) cta = QS_THREADS_LOG2;
      P.cta_log2 = (uint8_t)cta;
    }
    assign_runs(P);
    finish_tables(P);
    for (uint32_t s = 0; s < P.nsteps; ++s) {
      stats.n_steps++;
      stats.n_warp_syncs += P.steps[s].block_sync ? 0 : 1;
    }
    stats.n_layers += P.nlayers;
    for (size_t idx : taken) {
      stats.n_sign += ops[idx].kind == OP_SIGN ? 1 : 0;
      stats.n_dense += ops[idx].kind == OP_DENSE ? 1 : 0;
    }
    stats.n_passes++;
    out->items.push_back(std::move(it));
  }
  return QSIM_OK;
}

}  // namespace qs
