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

#include <algorithm>
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
  if (o.max_group <= 0) o.max_group = 3;
  if (o.max_dense_ops <= 0) o.max_dense_ops = 20;
  if (o.lookahead <= 0) o.lookahead = 600;
  if (o.merge_1q <= 0) o.merge_1q = 1;
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
      o.c = c; o.s = s; o.r0 = cplx(1.0, 0.0); o.r1 = r1;
      o.mat = {cplx(c, 0.0), -s * r1, cplx(s, 0.0), c * r1};
      t.has_pending = true;
      t.pending = {l0, cplx(0.0, 0.0), cplx(0.0, 0.0), l1};
      t.pending_diag = true;
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

// shape of a 2x2 matrix: lets the kernel skip the multiplications by exact zeros
uint8_t mat_form(const Op& op) {
  if (op.rot) return QS_FORM_ROT;
  auto zero = [](const cplx& v) { return v.real() == 0.0 && v.imag() == 0.0; };
  if (zero(op.mat[1]) && zero(op.mat[2])) return QS_FORM_DIAG;
  if (zero(op.mat[0]) && zero(op.mat[3])) return QS_FORM_ANTIDIAG;
  return QS_FORM_GENERAL;
}

// 8-double coefficient slot of one member of a 1Q step
void write_1q_slot(double* dst, const Op& op) {
  if (op.rot) {
    dst[0] = op.c; dst[1] = op.s;
    dst[2] = op.r0.real(); dst[3] = op.r0.imag();
    dst[4] = op.r1.real(); dst[5] = op.r1.imag();
    dst[6] = dst[7] = 0.0;
  } else {
    for (int e = 0; e < 4; ++e) { dst[2 * e] = op.mat[e].real(); dst[2 * e + 1] = op.mat[e].imag(); }
  }
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
  long walk(size_t first, uint64_t S, QsPass* pass, std::vector<size_t>* taken_idx) const {
    const uint64_t all = (n >= 64) ? ~0ull : ((1ull << n) - 1ull);
    uint64_t blocked_full = 0, blocked_diag = 0;
    int dense_taken = 0, sign_taken = 0, visited = 0;

    int lpos[64];
    {
      int t = 0;
      for (int b = 0; b < n; ++b) lpos[b] = (S >> b & 1) ? t++ : -1;
    }

    int nsteps = 0, ncoef = 0, npairs = 0;
    // Earliest-fit packing of single-qubit gates into 1Q steps: a gate on bit c may
    // join any existing step after the last matrix on c and not before the step
    // whose sign block holds the last CZ/Z on c (it commutes with everything else).
    int step_r[QS_MAX_STEPS];                // members of a 1Q step; -1 = dense step (closed)
    int last_dense[64], last_sign[64];
    for (int b = 0; b < 64; ++b) { last_dense[b] = -1; last_sign[b] = 0; }
    // Sign pairs taken but not yet attached: each waits for the first later step
    // whose group contains one of its bits, else for the pass's final block.
    uint64_t pend_mask = 0;
    uint8_t pend[256][2];
    int npend = 0;

    // (local position, outer bit) pairs per step; flattened into QsPass::pairs at the end
    constexpr int kMaxLo = 24;
    uint8_t step_lo[QS_MAX_STEPS][kMaxLo][2];
    int step_nlo[QS_MAX_STEPS];
    int total_lo = 0;

    // Earliest existing step that a matrix on bit c may join, given the pending
    // pairs that touch c (they would be attached to that step's sign block, so the
    // step must come after every earlier matrix on their partner bits).
    auto earliest_step = [&](int c) {
      int e = std::max(last_dense[c] + 1, last_sign[c]);
      if (pend_mask >> c & 1)
        for (int p = 0; p < npend; ++p) {
          const int a = pend[p][0], b = pend[p][1];
          if (a == c && b != c) e = std::max(e, last_dense[b] + 1);
          else if (b == c && a != c) e = std::max(e, last_dense[a] + 1);
        }
      return e;
    };
    auto pending_lo_count = [&](int c) {
      int cnt = 0;
      if (pend_mask >> c & 1)
        for (int p = 0; p < npend; ++p) {
          const int a = pend[p][0], b = pend[p][1];
          if ((a == c) != (b == c) && lpos[a == c ? b : a] < 0) ++cnt;
        }
      return cnt;
    };
    // Move every pending pair that touches bit c (factor f of step sidx) into that
    // step's sign block.
    auto attach = [&](int c, int sidx, int f, QsStep* st) {
      if (!(pend_mask >> c & 1)) return;
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
        last_sign[c] = std::max(last_sign[c], sidx);
        last_sign[e] = std::max(last_sign[e], sidx);
        if (e != c && lpos[e] < 0) {
          step_lo[sidx][step_nlo[sidx]][0] = (uint8_t)lpos[c];
          step_lo[sidx][step_nlo[sidx]][1] = (uint8_t)e;
          step_nlo[sidx]++;
          ++total_lo;
        }
        if (!st) continue;
        st->has_sign = 1;
        if (e == c) {
          st->zconst ^= (uint16_t)(1u << lpos[c]);
        } else if (lpos[e] >= 0) {
          st->ng[f] ^= (uint16_t)(1u << lpos[e]);
          // partner inside the same group (multi-qubit dense step): keep ng symmetric
          for (int f2 = 0; f2 < st->r; ++f2)
            if (f2 != f && st->gpos[f2] == lpos[e]) st->ng[f2] ^= (uint16_t)(1u << lpos[c]);
        }
      }
      npend = kept;
      pend_mask = new_mask;
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
      } else {
        const bool in_tile = (m & ~S) == 0;
        const bool free_bits = (m & blocked_full) == 0 && (op.diag || (m & blocked_diag) == 0);
        bool take = in_tile && free_bits && op.k <= QS_MAX_R && dense_taken < opt.max_dense_ops;
        int join = -1;
        if (take) {
          if (op.k == 1) {
            const int c = op.bits[0];
            const int nlo = pending_lo_count(c);
            for (int sidx = earliest_step(c); sidx < nsteps; ++sidx)
              if (step_r[sidx] >= 1 && step_r[sidx] < opt.max_group && step_nlo[sidx] + nlo <= kMaxLo) {
                join = sidx;
                break;
              }
          }
          if (join < 0) {
            const int need = (op.k == 1) ? 8 * opt.max_group + 2 * (1 << opt.max_group)
                                         : 2 * (1 << op.k) * (1 << op.k);
            int nlo = 0;
            for (int b : op.bits) nlo += pending_lo_count(b);
            if (nsteps >= QS_MAX_STEPS || ncoef + need > QS_MAX_COEF || nlo > kMaxLo ||
                total_lo + nlo + npend > QS_MAX_PAIRS)
              take = false;
          }
        }
        if (take && join >= 0) {
          const int c = op.bits[0];
          const int f = step_r[join];
          QsStep* st = pass ? &pass->steps[join] : nullptr;
          if (st) {
            st->gpos[f] = (uint8_t)lpos[c];
            st->form[f] = mat_form(op);
            write_1q_slot(pass->coef + st->coef_off + 8 * f, op);
            st->r++;
          }
          attach(c, join, f, st);
          step_r[join]++;
          last_dense[c] = join;
        } else if (take) {
          QsStep* st = pass ? &pass->steps[nsteps] : nullptr;
          if (st) *st = QsStep{};
          step_nlo[nsteps] = 0;
          const int dim = 1 << op.k;
          if (st) {
            st->coef_off = (uint16_t)ncoef;
            st->kind = (op.k == 1) ? QS_STEP_1Q : QS_STEP_DENSE;
            st->r = (uint8_t)op.k;
            for (int f = 0; f < op.k; ++f) st->gpos[f] = (uint8_t)lpos[op.bits[f]];
            if (op.k == 1) st->form[0] = mat_form(op);
            double* dst = pass->coef + ncoef;
            if (op.k == 1) write_1q_slot(dst, op);
            else
              for (int e = 0; e < dim * dim; ++e) { dst[2 * e] = op.mat[e].real(); dst[2 * e + 1] = op.mat[e].imag(); }
          }
          for (int f = 0; f < op.k; ++f) attach(op.bits[f], nsteps, f, st);
          if (op.k == 1) {
            ncoef += 8 * opt.max_group + 2 * (1 << opt.max_group);   // later members + phase table
            step_r[nsteps] = 1;
          } else {
            ncoef += 2 * dim * dim;
            step_r[nsteps] = -1;
          }
          for (int b : op.bits) last_dense[b] = nsteps;
          ++nsteps;
        }
        if (take) {
          ++dense_taken;
          if (taken_idx) taken_idx->push_back(i);
        } else if (op.diag) {
          blocked_diag |= m;
        } else {
          blocked_full |= m;
        }
      }
      if ((blocked_full & all) == all) break;
    }

    if (pass) {
      // phase tables of the steps that hold rotation-form members
      for (int sidx = 0; sidx < nsteps; ++sidx) {
        QsStep& st = pass->steps[sidx];
        if (st.kind != QS_STEP_1Q) continue;
        bool any = false;
        for (int f = 0; f < st.r; ++f) any |= st.form[f] == QS_FORM_ROT;
        if (!any) continue;
        st.has_phase = 1;
        st.ph_off = (uint16_t)(st.coef_off + 8 * opt.max_group);
        for (int m = 0; m < (1 << st.r); ++m) {
          cplx ph(1.0, 0.0);
          for (int f = 0; f < st.r; ++f) {
            if (st.form[f] != QS_FORM_ROT) continue;
            const double* slot = pass->coef + st.coef_off + 8 * f;
            const int bit = (m >> (st.r - 1 - f)) & 1;
            ph *= cplx(slot[2 + 2 * bit], slot[3 + 2 * bit]);
          }
          pass->coef[st.ph_off + 2 * m] = ph.real();
          pass->coef[st.ph_off + 2 * m + 1] = ph.imag();
        }
      }
      // flatten the per-step (local, outer) pair lists
      for (int sidx = 0; sidx < nsteps; ++sidx) {
        QsStep& st = pass->steps[sidx];
        st.pair_off = (uint16_t)npairs;
        st.n_lo = (uint16_t)step_nlo[sidx];
        for (int q = 0; q < step_nlo[sidx]; ++q) {
          pass->pairs[2 * npairs] = step_lo[sidx][q][0];
          pass->pairs[2 * npairs + 1] = step_lo[sidx][q][1];
          ++npairs;
        }
      }
      // whatever is still pending goes into the final block
      uint8_t* dst = pass->pairs + 2 * npairs;
      int w = 0, n_oo = 0, n_lo = 0;
      for (int p = 0; p < npend; ++p) {
        const int a = pend[p][0], b = pend[p][1];
        if (lpos[a] < 0 && lpos[b] < 0) { dst[w++] = (uint8_t)a; dst[w++] = (uint8_t)b; ++n_oo; }
      }
      for (int p = 0; p < npend; ++p) {
        const int a = pend[p][0], b = pend[p][1];
        if ((lpos[a] >= 0) != (lpos[b] >= 0)) {
          const int in = lpos[a] >= 0 ? a : b, outb = lpos[a] >= 0 ? b : a;
          dst[w++] = (uint8_t)lpos[in]; dst[w++] = (uint8_t)outb; ++n_lo;
        }
      }
      for (int p = 0; p < npend; ++p) {
        const int a = pend[p][0], b = pend[p][1];
        if (lpos[a] >= 0 && lpos[b] >= 0) {
          if (a == b) {
            pass->fin_zconst ^= (uint16_t)(1u << lpos[a]);
          } else {
            pass->fin_nsym[lpos[a]] ^= (uint16_t)(1u << lpos[b]);
            pass->fin_nsym[lpos[b]] ^= (uint16_t)(1u << lpos[a]);
          }
        }
      }
      pass->fin_has_sign = npend > 0 ? 1 : 0;
      pass->fin_pair_off = (uint16_t)npairs;
      pass->fin_n_oo = (uint16_t)n_oo;
      pass->fin_n_lo = (uint16_t)n_lo;
      npairs += n_oo + n_lo;
      pass->nsteps = (uint32_t)nsteps;
      pass->ncoef = (uint32_t)ncoef;
      pass->npairs = (uint32_t)npairs;
    }
    return (long)dense_taken * 4096 + std::min(sign_taken, 4095);
  }
};

// Order the free local positions of a step: the three fastest thread bits go to
// positions that differ mod 3 (conflict-free with the XOR-fold swizzle), lanes 3-4
// next, then the `nwarp` warp-owned positions (thread-id bits 5..7), then the
// per-thread iteration bits.
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
    if (!have[p % 3]) { have[p % 3] = true; order.push_back(p); }
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
  int s = 0;
  while (s < nsteps) {
    uint32_t used = 0;
    int e = s, maxr = 0;
    while (e < nsteps) {
      uint32_t u = used;
      const QsStep& st = P.steps[e];
      for (int f = 0; f < st.r; ++f) u |= 1u << st.gpos[f];
      const int r = std::max(maxr, (int)st.r);
      if (T - __builtin_popcount(u) < QS_WARP_BITS || T - r < QS_THREADS_LOG2) break;
      used = u;
      maxr = r;
      ++e;
    }
    if (e == s) {                       // no room for warp-owned positions: plain block-synchronised step
      order_free_positions(P.steps[s], T, nullptr, 0);
      P.steps[s].block_sync = 1;
      ++s;
      continue;
    }
    // Warp-owned positions: any QS_WARP_BITS positions that no step of the run uses as a
    // group bit.  Prefer a choice that leaves every step three lane positions with different
    // residues mod 3 (conflict-free 128-bit accesses under the XOR-fold swizzle, tile_exec.h);
    // the highest positions otherwise.
    int wp[4], nw = 0;
    {
      int cand[QS_MAX_T], nc = 0;
      for (int p = T - 1; p >= 0; --p)
        if (!(used >> p & 1)) cand[nc++] = p;
      auto steps_ok = [&](uint32_t wmask) {          // steps of the run left conflict-free
        int ok = 0;
        for (int i = s; i < e; ++i) {
          uint32_t g = 0;
          for (int f = 0; f < P.steps[i].r; ++f) g |= 1u << P.steps[i].gpos[f];
          int have = 0;
          for (int p = 0; p < T; ++p)
            if (!((g | wmask) >> p & 1)) have |= 1 << (p % 3);
          ok += have == 7;
        }
        return ok;
      };
      bool found = false;
      if (QS_WARP_BITS == 3 && nc >= 3) {
        int best = -1;
        for (int a = 0; a < nc; ++a)
          for (int b = a + 1; b < nc; ++b)
            for (int c = b + 1; c < nc; ++c) {
              const int ok = steps_ok((1u << cand[a]) | (1u << cand[b]) | (1u << cand[c]));
              if (ok > best) {                      // candidates come highest positions first
                best = ok;
                wp[0] = cand[a]; wp[1] = cand[b]; wp[2] = cand[c];
                nw = 3;
                found = true;
              }
            }
      }
      if (!found) {
        nw = 0;
        for (int i = 0; i < nc && nw < QS_WARP_BITS; ++i) wp[nw++] = cand[i];
      }
    }
    for (int i = s; i < e; ++i) {
      order_free_positions(P.steps[i], T, wp, QS_WARP_BITS);
      P.steps[i].block_sync = (i == e - 1) ? 1 : 0;
    }
    s = e;
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
    assign_runs(P);
    for (uint32_t s = 0; s < P.nsteps; ++s) {
      QsStep& st = P.steps[s];
      stats.n_steps++;
      stats.n_warp_syncs += st.block_sync ? 0 : 1;
      stats.n_dense += (st.kind == QS_STEP_1Q) ? st.r : 1;
    }
    for (size_t idx : taken) stats.n_sign += ops[idx].kind == OP_SIGN ? 1 : 0;
    stats.n_passes++;
    out->items.push_back(std::move(it));
  }
  return QSIM_OK;
}

}  // namespace qs
