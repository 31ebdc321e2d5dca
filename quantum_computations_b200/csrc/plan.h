// Pass / step descriptors shared by the host planner, the sm_100a kernels and
// the host emulator used by the CPU tests.
//
// A *pass* streams the whole state through shared memory once (one HBM read +
// one HBM write per amplitude).  The state index is split into T "tile" bits
// and n-T "outer" bits; every CTA owns the 2^T amplitudes of one outer value,
// stages them in shared memory, applies the pass's *steps* there and writes
// them back in place.  The low tile bits are always index bits 0..L-1 so every
// global access touches runs of 2^L consecutive complex128 amplitudes.
//
// A *step* is one shared-memory round trip: every thread pulls 2^r amplitudes
// (spanning r "group" bits of the tile) into registers, optionally flips signs,
// applies either r independent 2x2 matrices or one dense 2^r x 2^r matrix, and
// stores back.
//
// Sign blocks.  CZ and Z gates multiply amplitude i by (-1)^q(i), q a quadratic
// form over GF(2) in the index bits -- no memory traffic, any qubits.  They
// commute with every matrix that does not touch their bits, so the planner
// DEFERS each pending pair to the first later step whose group contains one of
// its bits (or to the pass's final block, applied on the way back to HBM).  A
// step's block therefore only holds pairs that touch its group bits, and its sign
// for amplitude m of a work item at local index j0 is
//     parity(m & W(j0)) + qg(m),
//     W_f = z_f + parity(j0 & ng[f])        (f = group factor)
// with ng[f] the in-tile partners of group bit f, z_f a Z on it (a local Z, or a
// CZ with an outer bit that is 1 for this tile), and qg the pairs inside the
// group: a few popcounts per work item.  Only the final block evaluates a full
// quadratic form  g + z.j + Q(j)  (tile-uniform bit g from pairs of outer bits,
// linear mask z, symmetric neighbour masks nsym).
//
// Index-bit convention inside the library: bit b of the linear index.  The
// reference numbers qubits from the most significant end, so reference qubit q
// of an n-qubit register is bit n-1-q (DV/numpy_quantum.py:243-247).
#pragma once
#include <stdint.h>

#define QS_MAX_R        4      // group bits per step (dense 16x16 at most)
#define QS_MAX_STEPS    32
#define QS_MAX_PAIRS    640    // (outer,outer) and (local,outer) sign pairs per pass
#define QS_MAX_COEF     2816   // doubles of matrix coefficients per pass
#ifndef QS_THREADS_LOG2
#define QS_THREADS_LOG2 8      // threads per CTA = 256 (7 and 9 are experiment switches)
#endif
// largest tile: 2^13 amplitudes = 128 KiB; a step's per-thread iteration table has 16 entries
#define QS_MAX_T        (QS_THREADS_LOG2 + 5 < 13 ? QS_THREADS_LOG2 + 5 : 13)
#define QS_THREADS      (1 << QS_THREADS_LOG2)
#define QS_WARP_BITS    (QS_THREADS_LOG2 - 5)                 // thread-id bits that select the warp
#define QS_MAX_ITER     (1 << (QS_MAX_T - QS_THREADS_LOG2))   // amplitudes per thread in load/store

enum QsStepKind : uint8_t {
  QS_STEP_1Q    = 0,   // r independent 2x2 matrices, one per group bit
  QS_STEP_DENSE = 1    // one 2^r x 2^r matrix on the r group bits
};

enum QsMatForm : uint8_t {
  QS_FORM_GENERAL  = 0,
  QS_FORM_DIAG     = 1,   // off-diagonal entries exactly zero
  QS_FORM_ANTIDIAG = 2,   // diagonal entries exactly zero (X-like)
  // unitary written as (real rotation [[c,-s],[s,c]]) x (diagonal phases): the
  // phases of all factors of a step are applied as ONE table over the group bits
  // (QsStep::ph_off), the rotation costs half the flops of a complex 2x2.
  // Coefficient slot: c, s, then unused.
  QS_FORM_ROT      = 3
};

struct QsStep {
  uint8_t  kind;
  uint8_t  r;                    // number of group bits
  uint8_t  gpos[QS_MAX_R];       // local position of matrix factor f (f=0: most significant)
  uint8_t  fpos[QS_MAX_T];       // the T-r free local positions, in thread-scatter order
  uint8_t  has_sign;             // 1 if the step's sign block is not empty
  uint8_t  form[QS_MAX_R];       // QS_STEP_1Q: shape of each 2x2 (QsMatForm), saves flops
  uint8_t  has_phase;            // 1 if a phase table (2^r complex, indexed by m) precedes the matrices
  uint8_t  block_sync;           // 1: __syncthreads() after this step; 0: the next step only needs data
                                 //    of the same warp (same warp-owned index bits), __syncwarp() is enough
  uint16_t ph_off;               // its offset (in doubles) in QsPass::coef
  uint16_t coef_off;             // first coefficient (in doubles) in QsPass::coef
  // sign block (only pairs touching a group bit)
  uint16_t pair_off;             // first (local position, outer global bit) pair in QsPass::pairs
  uint16_t n_lo;
  uint16_t zconst;               // local positions (group bits) carrying a Z
  uint16_t ng[QS_MAX_R];         // in-tile CZ partners (local positions) of group factor f
};

struct QsPass {
  uint32_t T;                    // tile bits
  uint32_t nsteps;
  uint32_t ncoef;
  uint32_t npairs;
  uint8_t  tile_bits[16];        // ascending global bit numbers of the local positions
  // final sign block, applied while storing back to global memory
  uint8_t  fin_has_sign;
  uint8_t  pad0;
  uint16_t fin_pair_off;         // first pair: n_oo (outer, outer) then n_lo (local, outer)
  uint16_t fin_n_oo;
  uint16_t fin_n_lo;
  uint16_t fin_zconst;
  uint16_t fin_nsym[QS_MAX_T];   // fin_nsym[p]: local positions CZ-coupled to local position p
  QsStep   steps[QS_MAX_STEPS];
  uint8_t  pairs[QS_MAX_PAIRS * 2];
  double   coef[QS_MAX_COEF];
};

// CUDA kernel parameters are limited to 32764 bytes (CUDA >= 12.1, sm_70+).
static_assert(sizeof(QsPass) <= 32000, "QsPass must fit in the kernel parameter space");

// Per-step lookup tables, built once per kernel launch in shared memory (they do
// not depend on the tile): where the thread id and the per-thread iteration
// counter land inside the tile.
struct QsStepTab {
  uint16_t jA[16];               // local-index bits of thread-id nibble 0
  uint16_t jB[32];               // local-index bits of thread-id bits 4..8
  uint32_t hi[16];               // iteration i: jhi | swz(jhi) << 16
  uint32_t sdepb[16];            // BYTE offset (swizzled slot * 16) of amplitude m of a work item
  uint16_t ng[QS_MAX_R];         // copy of QsStep::ng
  uint16_t qg;                   // bit m: parity of the CZ pairs inside the group for amplitude m
  uint8_t  gpos[QS_MAX_R];       // copy of QsStep::gpos
  uint8_t  all_rot;              // QS_STEP_1Q whose members are all QS_FORM_ROT: branch-free fast path
  uint8_t  pad[3];
};

// Tables for the load/store phases and the final sign block.
struct QsIoTab {
  uint64_t gbyte[QS_MAX_ITER];   // BYTE offset in the state of the global-index bits of iteration i
  uint32_t sbyte[QS_MAX_ITER];   // BYTE offset in the tile of swz(i << QS_THREADS_LOG2)
  uint16_t fin_neigh[QS_MAX_ITER];  // XOR of fin_nsym over the bits of i << QS_THREADS_LOG2
  uint64_t fin_q;                // bit i: Q(i << QS_THREADS_LOG2)
  uint64_t base_tab[4][64];      // tile number -> global index of the tile, 6 bits at a time
};
