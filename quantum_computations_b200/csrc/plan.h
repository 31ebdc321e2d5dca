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
// (spanning r "group" bits of the tile) into registers, optionally flips signs
// (a block of CZ/Z gates, which may involve ANY index bit), applies either r
// independent 2x2 matrices or one dense 2^r x 2^r matrix, and stores back.
//
// Sign blocks.  A block of CZ/Z gates multiplies amplitude i by (-1)^q(i) with q
// a quadratic form over GF(2) in the index bits.  Relative to a tile it splits
//   q = g  +  z . j  +  Q(j)           (j = local index inside the tile)
// into a tile-uniform bit g (pairs of outer bits), a linear part z (local Z
// gates: `zconst`; CZ between a local and an outer bit: the `lo` pairs, resolved
// per tile) and a quadratic part Q over local bits only, stored as symmetric
// neighbour masks `nsym`.  The kernels evaluate it with a handful of popcounts
// per work item instead of a loop over pairs per amplitude (tile_exec.h).
//
// Index-bit convention inside the library: bit b of the linear index.  The
// reference numbers qubits from the most significant end, so reference qubit q
// of an n-qubit register is bit n-1-q (DV/numpy_quantum.py:243-247).
#pragma once
#include <stdint.h>

#define QS_MAX_T        13     // largest tile: 2^13 amplitudes = 128 KiB
#define QS_MAX_R        4      // group bits per step (dense 16x16 at most)
#define QS_MAX_STEPS    56
#define QS_MAX_PAIRS    640    // (outer,outer) and (local,outer) sign pairs per pass
#define QS_MAX_COEF     2816   // doubles of matrix coefficients per pass
#define QS_THREADS_LOG2 8
#define QS_THREADS      (1 << QS_THREADS_LOG2)

enum QsStepKind : uint8_t {
  QS_STEP_1Q    = 0,   // r independent 2x2 matrices, one per group bit
  QS_STEP_DENSE = 1,   // one 2^r x 2^r matrix on the r group bits
  QS_STEP_SIGN  = 2    // sign block only (fused into the final store)
};

struct QsStep {
  uint8_t  kind;
  uint8_t  r;                    // number of group bits (0 for QS_STEP_SIGN)
  uint8_t  gpos[QS_MAX_R];       // local position of matrix factor f (f=0: most significant)
  uint8_t  fpos[QS_MAX_T];       // the T-r free local positions, in thread-scatter order
  uint8_t  has_sign;             // 1 if the step's sign block is not empty
  uint16_t coef_off;             // first coefficient (in doubles) in QsPass::coef
  // sign block applied to the amplitudes as they are loaded for this step
  uint16_t pair_off;             // first pair (index into QsPass::pairs, 2 bytes each)
  uint8_t  n_oo;                 // pairs with both bits outside the tile (global bit numbers)
  uint8_t  n_lo;                 // pairs (local position, outer global bit)
  uint16_t zconst;               // local positions carrying a Z
  uint16_t nsym[QS_MAX_T];       // nsym[p]: local positions CZ-coupled to local position p
};

struct QsPass {
  uint32_t T;                    // tile bits
  uint32_t nsteps;               // steps[nsteps-1] is always a QS_STEP_SIGN (maybe empty)
  uint32_t ncoef;
  uint32_t npairs;
  uint8_t  tile_bits[16];        // ascending global bit numbers of the local positions
  QsStep   steps[QS_MAX_STEPS];
  uint8_t  pairs[QS_MAX_PAIRS * 2];
  double   coef[QS_MAX_COEF];
};

// CUDA kernel parameters are limited to 32764 bytes (CUDA >= 12.1, sm_70+).
static_assert(sizeof(QsPass) <= 32000, "QsPass must fit in the kernel parameter space");
