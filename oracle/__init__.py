"""CPU oracle for the dv_simulator gate-application path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker or as the
timed CPU comparator.  The product path (``quantum_computations_b200``) never
imports this package and fails loudly when its CUDA library is missing.

Modules
-------
``dense_ref``  restates the reference algorithm itself: expand every gate to a
               dense 2^N x 2^N operator with Kronecker products, permute its
               tensor factors, multiply (O(4^N)); this is what the reference
               does and what the CPU baseline times.
``strided``    the same mathematics as strided tensor contractions (O(2^N)),
               used as the checker at sizes the dense form cannot reach.
``gkp_noise``  the analytic finite-squeezing logical-error model.

Parity pin: the reference ships no golden vectors or tests for this path
(SURVEY.md section 8c).  The oracle is pinned instead against outputs of the
reference itself, generated in the build container by
``tests/golden/make_golden.py`` (which imports /root/reference) and committed
under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks both oracle
modules against every stored vector.
"""
