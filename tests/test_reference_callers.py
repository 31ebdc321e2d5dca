"""Unmodified REFERENCE callers on top of this package (drop-in claim).

``compat.install()`` registers this package as ``simulators.dv_simulator``; the reference's own
scripts -- ``randomised_benchmarking.random_circ``, ``gkp_simulator.transpiler.MBGKPCircuit`` and
``dv_circuits.grover`` -- are then imported from /root/reference as they are and must build,
type-dispatch and simulate with this package's classes.  Runs in a subprocess (it rewires
``sys.modules``); skipped where the reference checkout does not exist (the GPU box)."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
PAPER = os.path.join(REF, "impact_of_finite_squeezing_on_near-term_quantum_computations_using_gkp_qubits")

SCRIPT = textwrap.dedent(r'''
    import pickle, sys
    import numpy as np
    sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
    sys.path.insert(0, {ref!r}); sys.path.insert(0, {paper!r})
    import quantum_computations_b200.compat as compat
    compat.install()
    from quantum_computations_b200 import gates, simulator, workloads
    from quantum_computations_b200.states import State
    from emu_backend import emu
    from golden.specs import as_oracle_ops
    from oracle import strided

    # 1. the reference's RB generator builds OUR gate objects, the same ones workloads.rb_random_circuit draws
    import randomised_benchmarking as rb                      # unmodified reference script
    assert rb.dv_gates is gates and rb.DVSimulator is simulator.Simulator
    for seed, depth in ((1, 8), (2, 15), (3, 20)):
        dv_circ, gkp_circ = rb.random_circ(2, depth, seed)
        assert all(type(g).__module__ == gates.__name__ for g in dv_circ)
        mine = workloads.rb_random_circuit(2, depth, np.random.default_rng(seed))
        assert [(type(g).__name__, list(g.indices)) for g in dv_circ] == \
               [(type(g).__name__, list(g.indices)) for g in mine]
        assert gkp_circ.depth() == depth

    # 2. the measurement-based transpiler dispatches on our classes (GKP/transpiler.py:41-63)
    from simulators.gkp_simulator.transpiler import MBGKPCircuit
    circ = MBGKPCircuit(3)
    for g in (gates.H(0), gates.CZ(0, 1), gates.P(2), gates.SWAP(1, 2), gates.Pdg(0), gates.I(1)):
        circ.add_gate(g)
    circ.fill()
    assert circ.depth() >= 3

    # 3. the reference's Grover circuit, simulated by this package (host emulator backend) == CPU oracle
    import dv_circuits as ccs
    for tagged in ([3, 6], [0, 4]):
        gcirc = ccs.grover(ccs.oracle(tagged))
        assert all(type(g).__module__ in (gates.__name__, simulator.__name__) for g in gcirc)
        out = simulator.Simulator(gcirc, backend=emu()).run(None)
        ref, _ = strided.run(as_oracle_ops(gcirc), np.ones(1, dtype=np.complex128))
        assert np.abs(out - ref).max() < 1e-12
        probs = np.abs(out) ** 2
        assert all(abs(probs[t] - 0.5) < 1e-12 for t in tagged)

    # 4. gate objects survive pickling (the reference farms samples out with multiprocessing)
    g2 = pickle.loads(pickle.dumps([gates.H(0), gates.RZ(1, 0.3), gates.CZ(0, 1)]))
    assert [type(g).__name__ for g in g2] == ["H", "RZ", "CZ"] and np.allclose(g2[1].matrix, gates.RZ(1, 0.3).matrix)
    print("REFERENCE-CALLERS-OK")
''')


@pytest.mark.skipif(not os.path.isdir(PAPER), reason="the reference checkout is not present on this machine")
def test_unmodified_reference_callers_run_on_this_package():
    out = subprocess.run([sys.executable, "-c", SCRIPT.format(root=ROOT, ref=REF, paper=PAPER)],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0 and "REFERENCE-CALLERS-OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
