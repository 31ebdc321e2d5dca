"""Plain-data (JSON-able) gate descriptions shared by the golden-vector
generator and the tests.  ``to_spec`` works on gates of either the reference or
this package; ``from_spec`` instantiates them from a given gates/simulator
module pair."""
from __future__ import annotations

import numpy as np


def to_spec(gate, matrices: dict) -> dict:
    """``matrices`` collects explicit matrices (generic gates) under fresh keys."""
    name = type(gate).__name__
    if name == "ClassicalControl":
        return {"name": "ClassicalControl", "pos": list(gate._pos), "neg": list(gate._neg),
                "gate": to_spec(gate.gate, matrices)}
    spec = {"name": name, "indices": [int(i) for i in gate.indices]}
    if name == "RZ":
        spec["angle"] = float(gate.angle)
    elif name == "Insert":
        spec["state"] = gate.state.name
    elif name in ("M", "MZ", "MX"):
        spec.update(name="M", theta=float(gate.theta), phi=float(gate.phi),
                    result=None if gate.result is None else int(gate.result))
    elif name == "Channel":
        key = f"kraus{len(matrices)}"
        matrices[key] = np.stack([np.asarray(k, dtype=np.complex128) for k in gate.kraus])
        spec["kraus"] = key
    elif name == "Gate":
        key = f"mat{len(matrices)}"
        matrices[key] = np.asarray(gate.matrix)
        spec["matrix"] = key
    return spec


def from_spec(spec: dict, gates, simulator=None, matrices=None, channels=None, states=None):
    name = spec["name"]
    if name == "ClassicalControl":
        inner = from_spec(spec["gate"], gates, simulator, matrices, channels, states)
        return simulator.ClassicalControl(inner, list(spec["pos"]), list(spec["neg"]))
    idx = list(spec.get("indices", []))
    if name == "RZ":
        return gates.RZ(idx[0], spec["angle"])
    if name == "Insert":
        st = states.State if states is not None else gates.State
        return gates.Insert(idx[0], st[spec["state"]])
    if name == "M":
        return gates.M(idx[0], spec["theta"], spec["phi"], result=spec["result"])
    if name == "Channel":
        return channels.Channel(idx, list(matrices[spec["kraus"]]))
    if name == "Gate":
        return gates.Gate(idx, matrices[spec["matrix"]])
    return getattr(gates, name)(*idx)


def as_oracle_ops(circuit) -> list:
    """Translate gate objects (either package) to the oracle's plain tuples."""
    ops = []
    for gate in circuit:
        name = type(gate).__name__
        if name == "Insert":
            ops.append(("ins", gate.indices[0], np.asarray(gate.matrix)[0, :]))
        elif name in ("M", "MZ", "MX"):
            ops.append(("m", gate.indices[0], gate.theta, gate.phi, gate.result))
        elif name == "Channel":
            ops.append(("kraus", list(gate.indices), list(gate.kraus)))
        else:
            ops.append(("u", list(gate.indices), np.asarray(gate.matrix)))
    return ops
