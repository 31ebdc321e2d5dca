"""State vectors larger than one GPU: shard by the top qubits, swap on demand.

The reference has no distributed path at all; this module
is what ``BASELINE.json`` configuration C5 (34 qubits over 8 B200s) needs.

Layout.  A register of ``n`` qubits over ``P = 2^g`` ranks: rank ``r`` holds the
``2^(n-g)`` amplitudes whose top ``g`` index bits equal ``r``.  The *physical*
index bits ``0 .. n-g-1`` are local, bits ``n-g .. n-1`` are the rank number.
A table maps every logical index bit (reference qubit ``q`` is logical bit
``n-1-q``) to the physical bit it currently occupies.

Execution.  Gates are consumed in program order:

* a gate whose non-diagonal action touches only local bits joins the current
  local segment, which is run by the single-GPU fused planner on the shard;
* diagonal gates (Z, RZ, P, T, CZ, ...) never need communication: for the bits
  that live in the rank number the diagonal is restricted to this rank's values
  and becomes a smaller diagonal gate (or a scalar) on the shard;
* a non-diagonal gate on a rank bit triggers a *global<->local swap*: that bit
  changes places with the local bit whose next non-diagonal use lies farthest in
  the future.  Every rank keeps the half of its shard that already agrees with
  its rank bit and exchanges the other half with the partner rank
  ``r ^ 2^(bit)`` -- one send and one receive of half a shard per rank, all
  ``P/2`` pairs at once over NVLink/NVSwitch (``torch.distributed`` P2P on
  NCCL).  When the local bit is the top local bit the halves are contiguous and
  travel without staging; otherwise ``qsim_swap_pack`` / ``qsim_swap_unpack``
  gather and scatter them.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi, engine


def _is_diagonal(m: np.ndarray) -> bool:
    return not np.any(m - np.diag(np.diagonal(m)))


class Comm:
    """Thin wrapper over ``torch.distributed`` point-to-point exchange of
    complex128 buffers (moved as float64 pairs so gloo works too)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.size = dist.get_world_size(group)
        if self.size & (self.size - 1):
            raise ValueError("the number of ranks must be a power of two")
        self.bytes_exchanged = 0

    def exchange(self, send_t, recv_t, partner: int) -> None:
        """Send ``send_t`` to and receive ``recv_t`` from ``partner`` (torch tensors)."""
        import torch
        dist = self.dist
        s = torch.view_as_real(send_t) if send_t.is_complex() else send_t
        r = torch.view_as_real(recv_t) if recv_t.is_complex() else recv_t
        ops = [dist.P2POp(dist.isend, s, partner, self.group), dist.P2POp(dist.irecv, r, partner, self.group)]
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        self.bytes_exchanged += s.numel() * s.element_size()

    def allreduce_sum(self, values: np.ndarray) -> np.ndarray:
        import torch
        t = torch.from_numpy(np.ascontiguousarray(values, dtype=np.float64))
        dev = getattr(self, "device", None)
        if dev is not None:
            t = t.to(dev)
        self.dist.all_reduce(t, group=self.group)
        return t.cpu().numpy()


class ShardedState:
    """One rank's shard plus the logical->physical bit table (same on all ranks)."""

    def __init__(self, n: int, comm: Comm, backend=None, as_torch=None):
        self.n = int(n)
        self.comm = comm
        self.g = comm.size.bit_length() - 1
        if self.g >= self.n:
            raise ValueError("more ranks than amplitudes")
        self.n_local = self.n - self.g
        self.backend = backend or engine.get_backend()
        self.phys = list(range(self.n))           # phys[logical bit] = physical bit
        self.buf = self.backend.empty(1 << self.n_local)
        self.swaps = 0
        self.swap_seconds = 0.0
        self._peer = None
        # how to view a backend buffer as a torch tensor for the communicator
        self._as_torch = as_torch or (lambda b: b)

    # -- initial states ---------------------------------------------------------------------
    def set_product(self, vectors) -> None:
        """|psi> = kron of single-qubit kets (qubit 0 first).  The rank bits select
        one amplitude of each of the first g kets: a scalar on this rank."""
        vecs = [np.asarray(v, dtype=np.complex128) for v in vectors]
        assert len(vecs) == self.n
        self.phys = list(range(self.n))
        scale = 1.0 + 0.0j
        for q in range(self.g):                    # reference qubit q = logical bit n-1-q = rank bit
            bit = (self.comm.rank >> (self.g - 1 - q)) & 1
            scale *= vecs[q][bit]
        local = [v.copy() for v in vecs[self.g:]]
        local[0] = local[0] * scale
        amps = np.ascontiguousarray(np.stack(local))
        be = self.backend
        _capi.check(be.lib, be.lib.qsim_init_product(
            be.ptr(self.buf), self.n_local, amps.view(np.float64).ctypes.data_as(_capi.c_double_p), be.stream()))

    # -- reductions ---------------------------------------------------------------------------
    def norm(self) -> float:
        be = self.backend
        out = np.zeros(2)
        _capi.check(be.lib, be.lib.qsim_reduce_norm2(be.ptr(self.buf), C.c_uint64(1 << self.n_local),
                                                     out.ctypes.data_as(_capi.c_double_p), be.stream()))
        return float(np.sqrt(self.comm.allreduce_sum(out[:1])[0]))

    # -- swaps -----------------------------------------------------------------------------------
    CHUNK_LOG2 = 26            # amplitudes per pipeline chunk (1 GiB)

    def _peer_setup(self, chunk: int):
        """Allocate the four staging chunks with the library (plain cudaMalloc, so they can be
        exported through CUDA IPC) and open every other rank's receive chunks once."""
        import os
        be, lib = self.backend, self.backend.lib
        if self._peer is not None and self._peer["chunk"] == chunk:
            return self._peer
        use_ipc = getattr(be, "name", "") == "cuda" and os.environ.get("QSIM_SWAP_IPC", "1") != "0"
        peer = {"chunk": chunk, "ipc": use_ipc}
        if not use_ipc:
            peer["send"] = [be.empty(chunk) for _ in range(2)]
            peer["recv"] = [be.empty(chunk) for _ in range(2)]
            self._peer = peer
            return peer
        import torch
        dev = be.device.index
        ptrs = []
        for _ in range(4):
            out = C.c_void_p()
            _capi.check(lib, lib.qsim_peer_alloc(dev, C.c_uint64(16 * chunk), C.byref(out)))
            ptrs.append(out.value)
        peer["send_ptr"], peer["recv_ptr"] = ptrs[:2], ptrs[2:]
        handles = []
        for p in peer["recv_ptr"]:
            h = C.create_string_buffer(64)
            _capi.check(lib, lib.qsim_ipc_export(C.c_void_p(p), h))
            handles.append(h.raw)
        gathered = [None] * self.comm.size
        self.comm.dist.all_gather_object(gathered, handles, group=self.comm.group)
        peer["remote_recv"] = {}
        for r, hs in enumerate(gathered):
            if r == self.comm.rank:
                continue
            opened = []
            for h in hs:
                out = C.c_void_p()
                _capi.check(lib, lib.qsim_ipc_import(dev, h, C.byref(out)))
                opened.append(out.value)
            peer["remote_recv"][r] = opened
        peer["token_out"] = torch.zeros(2, dtype=torch.float64, device=be.device)
        peer["token_in"] = torch.zeros(2, dtype=torch.float64, device=be.device)
        self._peer = peer
        return peer

    def swap(self, global_phys: int, local_phys: int) -> None:
        """Exchange physical rank bit ``global_phys`` with physical local bit ``local_phys``.

        The travelling half shard is cut into chunks that flow through a three-stage
        pipeline on three streams: gather (``qsim_swap_pack``) into a staging chunk,
        transfer, scatter (``qsim_swap_unpack``).  On GPUs the transfer is a copy-engine
        PULL of the partner's staging chunk over NVLink (CUDA IPC + ``qsim_peer_copy``),
        ordered between the two processes by a stream-ordered NCCL token exchange, so
        no SM is spent on communication; elsewhere (gloo tests) it is a send/recv."""
        import time
        be, lib = self.backend, self.backend.lib
        gi = global_phys - self.n_local            # bit of the rank number
        keep = (self.comm.rank >> gi) & 1
        partner = self.comm.rank ^ (1 << gi)
        half = 1 << (self.n_local - 1)
        qubit = self.n_local - 1 - local_phys      # local reference-style qubit number
        chunk = min(half, 1 << self.CHUNK_LOG2)
        nchunks = half // chunk
        peer = self._peer_setup(chunk)
        ipc = peer["ipc"]
        be.synchronize()                           # so the timer below sees the swap alone
        t0 = time.perf_counter()
        if ipc:
            send_ptr, recv_ptr = peer["send_ptr"], peer["recv_ptr"]
            remote = peer["remote_recv"][partner]
        else:
            send_ptr = [be.ptr(b) for b in peer["send"]]
            recv_ptr = [be.ptr(b) for b in peer["recv"]]
        pipe = be.pipeline(3)                      # stream contexts + events (no-ops on the host emulator)
        packed, ordered, moved, unpacked = {}, {}, {}, {}

        def token():                               # both ranks have reached this point of their comm streams
            self.comm.exchange(peer["token_out"], peer["token_in"], partner)

        def pack(c):
            with pipe.stage(0):
                if c >= 2:
                    pipe.wait(moved[c - 2])        # send[c%2] has left (our own push / send of chunk c-2)
                _capi.check(lib, lib.qsim_swap_pack(be.ptr(self.buf), C.c_void_p(send_ptr[c % 2]), self.n_local,
                                                    qubit, keep, C.c_uint64(c * chunk), C.c_uint64(chunk),
                                                    be.stream()))
                packed[c] = pipe.record()

        def unpack(c):
            with pipe.stage(2):
                # IPC: the partner's push of chunk c precedes its token c+1 on its comm stream
                pipe.wait(ordered[c + 1] if ipc else moved[c])
                _capi.check(lib, lib.qsim_swap_unpack(be.ptr(self.buf), C.c_void_p(recv_ptr[c % 2]), self.n_local,
                                                      qubit, keep, C.c_uint64(c * chunk), C.c_uint64(chunk),
                                                      be.stream()))
                unpacked[c] = pipe.record()

        for c in range(nchunks):
            pack(c)
            with pipe.stage(1):
                pipe.wait(packed[c])
                if c >= 2:
                    pipe.wait(unpacked[c - 2])     # our recv[c%2] is free again
                if ipc:
                    # token c: both chunks c are packed and both recv[c%2] are free -> PUSH ours
                    # into the partner's staging with the copy engine (writes are the fast
                    # direction of NVLink P2P)
                    token()
                    ordered[c] = pipe.record()
                    _capi.check(lib, lib.qsim_peer_copy(C.c_void_p(remote[c % 2]), C.c_void_p(send_ptr[c % 2]),
                                                        C.c_uint64(16 * chunk), be.stream()))
                    self.comm.bytes_exchanged += 16 * chunk
                else:
                    self.comm.exchange(self._as_torch(peer["send"][c % 2]), self._as_torch(peer["recv"][c % 2]),
                                       partner)
                moved[c] = pipe.record()
            if ipc:
                if c >= 1:
                    unpack(c - 1)
            else:
                unpack(c)
        if ipc:
            with pipe.stage(1):
                token()                            # every push has landed on both sides
                ordered[nchunks] = pipe.record()
            unpack(nchunks - 1)
        pipe.join()
        be.synchronize()
        self.swap_seconds += time.perf_counter() - t0
        self.swaps += 1
        # the two logical bits trade places
        la, lb = self.phys.index(global_phys), self.phys.index(local_phys)
        self.phys[la], self.phys[lb] = local_phys, global_phys

    # -- gathering (tests / small registers only) -----------------------------------------------------
    def gather_numpy(self) -> np.ndarray:
        """Full state in logical order on every rank."""
        import torch
        local = torch.view_as_real(self._as_torch(self.buf)).contiguous()
        parts = [torch.empty_like(local) for _ in range(self.comm.size)]
        self.comm.dist.all_gather(parts, local, group=self.comm.group)
        full = np.concatenate([torch.view_as_complex(p).cpu().numpy() for p in parts])   # physical order
        # physical index -> logical index: move physical bit phys[l] to logical bit l
        cube = full.reshape((2,) * self.n)                    # axis a <-> physical bit n-1-a
        axes = [self.n - 1 - self.phys[self.n - 1 - a] for a in range(self.n)]   # logical axis a takes that physical axis
        return np.ascontiguousarray(cube.transpose(axes)).reshape(-1)


class ShardedSimulator:
    """Runs a list of matrix gates (this package's ``Gate`` objects) on a ShardedState."""

    def __init__(self, circuit, state: ShardedState, plan_options=None):
        self.circuit = circuit
        self.state = state
        self.plan_options = plan_options
        self.stats = {"segments": 0, "passes": 0, "swaps": 0, "local_gates": 0}

    # ---- scheduling helpers ---------------------------------------------------------------------
    def _lower(self):
        """[(logical_bits (factor order), matrix, is_diag)]"""
        n = self.state.n
        out = []
        for gate in self.circuit:
            for targets, matrix in gate.lowered(n, False):
                m = np.asarray(matrix, dtype=np.complex128)
                out.append(([n - 1 - q for q in targets], m, _is_diagonal(m)))
        return out

    def _restrict_diagonal(self, bits_phys, matrix):
        """Diagonal gate with some targets in the rank number: keep this rank's entries."""
        st = self.state
        k = len(bits_phys)
        diag = np.diagonal(matrix).reshape((2,) * k)
        index, local_bits = [], []
        for f, p in enumerate(bits_phys):
            if p >= st.n_local:
                index.append((st.comm.rank >> (p - st.n_local)) & 1)
            else:
                index.append(slice(None))
                local_bits.append(p)
        sub = np.asarray(diag[tuple(index)]).reshape(-1)
        if not local_bits:                           # pure scalar: put it on local bit 0
            return [0], np.diag([sub[0], sub[0]])
        return local_bits, np.diag(sub)

    def compile(self):
        """Walk the circuit once, tracking where every logical bit lives, and build
        the schedule: fused local plans separated by global<->local swaps.  The
        schedule starts from the identity layout (``set_product``)."""
        st = self.state
        ops = self._lower()
        nloc = st.n_local
        phys = list(range(st.n))                      # simulated layout
        uses = {}
        for idx, (bits, _m, is_diag) in enumerate(ops):
            if not is_diag:
                for b in bits:
                    uses.setdefault(b, []).append(idx)
        cursor = {b: 0 for b in uses}

        def next_use(bit, now):
            lst = uses.get(bit)
            if not lst:
                return 1 << 60
            c = cursor[bit]
            while c < len(lst) and lst[c] < now:
                c += 1
            cursor[bit] = c
            return lst[c] if c < len(lst) else 1 << 60

        schedule = []

        def flush(segment):
            if not segment:
                return
            plan = engine.Plan(st.backend, nloc, segment, self.plan_options)
            schedule.append(("plan", plan))
            self.stats["segments"] += 1
            self.stats["passes"] += plan.stats["n_passes"]
            self.stats["local_gates"] += len(segment)

        saved = st.phys
        try:
            st.phys = phys                             # _restrict_diagonal reads the layout
            pending = list(range(len(ops)))
            while pending:
                # Classify what could run in the current layout.  A gate that cannot
                # (non-diagonal on a rank bit) blocks its qubits; gates behind it are held
                # back only if they do not commute with what is blocked (diagonal gates
                # commute with each other).
                ready, deferred = [], []
                blocked_full, blocked_diag = set(), set()
                for idx in pending:
                    bits, m, is_diag = ops[idx]
                    free = not (blocked_full & set(bits)) and (is_diag or not (blocked_diag & set(bits)))
                    if free and (is_diag or all(phys[b] < nloc for b in bits)):
                        ready.append(idx)
                    else:
                        deferred.append(idx)
                        (blocked_diag if is_diag else blocked_full).update(bits)

                def run_ops(indices):
                    segment = []
                    for idx in indices:
                        bits, m, is_diag = ops[idx]
                        p_bits = [phys[b] for b in bits]
                        if is_diag and any(p >= nloc for p in p_bits):
                            p_bits, m = self._restrict_diagonal(p_bits, m)
                        segment.append(([nloc - 1 - p for p in p_bits], m))
                    flush(segment)

                def run_ready():
                    run_ops(ready)
                    ready.clear()

                if not deferred:
                    run_ready()
                    break
                # Bring in the rank bits of the first blocked gate.  Ready gates are NOT run
                # at every swap: gates that touch neither swapped qubit commute with the swap,
                # so they keep accumulating into bigger, better-packed fused plans.  They must
                # run first only if the evicted qubit still has a non-diagonal gate among them.
                head = deferred[0]
                bits = ops[head][0]
                for b in bits:
                    if phys[b] < nloc:
                        continue
                    busy = {}
                    for idx in ready:
                        if not ops[idx][2]:
                            for q in ops[idx][0]:
                                busy[q] = busy.get(q, 0) + 1
                    local_logical = [l for l in range(st.n) if phys[l] < nloc and l not in bits]
                    victim = max(local_logical, key=lambda l: (-busy.get(l, 0), next_use(l, head), phys[l]))
                    if victim in busy:
                        run_ready()       # (running only the closure of the victim's gates was
                                          #  tried: many small, badly packed plans -- more passes)
                    schedule.append(("swap", phys[b], phys[victim]))
                    phys[b], phys[victim] = phys[victim], phys[b]
                    self.stats["swaps"] += 1
                pending = sorted(ready + deferred)
        finally:
            st.phys = saved
        self._schedule = schedule
        return schedule

    def run(self) -> ShardedState:
        """Execute the schedule on the state (which must be in the identity layout,
        e.g. right after ``set_product``)."""
        st = self.state
        if getattr(self, "_schedule", None) is None:
            self.compile()
        if st.phys != list(range(st.n)):
            raise ValueError("the schedule assumes the identity layout; call set_product first")
        for item in self._schedule:
            if item[0] == "plan":
                item[1].execute(st.buf)
            else:
                st.swap(item[1], item[2])
        return st
