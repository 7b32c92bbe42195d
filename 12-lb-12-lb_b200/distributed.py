"""Host layer for several GPUs of one node: one process per GPU, the lattice cut into y-slabs.

The data path has exactly two exchanges per time step:
  * halo: the 18 population rows (3 outgoing directions x 6 distributions) that leave the slab on
    each side travel to the periodic neighbours                       -- point-to-point send/recv
  * Poisson: the half spectrum of the local rows is transposed to "all rows of my spectral
    columns" and back                                                  -- two all-to-alls
plus one row of the potential per side for E = -grad(phi).  `SlabDriver` sequences these around the
library's per-slab kernels (plbm_step_local, plbm_halo_pack/unpack, plbm_poisson_stage) through
torch.distributed (NCCL on GPUs).  The kernels come from a backend object so that the sequencing,
the neighbour map and the split sizes can be exercised on CPU with the gloo backend and a test
double (tests/slab_double.py); the product backend is `CudaSlabBackend` below, a thin ctypes binding
with no arithmetic of its own.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .api import PlasmaLBM, PlbmError, _check


class PlbmExchange(C.Structure):
    _fields_ = [("nranks", C.c_int), ("rank", C.c_int),
                ("halo_send_lo", C.c_void_p), ("halo_send_hi", C.c_void_p), ("halo_recv_lo", C.c_void_p), ("halo_recv_hi", C.c_void_p),
                ("halo_count", C.c_longlong),
                ("phi_first_row", C.c_void_p), ("phi_last_row", C.c_void_p), ("phi_below", C.c_void_p), ("phi_above", C.c_void_p),
                ("phi_count", C.c_longlong),
                ("t1", C.c_void_p), ("t2", C.c_void_p),
                ("slab_y0", C.c_int * 17), ("slab_k0", C.c_int * 17)]


def slab_of(NY: int, rank: int, nranks: int):
    """The library's decomposition rule (plbm_slab_of): balanced, even-sized slabs."""
    from .api import load_library
    lib = load_library()
    lib.plbm_slab_of.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    y0, nyl = C.c_int(), C.c_int()
    _check(lib, lib.plbm_slab_of(NY, rank, nranks, C.byref(y0), C.byref(nyl)), "plbm_slab_of")
    return y0.value, nyl.value


class _DeviceMemory:
    """A raw device allocation exposed through __cuda_array_interface__ (zero-copy torch view)."""

    def __init__(self, ptr: int, count: int):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 2}


class CudaSlabBackend:
    """One slab on one GPU: the library context plus torch views of its exchange buffers."""

    def __init__(self, NX: int, NY: int, rank: int, nranks: int, poisson: str = "fft", device: int = 0, **si):
        import torch
        self.torch = torch
        self.sim = PlasmaLBM(NX, NY, poisson=poisson, rank=rank, nranks=nranks, device=device, **si)
        lib = self.sim.lib
        for name in ("plbm_step_local", "plbm_poisson_stage"):
            getattr(lib, name).argtypes = [C.c_void_p, C.c_int]
        for name in ("plbm_halo_pack", "plbm_halo_unpack"):
            getattr(lib, name).argtypes = [C.c_void_p]
        lib.plbm_exchange_info.argtypes = [C.c_void_p, C.POINTER(PlbmExchange)]
        lib.plbm_peer_export.argtypes = [C.c_void_p, C.c_void_p]
        lib.plbm_peer_attach.argtypes = [C.c_void_p, C.c_void_p]
        lib.plbm_peer_barrier.argtypes = [C.c_void_p]
        lib.plbm_peer_check.argtypes = [C.c_void_p]
        lib.plbm_peer_detach.argtypes = [C.c_void_p]
        lib.plbm_step_peer.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_float)]
        lib.plbm_halo_push.argtypes = [C.c_void_p]
        lib.plbm_phi_rows_push.argtypes = [C.c_void_p]
        self.lib = lib
        x = PlbmExchange()
        _check(lib, lib.plbm_exchange_info(self.sim._h, C.byref(x)), "plbm_exchange_info")
        self.rank, self.nranks = rank, nranks
        self.slab_y0 = list(x.slab_y0[: nranks + 1])
        self.slab_k0 = list(x.slab_k0[: nranks + 1])
        self.NX, self.NY = NX, NY
        self.has_poisson = poisson == "fft"
        dev = torch.device("cuda", device)
        view = lambda ptr, n: torch.as_tensor(_DeviceMemory(ptr, n), device=dev)
        self.halo_send_lo = view(x.halo_send_lo, x.halo_count); self.halo_send_hi = view(x.halo_send_hi, x.halo_count)
        self.halo_recv_lo = view(x.halo_recv_lo, x.halo_count); self.halo_recv_hi = view(x.halo_recv_hi, x.halo_count)
        self.phi_first_row = view(x.phi_first_row, x.phi_count); self.phi_last_row = view(x.phi_last_row, x.phi_count)
        self.phi_below = view(x.phi_below, x.phi_count); self.phi_above = view(x.phi_above, x.phi_count)
        if self.has_poisson:
            nh, nyl = NY // 2 + 1, self.slab_y0[rank + 1] - self.slab_y0[rank]
            nkl = self.slab_k0[rank + 1] - self.slab_k0[rank]
            self.t1 = view(x.t1, 2 * nh * nyl)
            self.t2 = view(x.t2, 2 * max(nkl, 1) * NX)[: 2 * nkl * NX]
        # NCCL work is ordered against the library's own stream
        self.stream = torch.cuda.ExternalStream(self.sim.stream, device=dev)

    def step_local(self, want_fields: bool):
        _check(self.lib, self.lib.plbm_step_local(self.sim._h, int(want_fields)), "plbm_step_local")

    def halo_pack(self):
        _check(self.lib, self.lib.plbm_halo_pack(self.sim._h), "plbm_halo_pack")

    def halo_unpack(self):
        _check(self.lib, self.lib.plbm_halo_unpack(self.sim._h), "plbm_halo_unpack")

    def poisson_stage(self, stage: int):
        _check(self.lib, self.lib.plbm_poisson_stage(self.sim._h, stage), "plbm_poisson_stage")

    # peer-memory transposes (plbm.h: plbm_peer_*); PEER_BLOB_BYTES = PLBM_PEER_BLOB_BYTES
    PEER_BLOB_BYTES = 512

    def peer_export(self) -> bytes:
        blob = C.create_string_buffer(self.PEER_BLOB_BYTES)
        _check(self.lib, self.lib.plbm_peer_export(self.sim._h, blob), "plbm_peer_export")
        return blob.raw

    def peer_attach(self, blobs: bytes):
        _check(self.lib, self.lib.plbm_peer_attach(self.sim._h, C.create_string_buffer(blobs, len(blobs))), "plbm_peer_attach")
        self._peers_attached = True

    def peer_barrier(self):
        _check(self.lib, self.lib.plbm_peer_barrier(self.sim._h), "plbm_peer_barrier")

    def halo_push(self):
        _check(self.lib, self.lib.plbm_halo_push(self.sim._h), "plbm_halo_push")

    # plbm.h: PLBM_PEER_STAGES; main stream: k1 .. phi_wait, Poisson stream: charge_pull .. barrier3
    STAGES = ("k1", "halo_push", "halo_barrier", "halo_unpack", "phi_wait", "charge_pull", "p1_rows_fwd", "barrier1", "p2_cols_peer",
              "barrier2", "p3_rows_inv_phi_rows", "barrier3")

    def step_peer(self, nsteps: int, want_fields: bool = False, stage_ms=None):
        """The whole peer-memory step sequence nsteps times in one library call (plbm_step_peer; by default with the Poisson
        solve of a step on a second stream beside its K1).  stage_ms: a dict that receives the accumulated device time [ms] of
        the stages (STAGES; at most the last 32 steps are timed)."""
        buf = (C.c_float * len(self.STAGES))() if stage_ms is not None else None
        _check(self.lib, self.lib.plbm_step_peer(self.sim._h, nsteps, int(want_fields), buf), "plbm_step_peer")
        if stage_ms is not None:
            for k, name in enumerate(self.STAGES):
                stage_ms[name] = stage_ms.get(name, 0.0) + float(buf[k])
            stage_ms["timed_steps"] = stage_ms.get("timed_steps", 0) + min(nsteps, 32)

    def phi_rows_push(self):
        _check(self.lib, self.lib.plbm_phi_rows_push(self.sim._h), "plbm_phi_rows_push")

    def peer_detach(self):
        _check(self.lib, self.lib.plbm_peer_detach(self.sim._h), "plbm_peer_detach")
        self._peers_attached = False

    def peer_check(self):
        _check(self.lib, self.lib.plbm_peer_check(self.sim._h), "plbm_peer_check")

    def stream_context(self):
        return self.torch.cuda.stream(self.stream)

    def fields(self, names=None):
        return self.sim.fields(names) if names else self.sim.fields()

    def sync(self):
        self.sim.sync()
        if getattr(self, "_peers_attached", False):
            self.peer_check()          # a barrier that gave up (dead peer) must not pass silently

    def close(self):
        self.sim.close()


class SlabDriver:
    """Sequences one time step over all slabs.  `backend` provides the per-slab kernels and the
    exchange buffers (torch tensors on the backend's device); `dist` is torch.distributed."""

    def __init__(self, backend, group=None, peer_memory=None):
        """peer_memory: True = require the peer-memory transposes, False = all-to-alls through torch.distributed,
        None = use peer memory when the backend offers it and every rank could attach (env PLBM_NO_PEER=1 disables)."""
        import os
        import torch.distributed as dist
        self.dist = dist
        self.b = backend
        self.group = group
        self.rank, self.nranks = backend.rank, backend.nranks
        if self.nranks < 2:
            raise PlbmError("SlabDriver needs at least two slabs; use PlasmaLBM for one GPU")
        self.up = (self.rank + 1) % self.nranks       # owns the rows above mine (periodic)
        self.down = (self.rank - 1) % self.nranks
        y0, k0 = backend.slab_y0, backend.slab_k0
        nyl = y0[self.rank + 1] - y0[self.rank]
        nkl = k0[self.rank + 1] - k0[self.rank]
        # all-to-all split sizes in doubles (complex = 2): T1 -> T2 and back
        self.t1_splits = [2 * (k0[d + 1] - k0[d]) * nyl for d in range(self.nranks)]
        self.t2_splits = [2 * nkl * (y0[s + 1] - y0[s]) for s in range(self.nranks)]
        self.peer = False
        if peer_memory is None:
            peer_memory = None if os.environ.get("PLBM_NO_PEER", "0") in ("", "0") else False
        if peer_memory is not False and backend.has_poisson and hasattr(backend, "peer_export"):
            self.peer = self._attach_peers(required=bool(peer_memory))

    def close(self):
        """Collective: unmap the peers' memory on every rank before any rank frees its own (call before backend.close())."""
        if self.peer:
            self.b.peer_check()
            self.b.peer_detach()
            self.peer = False
            self.dist.barrier(group=self.group)

    def _attach_peers(self, required: bool) -> bool:
        """Gather every rank's IPC blob and map the peers' spectra; all ranks agree on the outcome."""
        import torch
        b, d = self.b, self.dist
        dev = b.stream.device
        ok, blob, err = 1, bytes(b.PEER_BLOB_BYTES), ""
        try:
            blob = b.peer_export()
        except PlbmError as e:
            ok, err = 0, str(e)
        mine = torch.tensor(list(blob) + [ok], dtype=torch.uint8, device=dev)
        every = [torch.empty_like(mine) for _ in range(self.nranks)]
        d.all_gather(every, mine, group=self.group)
        every = [t.cpu().numpy().tobytes() for t in every]
        if all(t[-1] for t in every):
            try:
                b.peer_attach(b"".join(t[:-1] for t in every))
            except PlbmError as e:
                ok, err = 0, str(e)
        else:
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=dev)
        d.all_reduce(flag, op=d.ReduceOp.MIN, group=self.group)     # also orders the flag memset before any barrier
        if int(flag.item()) == 0:
            if required:
                raise PlbmError(f"peer-memory transposes were required but are unavailable on rank {self.rank}: {err or 'a peer failed'}")
            return False
        return True

    def _sendrecv(self, send_up, send_down, recv_from_down, recv_from_up, wait=True):
        """What I send up is what my upper neighbour receives from below, and vice versa.  With two
        slabs both neighbours are the same rank: the order send(up), send(down) / recv(down), recv(up)
        pairs the messages correctly because point-to-point traffic between two ranks is ordered.
        wait=False returns the pending work handles (the caller waits before it uses the buffers)."""
        d = self.dist
        ops = [d.P2POp(d.isend, send_up, self.up, self.group), d.P2POp(d.isend, send_down, self.down, self.group),
               d.P2POp(d.irecv, recv_from_down, self.down, self.group), d.P2POp(d.irecv, recv_from_up, self.up, self.group)]
        works = d.batch_isend_irecv(ops)
        if wait:
            for w in works:
                w.wait()
            return []
        return works

    def step(self, nsteps: int = 1, want_fields: bool = False, timing=None, stage_ms=None):
        """nsteps time steps.  timing: optional list that receives one (start, end) CUDA event pair
        around every fused collide-stream launch (CUDA backend only).  stage_ms: optional dict for the per-stage device
        times of the peer-memory path (CudaSlabBackend.step_peer).

        Order inside a step: the population halo exchange only feeds the NEXT step's K1, so it is
        started right after K1 and completed after the Poisson stages, which it overlaps."""
        b, d = self.b, self.dist
        if self.peer and timing is None and hasattr(b, "step_peer"):
            # no messages and no host round trip per kernel: the library issues the whole sequence itself
            b.step_peer(nsteps, want_fields, stage_ms)
            return
        with b.stream_context():
            for t in range(nsteps):
                if timing is not None:
                    e0, e1 = b.torch.cuda.Event(enable_timing=True), b.torch.cuda.Event(enable_timing=True)
                    e0.record()
                b.step_local(want_fields and t == nsteps - 1)
                if timing is not None:
                    e1.record()
                    timing.append((e0, e1))
                if self.peer:
                    # no messages at all: neighbours' buffers and every slab's half spectrum are written / read in
                    # place through peer memory (NVLink), ordered by three flag barriers on the library's stream
                    b.halo_push()
                    b.poisson_stage(0)
                    b.peer_barrier()
                    b.halo_unpack()
                    b.poisson_stage(4)
                    b.peer_barrier()
                    b.poisson_stage(5)        # P3, boundary rows of phi stored into the neighbours as well
                    b.peer_barrier()
                    b.poisson_stage(3)
                    continue
                # halo: top row's upward populations go up, bottom row's downward populations go down
                b.halo_pack()
                pending = self._sendrecv(b.halo_send_hi, b.halo_send_lo, b.halo_recv_lo, b.halo_recv_hi, wait=False)
                # spectral Poisson with two transposes
                b.poisson_stage(0)
                if b.has_poisson:
                    d.all_to_all_single(b.t2, b.t1, self.t2_splits, self.t1_splits, group=self.group)
                    b.poisson_stage(1)
                    d.all_to_all_single(b.t1, b.t2, self.t1_splits, self.t2_splits, group=self.group)
                    b.poisson_stage(2)
                    # my top phi row is the row below my upper neighbour's slab, my bottom row the one above my lower neighbour's
                    self._sendrecv(b.phi_last_row, b.phi_first_row, b.phi_below, b.phi_above)
                    b.poisson_stage(3)
                for w in pending:
                    w.wait()
                b.halo_unpack()

    def refresh_halos(self):
        """Exchange the population halo rows of the current state.  Only a state that a time step produced has anything to
        exchange: plbm_initialize and plbm_upload_state fill a slab's halo rows themselves (every rank writes each row its own
        cells pull from), and the library turns the pack / push / unpack calls into no-ops until the next step -- an exchange
        right after an upload would overwrite the neighbours' correct halo rows with rows the upload never wrote."""
        b = self.b
        with b.stream_context():
            if self.peer:
                b.peer_barrier()          # nobody is still unpacking an earlier exchange
                b.halo_push()
                b.peer_barrier()
                b.halo_unpack()
                return
            b.halo_pack()
            self._sendrecv(b.halo_send_hi, b.halo_send_lo, b.halo_recv_lo, b.halo_recv_hi)
            b.halo_unpack()

    def gather_fields(self, names):
        """All slabs' fields on every rank as full [NY, NX] arrays (tests, output)."""
        import torch
        mine = self.b.fields(names)
        rows = [self.b.slab_y0[s + 1] - self.b.slab_y0[s] for s in range(self.nranks)]
        pad = max(rows)
        nccl = self.dist.get_backend(self.group) == "nccl"
        dev = self.b.halo_send_lo.device if nccl else torch.device("cpu")
        out = {}
        for n in names:
            t = torch.zeros((pad, self.b.NX), dtype=torch.float64)
            t[: rows[self.rank]] = torch.from_numpy(np.ascontiguousarray(mine[n]))
            parts = [torch.empty((pad, self.b.NX), dtype=torch.float64, device=dev) for _ in range(self.nranks)]
            self.dist.all_gather(parts, t.to(dev), group=self.group)
            out[n] = torch.cat([p[: rows[s]].cpu() for s, p in enumerate(parts)], 0).numpy()
        return out
