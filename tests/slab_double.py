"""CPU test double of the per-slab backend (tests only).

Implements the same stage semantics as the CUDA library's multi-slab entry points
(plbm_step_local, plbm_halo_pack/unpack, plbm_poisson_stage) with numpy, using the CPU checker's
cell-local phases and FFT pieces for the arithmetic, and the same buffer layouts.  With it the
product's SlabDriver (12-lb-12-lb_b200/distributed.py) -- neighbour map, message pairing, all-to-all
split sizes, stage order -- runs under torch.distributed's gloo backend without a GPU, and the result
must equal the single-domain checker bit for bit.
"""
import contextlib
import math

import os

import numpy as np
import torch

from oracle import oracle as O

CX = [0, 1, 0, -1, 0, 1, -1, -1, 1]
CY = [0, 0, 1, 0, -1, 1, 1, -1, -1]
DIR_UP = [2, 5, 6]      # cy = +1: leave through the top of the slab
DIR_DOWN = [4, 7, 8]    # cy = -1


def _slab_cut(NY, r, R, alpha):
    """Python restatement of the library's cut (csrc/plbm_api.cu: slab_cut): the rows of the reference's central block weigh 1 + alpha."""
    if r <= 0:
        return 0
    if r >= R:
        return NY
    if alpha == 0.0:
        return 2 * (((NY // 2) * r) // R)
    b0, b1 = float(NY // 4 + 1), float((3 * NY) // 4)
    wb = (1.0 + alpha) * (b1 - b0)
    t = (NY + alpha * (b1 - b0)) * r / R
    y = t if t <= b0 else (b0 + (t - b0) / (1.0 + alpha) if t <= b0 + wb else t - alpha * (b1 - b0))
    return min(max(2 * int(y / 2.0 + 0.5), 0), NY)


def slab_rule(NY, rank, nranks):
    alpha = float(os.environ.get("PLBM_SLAB_ALPHA", "0.17"))
    if not 0.0 <= alpha <= 4.0:
        alpha = 0.0
    if alpha > 0.0 and any(_slab_cut(NY, r + 1, nranks, alpha) - _slab_cut(NY, r, nranks, alpha) < 2 for r in range(nranks)):
        alpha = 0.0                       # small lattices: the equal split
    a, b = _slab_cut(NY, rank, nranks, alpha), _slab_cut(NY, rank + 1, nranks, alpha)
    return a, b - a


class SlabDouble:
    def __init__(self, NX, NY, rank, nranks, poisson="fft"):
        self.NX, self.NY, self.rank, self.nranks = NX, NY, rank, nranks
        self.poisson = poisson
        self.has_poisson = poisson == "fft"
        self.slab_y0 = [slab_rule(NY, r, nranks)[0] for r in range(nranks)] + [NY]
        nh = NY // 2 + 1
        self.nh = nh
        self.slab_k0 = [(nh * r) // nranks for r in range(nranks + 1)]
        self.y0 = self.slab_y0[rank]
        self.nyl = self.slab_y0[rank + 1] - self.y0
        self.nkl = self.slab_k0[rank + 1] - self.slab_k0[rank]
        self.units = O.units_from_si()
        self.cell = O.PortOracle(NX, self.nyl, poisson="none", initialize=False)   # cell-local phases only
        w = [4.0 / 9.0] + [1.0 / 9.0] * 4 + [1.0 / 36.0] * 4
        # post-collision planes [species][kind][dir][row 0..nyl+1][x]; plane value at (x', y') = f_i at (x'+cx, y'+cy)
        self.A = np.zeros((3, 2, 9, self.nyl + 2, NX))
        for rr in range(self.nyl + 2):
            ysg = (self.y0 + rr - 1) % NY
            for i in range(9):
                x = (np.arange(NX) + CX[i]) % NX
                y = (ysg + CY[i]) % NY
                inside = (x >= NX // 4 + 1) & (x < 3 * NX // 4) & (y >= NY // 4 + 1) & (y < 3 * NY // 4)
                for s in range(3):
                    on = np.ones(NX, bool) if s == 2 else inside
                    self.A[s, 0, i, rr] = np.where(on, w[i] * self.units.rho_init[s], 0.0)
                    self.A[s, 1, i, rr] = np.where(on, w[i] * self.units.T_init[s], 0.0)
        self.Ex = np.full((self.nyl, NX), self.units.Ex_ext)
        self.Ey = np.full((self.nyl, NX), self.units.Ey_ext)
        self.rho_q = np.zeros((self.nyl, NX))
        self.phi = np.zeros((self.nyl, NX))
        self.macro = {}
        self.poisson_called = False
        z = lambda n: torch.zeros(n, dtype=torch.float64)
        self.halo_send_lo, self.halo_send_hi, self.halo_recv_lo, self.halo_recv_hi = z(18 * NX), z(18 * NX), z(18 * NX), z(18 * NX)
        self.phi_below, self.phi_above = z(NX), z(NX)
        phi_t = torch.from_numpy(self.phi)
        self.phi_first_row, self.phi_last_row = phi_t[0], phi_t[self.nyl - 1]
        self.t1, self.t2 = z(2 * nh * self.nyl), z(2 * self.nkl * NX)
        self.plan = O.port_lib().offt_plan2d_create(NX, NY)

    # ---- K1 on the slab ---------------------------------------------------------------------
    def step_local(self, want_fields):
        nyl, NX = self.nyl, self.NX
        for s in range(3):
            for kind, dst in ((0, self.cell.f(s)), (1, self.cell.g(s))):
                for i in range(9):
                    rows = self.A[s, kind, i, 1 - CY[i]: 1 - CY[i] + nyl]          # pull from row y - cy (storage row y + 1 - cy)
                    dst[:, :, i] = np.roll(rows, CX[i], axis=1)                      # and from column x - cx
        self.cell.scalar(O.PO_EX)[...] = self.Ex
        self.cell.scalar(O.PO_EY)[...] = self.Ey
        self.cell.update_macro()
        self.macro = {k: v for k, v in self.cell.fields().items() if k not in ("Ex", "Ey", "phi")}
        self.rho_q[...] = self.macro["rho_q"]
        self.cell.compute_equilibrium()
        self.cell.thermal_collisions()
        self.cell.collisions()
        for s in range(3):
            self.A[s, 0, :, 1:nyl + 1] = np.moveaxis(self.cell.f(s), 2, 0)
            self.A[s, 1, :, 1:nyl + 1] = np.moveaxis(self.cell.g(s), 2, 0)

    # ---- K5 -----------------------------------------------------------------------------------
    def halo_pack(self):
        hi = self.halo_send_hi.numpy().reshape(6, 3, self.NX)
        lo = self.halo_send_lo.numpy().reshape(6, 3, self.NX)
        for s in range(3):
            for kind in range(2):
                for j in range(3):
                    hi[s * 2 + kind, j] = self.A[s, kind, DIR_UP[j], self.nyl]
                    lo[s * 2 + kind, j] = self.A[s, kind, DIR_DOWN[j], 1]

    def halo_unpack(self):
        from_below = self.halo_recv_lo.numpy().reshape(6, 3, self.NX)
        from_above = self.halo_recv_hi.numpy().reshape(6, 3, self.NX)
        for s in range(3):
            for kind in range(2):
                for j in range(3):
                    self.A[s, kind, DIR_UP[j], 0] = from_below[s * 2 + kind, j]
                    self.A[s, kind, DIR_DOWN[j], self.nyl + 1] = from_above[s * 2 + kind, j]

    # ---- spectral Poisson, staged -------------------------------------------------------------
    def poisson_stage(self, stage):
        lib = O.port_lib()
        NX, NY, nyl, nh = self.NX, self.NY, self.nyl, self.nh
        if not self.poisson_called:
            self.poisson_called = True
            self.phi[...] = 0.0
            if self.poisson == "none":
                self.Ex[...] = 0.0; self.Ey[...] = 0.0
        if self.poisson == "none":
            return
        if stage == 0:
            H = np.zeros((nyl, nh), dtype=np.complex128)
            rq = np.ascontiguousarray(self.rho_q)
            lib.offt_rows_fwd(self.plan, rq.ctypes.data, nyl, H.ctypes.data)
            self.t1.numpy().view(np.complex128)[...] = np.ascontiguousarray(H.T).reshape(-1)      # [k][r_local]
        elif stage == 1:
            T2 = self.t2.numpy().view(np.complex128)
            k0 = self.slab_k0[self.rank]
            sx2 = np.array([math.sin(math.pi * (i if i <= NX // 2 else i - NX) / NX) ** 2 for i in range(NX)])
            for kl in range(self.nkl):
                col = np.zeros(NX, dtype=np.complex128)
                for s in range(self.nranks):
                    a, rows = self.slab_y0[s], self.slab_y0[s + 1] - self.slab_y0[s]
                    off = self.nkl * a + kl * rows
                    col[a:a + rows] = T2[off:off + rows]
                lib.offt_col(self.plan, -1, col.ctypes.data)
                siny = math.sin(math.pi * (k0 + kl) / NY)
                denom = 4.0 * (sx2 + siny * siny)
                ok = denom > 1e-15
                with np.errstate(all="ignore"):
                    col = np.where(ok, (col.real / np.where(ok, denom, 1.0)) + 1j * (col.imag / np.where(ok, denom, 1.0)), 0.0)
                col = np.ascontiguousarray(col)
                lib.offt_col(self.plan, +1, col.ctypes.data)
                for s in range(self.nranks):
                    a, rows = self.slab_y0[s], self.slab_y0[s + 1] - self.slab_y0[s]
                    off = self.nkl * a + kl * rows
                    T2[off:off + rows] = col[a:a + rows]
        elif stage == 2:
            H = np.ascontiguousarray(self.t1.numpy().view(np.complex128).reshape(nh, nyl).T)
            out = np.zeros((nyl, NY))
            lib.offt_rows_inv(self.plan, H.ctypes.data, nyl, out.ctypes.data)
            self.phi[...] = out * (1.0 / (NX * NY))
        elif stage == 3:
            ext = np.vstack([self.phi_below.numpy()[None], self.phi, self.phi_above.numpy()[None]])
            self.Ex[...] = -0.5 * (np.roll(self.phi, -1, axis=1) - np.roll(self.phi, 1, axis=1))
            self.Ey[...] = -0.5 * (ext[2:] - ext[:-2])

    def stream_context(self):
        return contextlib.nullcontext()

    def fields(self, names=None):
        all_ = dict(self.macro)
        all_.update(Ex=self.Ex.copy(), Ey=self.Ey.copy(), phi=self.phi.copy(), rho_q=self.rho_q.copy())
        return {n: all_[n] for n in (names or all_.keys())}
