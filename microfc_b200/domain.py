"""Domain decomposition and ghosted grid metrics: the host-side work the unchanged Fortran
host does before it reaches the hot path, needed here by the Python driver that plays
``p_main`` (no Fortran compiler exists in this image).

* :func:`processor_topology` -- ``s_mpi_decompose_computational_domain``
  (src/simulation/m_mpi_proxy.fpp:134-328): px x py minimising |Mx/px - Ny/py| subject to
  >= 5*weno_order cells per rank and direction, ties to the larger px; 1-D splits in x.
  Three factors in 3-D (EXTENSION).
* :class:`RankLayout` -- local sizes (remainder cells to the lowest coordinates, :229-239,
  :287-297), ``start_idx`` (:257-263,:313-319) and neighbour ranks on the periodic, row-major
  Cartesian communicator (:215-222,:242-255,:300-311).
* :func:`ghosted_metrics` -- ``s_read_parallel_data_files`` metrics
  (src/simulation/m_start_up.fpp:434-438) and ``s_populate_grid_variables_buffers``
  (:517-655), where processor boundaries receive the neighbour's cell widths
  (m_mpi_proxy.fpp:338-455).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Tuple

import numpy as np

from .case import CaseConfig

NUM_STCLS_MIN = 5   # m_global_parameters.fpp:33


def processor_topology(num_procs: int, cfg: CaseConfig) -> Tuple[int, int, int]:
    lim = NUM_STCLS_MIN * cfg.weno_order
    Mx, Ny, Nz = cfg.m + 1, cfg.n + 1, cfg.p + 1
    nd = cfg.num_dims
    if nd == 1:
        return (num_procs, 1, 1)
    best = None
    if nd == 2:
        px, py = 1, num_procs
        fct_min = 10.0 * abs(Mx / float(px) - Ny / float(py))
        for i in range(1, num_procs + 1):
            if num_procs % i == 0 and Mx // i >= lim:
                j = num_procs // i
                f = abs(Mx / float(i) - Ny / float(j))
                if fct_min >= f and Ny / float(j) >= lim:
                    px, py, fct_min, best = i, j, f, True
        if best is None and num_procs > 1:
            raise ValueError("Unsupported combination of values of num_procs, m, n and weno_order. Exiting ...")
        return (px, py, 1)
    px, py, pz = 1, 1, num_procs
    fct_min = 10.0 * abs(Mx / float(px) - Ny / float(py)) + 10.0 * abs(Ny / float(py) - Nz / float(pz))
    for i in range(1, num_procs + 1):
        if num_procs % i == 0 and Mx // i >= lim:
            for j in range(1, num_procs // i + 1):
                if (num_procs // i) % j == 0 and Ny // j >= lim:
                    k = num_procs // (i * j)
                    f = abs(Mx / float(i) - Ny / float(j)) + abs(Ny / float(j) - Nz / float(k))
                    if fct_min >= f and Nz / float(k) >= lim:
                        px, py, pz, fct_min, best = i, j, k, f, True
    if best is None and num_procs > 1:
        raise ValueError("Unsupported combination of values of num_procs, m, n, p and weno_order. Exiting ...")
    return (px, py, pz)


@dataclass
class RankLayout:
    rank: int
    num_procs: int
    np_dir: Tuple[int, int, int]
    coords: Tuple[int, int, int]
    N: List[int]                 # local m, n, p
    start_idx: List[int]
    bc: List[List[int]]          # [dir][side]: < 0 physical code, >= 0 neighbour rank

    @property
    def shape(self):
        """(Nz, Ny, Nx) local cells."""
        return (self.N[2] + 1, self.N[1] + 1, self.N[0] + 1)

    def interior_slices(self):
        s = self.start_idx
        return (slice(s[2], s[2] + self.N[2] + 1), slice(s[1], s[1] + self.N[1] + 1), slice(s[0], s[0] + self.N[0] + 1))


def _cart_rank(np_dir, c, nd):
    c = [c[d] % np_dir[d] for d in range(3)]
    if nd == 1:
        return c[0]
    if nd == 2:
        return c[0] * np_dir[1] + c[1]
    return (c[0] * np_dir[1] + c[1]) * np_dir[2] + c[2]


def rank_layout(rank: int, num_procs: int, cfg: CaseConfig) -> RankLayout:
    nd = cfg.num_dims
    npd = processor_topology(num_procs, cfg)
    if nd == 1:
        coords = [rank, 0, 0]
    elif nd == 2:
        coords = [rank // npd[1], rank % npd[1], 0]
    else:
        coords = [rank // (npd[1] * npd[2]), (rank // npd[2]) % npd[1], rank % npd[2]]
    Nglb = [cfg.m, cfg.n, cfg.p]
    N, start, bc = [0, 0, 0], [0, 0, 0], [list(b) for b in cfg.bc]
    for d in range(nd):
        rem = (Nglb[d] + 1) % npd[d]
        Nl = (Nglb[d] + 1) // npd[d] - 1
        if coords[d] < rem:
            Nl += 1
        N[d] = Nl
        start[d] = (Nl + 1) * coords[d] if coords[d] < rem else (Nl + 1) * coords[d] + rem
        if num_procs > 1:
            c = list(coords)
            if coords[d] > 0 or cfg.bc[d][0] == -1:
                c[d] = coords[d] - 1
                bc[d][0] = _cart_rank(npd, c, nd)
            c = list(coords)
            if coords[d] < npd[d] - 1 or cfg.bc[d][1] == -1:
                c[d] = coords[d] + 1
                bc[d][1] = _cart_rank(npd, c, nd)
    return RankLayout(rank, num_procs, npd, tuple(coords), N, start, bc)


@dataclass
class Metrics:
    """Ghosted metric arrays of one rank, stored with element 0 <-> the reference's lowest
    index: cb(-1-b:N+b), cc(-b:N+b), ds(-b:N+b)."""
    cb: List[np.ndarray] = field(default_factory=list)
    cc: List[np.ndarray] = field(default_factory=list)
    ds: List[np.ndarray] = field(default_factory=list)


def ghosted_metrics(layout: RankLayout, cfg: CaseConfig, cb_glb: List[np.ndarray],
                    all_layouts: List[RankLayout] | None = None) -> Metrics:
    """Metrics of ``layout``'s rank.  ``all_layouts`` (every rank's layout) is needed when a
    side is a processor boundary, to fetch the neighbour's cell widths."""
    b = cfg.buff_size
    out = Metrics()
    for d in range(cfg.num_dims):
        N = layout.N[d]
        cb = np.zeros(N + 2 + 2 * b); cc = np.zeros(N + 1 + 2 * b); ds = np.zeros(N + 1 + 2 * b)
        oc = b + 1       # cb index of element 0 is oc (cb(-1-b) is element 0)
        o = b            # cc/ds index of element 0
        s0 = layout.start_idx[d]
        cb[oc - 1:oc + N + 1] = cb_glb[d][s0:s0 + N + 2]                       # x_cb(-1:m), m_start_up.fpp:434
        ds[o:o + N + 1] = cb[oc:oc + N + 1] - cb[oc - 1:oc + N]                # :436
        cc[o:o + N + 1] = cb[oc - 1:oc + N] + ds[o:o + N + 1] / 2.0            # :438

        def nb_widths(nb_rank, first: bool):
            nl = all_layouts[nb_rank]
            s = nl.start_idx[d]
            w = cb_glb[d][s + 1:s + nl.N[d] + 2] - cb_glb[d][s:s + nl.N[d] + 1]
            return w[:b] if first else w[nl.N[d] + 1 - b:]

        bcb, bce = layout.bc[d]
        for i in range(1, b + 1):                                              # :527-541
            if bcb <= -3: ds[o - i] = ds[o]
            elif bcb == -2: ds[o - i] = ds[o + i - 1]
            elif bcb == -1: ds[o - i] = ds[o + N - (i - 1)]
        if bcb >= 0:
            ds[o - b:o] = nb_widths(bcb, first=False)
        for i in range(1, b + 1):
            cb[oc - 1 - i] = cb[oc - i] - ds[o - i]                            # :545-547
        for i in range(1, b + 1):
            cc[o - i] = cc[o + 1 - i] - (ds[o + 1 - i] + ds[o - i]) / 2.0      # :550-552
        for i in range(1, b + 1):                                              # :558-572
            if bce <= -3: ds[o + N + i] = ds[o + N]
            elif bce == -2: ds[o + N + i] = ds[o + N - (i - 1)]
            elif bce == -1: ds[o + N + i] = ds[o + i - 1]
        if bce >= 0:
            ds[o + N + 1:o + N + 1 + b] = nb_widths(bce, first=True)
        for i in range(1, b + 1):
            cb[oc + N + i] = cb[oc + N + (i - 1)] + ds[o + N + i]              # :576-578
        for i in range(1, b + 1):
            cc[o + N + i] = cc[o + N + (i - 1)] + (ds[o + N + (i - 1)] + ds[o + N + i]) / 2.0   # :581-583
        out.cb.append(cb); out.cc.append(cc); out.ds.append(ds)
    return out
