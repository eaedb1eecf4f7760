"""Worker of tests/test_multirank_gloo.py (launched with torch.distributed.run, backend gloo,
world_size >= 2, CPU only).  Exercises the host-side multi-rank logic the CUDA library relies
on, in real separate processes:

  * every rank derives its own block from (rank, num_procs, case) alone -- no communication --
    and the blocks agree (domain.rank_layout, m_mpi_proxy.fpp:134-328);
  * the ghost-cell exchange SCHEDULE of mfc_api.cu::fill_ghosts -- direction by direction,
    sends [to beg, to end], receives [from end, from beg], transverse extent including the
    ghosts of earlier directions (m_mpi_proxy.fpp:736-739) -- run over gloo send/recv on numpy
    buffers, followed by the physical boundary fill, yields exactly the globally padded array;
  * rank 0's 128-byte id reaches every rank through the broadcast helper bench.py uses;
  * the stability extrema reduce with MAX/MIN like m_mpi_common.fpp:155-165.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from microfc_b200 import cases, pre_process  # noqa: E402
from microfc_b200.domain import rank_layout  # noqa: E402


def pad_global(q, cfg, b):
    """The global array with b ghost layers per active direction, filled by the physical BCs in
    the reference's order x, y, z (m_rhs.fpp:686-908)."""
    nd, nf = cfg.num_dims, cfg.num_fluids
    out = q
    for d in range(nd):
        ax = 3 - d                                       # (E, z, y, x)
        beg, end = cfg.bc[d]
        n = out.shape[ax]
        idx = lambda s: tuple(s if a == ax else slice(None) for a in range(4))
        if beg == -1:
            lo, hi = out[idx(slice(n - b, n))], out[idx(slice(0, b))]
        else:
            def side(code, first):
                if code <= -3:
                    e = out[idx(slice(0, 1) if first else slice(n - 1, n))]
                    return np.repeat(e, b, axis=ax)
                m = out[idx(slice(0, b) if first else slice(n - b, n))]
                m = np.flip(m, axis=ax).copy()
                m[nf + d] *= -1.0                        # reflective: normal momentum negated
                return m
            lo, hi = side(beg, True), side(end, False)
        out = np.concatenate([lo, out, hi], axis=ax)
    return out


def exchange(qg, lay, cfg, b, world):
    """mfc_api.cu::fill_ghosts over gloo; qg: this rank's ghosted (E, z, y, x) array."""
    nd, nf = cfg.num_dims, cfg.num_fluids
    for d in range(nd):
        ax = 3 - d
        N = lay.N[d] + 1
        def sl(s):
            # earlier directions ghosted, later ones interior only
            out = [slice(None)]
            for a in (2, 1, 0):
                if a == d:
                    out.append(s)
                elif a < d or a >= nd:
                    out.append(slice(None))
                else:
                    out.append(slice(b, b + lay.N[a] + 1))
            return tuple(out)
        nb = lay.bc[d]
        eff = [(-1 if c == lay.rank else c) for c in nb]                    # self-neighbour: plain periodic
        reqs, recv = [], {}
        send_src = {0: slice(b, 2 * b), 1: slice(N, N + b)}                # first b / last b interior layers
        for s in (0, 1):
            if eff[s] >= 0:
                buf = torch.from_numpy(np.ascontiguousarray(qg[sl(send_src[s])]))
                reqs.append(dist.isend(buf, dst=eff[s], tag=s if eff[0] != eff[1] else 0))
        for s in (1, 0):
            if eff[s] >= 0:
                shape = qg[sl(send_src[s])].shape
                recv[s] = torch.empty(shape, dtype=torch.float64)
                reqs.append(dist.irecv(recv[s], src=eff[s], tag=(1 - s) if eff[0] != eff[1] else 0))
        for r in reqs:
            r.wait()
        dst = {0: slice(0, b), 1: slice(N + b, N + 2 * b)}
        for s, t in recv.items():
            qg[sl(dst[s])] = t.numpy()
        for s in (0, 1):                                                    # physical sides (k_bc)
            code = eff[s]
            if code >= 0:
                continue
            for jj in range(1, b + 1):
                if code <= -3:
                    src = b if s == 0 else b + N - 1
                elif code == -2:
                    src = b + jj - 1 if s == 0 else b + N - jj
                else:
                    src = b + N - jj if s == 0 else b + jj - 1
                dsti = b - jj if s == 0 else b + N - 1 + jj
                val = qg[sl(slice(src, src + 1))].copy()
                if code == -2:
                    val[nf + d] *= -1.0
                qg[sl(slice(dsti, dsti + 1))] = val
    return qg


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    ok = True
    for name in sys.argv[1:]:
        d = {"sod_1d": lambda: cases.sod_1d(Nx=99), "shockbubble_2d": lambda: cases.shockbubble_2d_cells(100, 60),
             "shearlayer_2d": lambda: cases.shearlayer_2d(Nx=63, Ny=55), "shockdroplet_2d": lambda: cases.shockdroplet_2d(Nx=99, Ny=59),
             "shockbubble_3d": lambda: cases.shockbubble_3d(nc=50)}[name]()
        cfg = cases.config(d)
        cb = pre_process.generate_grid(cfg)
        q0 = pre_process.generate_initial_condition(cfg, cb)
        b, nd, E = cfg.buff_size, cfg.num_dims, cfg.sys_size
        lay = rank_layout(rank, world, cfg)
        # 1. layouts agree across processes
        mine = [list(lay.N), list(lay.start_idx), [list(x) for x in lay.bc], list(lay.coords)]
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        for r in range(world):
            other = rank_layout(r, world, cfg)
            assert gathered[r] == [list(other.N), list(other.start_idx), [list(x) for x in other.bc], list(other.coords)]
        # 2. halo exchange schedule
        shape = tuple(lay.N[a] + 1 + 2 * b if a < nd else 1 for a in (2, 1, 0))
        qg = np.full((E,) + shape, np.nan)
        inner = tuple([slice(None)] + [slice(b, b + lay.N[a] + 1) if a < nd else slice(0, 1) for a in (2, 1, 0)])
        qg[inner] = q0[(slice(None),) + lay.interior_slices()]
        qg = exchange(qg, lay, cfg, b, world)
        glob = pad_global(q0, cfg, b)
        want = glob[tuple([slice(None)] + [slice(lay.start_idx[a], lay.start_idx[a] + lay.N[a] + 1 + 2 * b) if a < nd else slice(0, 1)
                                            for a in (2, 1, 0)])]
        # later directions' ghosts of EARLIER-direction slabs are not filled (interior-only
        # transverse extent): compare where the schedule defines values
        mask = ~np.isnan(qg)
        assert mask[inner].all()
        for a in range(nd):                                                  # every face slab is defined
            face = [slice(None)] + [slice(b, b + lay.N[c] + 1) if c < nd else slice(0, 1) for c in (2, 1, 0)]
            face[3 - a] = slice(None)
            assert mask[tuple(face)].all(), (name, a)
        same = np.array_equal(qg[mask], want[mask])
        ok = ok and same
        assert same, f"rank {rank}: {name} ghost cells differ from the padded global array"
        # 3. unique-id broadcast helper
        obj = [bytes(range(128)) if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        assert obj[0] == bytes(range(128))
        # 4. stability extrema: MAX / MIN over ranks
        loc = torch.tensor([float(np.abs(qg[inner]).max()), -float(np.abs(qg[inner]).min())], dtype=torch.float64)
        dist.all_reduce(loc, op=dist.ReduceOp.MAX)
        assert loc[0].item() == np.abs(q0).max() and -loc[1].item() == np.abs(q0).min()
    dist.barrier()
    if rank == 0:
        print("GLOO_WORKER_OK" if ok else "GLOO_WORKER_FAIL")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
