"""Host-side logic of the N > 1 path on CPU: two gloo ranks each own a slab of the state grid
(planned by the library's own reach analysis, bellman_plan_slabs), compute their slab with the
oracle, and exchange exactly the halo ranges the library exchanges (owner's slab ∩ reader's
extended range).  The stitched result must equal the single-process sweep bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_desc(kind):
    import bellman_b200 as bb
    if kind == "kirk":
        o = bb.Dynamic_Solver()
        o.dx, o.du, o.N = 48, 24, 9
        return o._build(), 1
    sp = bb.Solver_pos_att()
    sp.n_mesh_x, sp.n_mesh_v, sp.n_mesh_t, sp.n_mesh_w = 8, 7, 6, 9
    return sp.channel_desc(0), 3


def _exchange(J, slabs, rank, world, inner, outer):
    """Fill my halo from the owners: same pairing rule as exchange_halo() in bellman_api.cu."""
    n_p = J.shape[1]
    me = slabs[rank]
    view = J.reshape(J.shape[0], outer, n_p, inner) if False else None
    reqs = []
    bufs = []
    for q in range(world):
        if q == rank:
            continue
        o = slabs[q]
        slo, shi = max(o[2], me[0]), min(o[3], me[1])
        rlo, rhi = max(me[2], o[0]), min(me[3], o[1])
        if slo < shi:
            t = torch.from_numpy(np.ascontiguousarray(J[:, :, slo:shi, :]))
            reqs.append(dist.isend(t, q))
        if rlo < rhi:
            t = torch.empty((J.shape[0], J.shape[1], rhi - rlo, J.shape[3]), dtype=torch.float64)
            reqs.append(dist.irecv(t, q))
            bufs.append((rlo, rhi, t))
    for r in reqs:
        r.wait()
    for rlo, rhi, t in bufs:
        J[:, :, rlo:rhi, :] = t.numpy()


def _worker(rank, world, port, kind, n_stages, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bellman_b200 as bb
    from oracle import cbind
    d, part_dim = _make_desc(kind)
    slabs = bb.plan_slabs(d, part_dim, world)
    lo, hi, elo, ehi = slabs[rank]
    inner = int(np.prod(d.n[:part_dim]))
    outer = int(np.prod(d.n[part_dim + 1:]))
    n_p = d.n[part_dim]
    # full-size array; only [elo, ehi) along part_dim is ever valid on this rank
    J = np.full((d.P, outer, n_p, inner), np.nan)
    J[:, :, elo:ehi, :] = 0.0                                  # J_N = 0
    idx = None
    for _ in range(n_stages):
        Jn = np.where(np.isnan(J), 1e300, J).reshape(d.P, -1)  # poison outside my extended range
        Jo, Io = cbind.stage(d, Jn, part_dim=part_dim, own_lo=lo, own_hi=hi)
        assert np.all(np.abs(Jo.reshape(J.shape)[:, :, lo:hi, :]) < 1e200), "read outside the planned halo"
        J = np.full_like(J, np.nan)
        J[:, :, lo:hi, :] = Jo.reshape(J.shape)[:, :, lo:hi, :]
        idx = Io.reshape(J.shape)[:, :, lo:hi, :].copy()
        _exchange(J, slabs, rank, world, inner, outer)
        assert not np.isnan(J[:, :, elo:ehi, :]).any(), "halo exchange left a hole"
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), J=J[:, :, lo:hi, :], idx=idx, lo=lo, hi=hi)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("kind,world", [("kirk", 2), ("pos_att", 2), ("kirk", 3)])
def test_two_rank_slab_sweep_matches_single_process(tmp_path, oracle_lib, bellman, kind, world):
    n_stages = 4
    port = _free_port()
    mp.spawn(_worker, args=(world, port, kind, n_stages, str(tmp_path)), nprocs=world, join=True)
    d, part_dim = _make_desc(kind)
    ref = oracle_lib.sweep(d, n_stages=n_stages)
    inner = int(np.prod(d.n[:part_dim]))
    outer = int(np.prod(d.n[part_dim + 1:]))
    shape = (d.P, outer, d.n[part_dim], inner)
    Jref = ref["J_last"].reshape(shape)
    Iref = ref["idx_last"].reshape(shape)
    covered = 0
    for r in range(world):
        z = np.load(tmp_path / f"rank{r}.npz")
        lo, hi = int(z["lo"]), int(z["hi"])
        assert np.array_equal(z["J"], Jref[:, :, lo:hi, :])
        assert np.array_equal(z["idx"], Iref[:, :, lo:hi, :])
        covered += hi - lo
    assert covered == d.n[part_dim]
