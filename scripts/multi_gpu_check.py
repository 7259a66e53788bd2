"""Run under torchrun (one rank per GPU): slab-partitioned sweep through libbellman.so + NCCL halo
exchange, every rank's slab compared bit for bit with the single-process CPU oracle.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P scripts/multi_gpu_check.py [kirk|attitude|pos_att]
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bellman_b200 as bb  # noqa: E402
from oracle import cbind  # noqa: E402


def make(kind):
    if kind == "kirk":
        o = bb.Dynamic_Solver()
        return bb.tables.kirk_desc(o.A, o.B, o.Q, o.R, 12, o.x_min, o.x_max, 256, o.u_min, o.u_max, 48,
                                   store_J_all=False, store_idx_all=False)
    if kind == "kirk_odd":
        # odd leading dimension: D = 2 grids pad it to even (TMA stride rule), which the neighbours'
        # peer-store addressing must reproduce for BOTH partition dimensions
        o = bb.Dynamic_Solver()
        t = bb.tables
        s0, s1, u = t.linspace(-2.5, 3.0, 255), t.linspace(-2.5, 3.0, 120), t.linspace(-40.0, 10.0, 24)
        A, B = o.A, o.B.ravel()
        row = lambda x: np.ascontiguousarray(x).reshape(1, -1)
        return t.Desc(n=[255, 120], C=24, N=12, grid=[row(s0), row(s1)], src_a=[0, 0], src_b=[1, 1],
                      Ta=[row(A[0, 0] * s0), row(A[1, 0] * s0)], Tb=[row(A[0, 1] * s1), row(A[1, 1] * s1)],
                      Tc=[row(B[0] * u), row(B[1] * u)], q_order=[0, 1],
                      q=[row(0.25 * s0 * s0), row(0.05 * s1 * s1)], r=row(0.05 * u * u)).validate()
    if kind == "kirk_host":
        o = bb.Dynamic_Solver()
        t = bb.tables
        s0, s1, u = t.linspace(-2.5, 3.0, 256), t.linspace(-2.5, 3.0, 1024), t.linspace(-40.0, 10.0, 48)
        A, B = o.A, o.B.ravel()
        row = lambda x: np.ascontiguousarray(x).reshape(1, -1)
        return t.Desc(n=[256, 1024], C=48, N=8, grid=[row(s0), row(s1)], src_a=[0, 0], src_b=[1, 1],
                      Ta=[row(A[0, 0] * s0), row(A[1, 0] * s0)], Tb=[row(A[0, 1] * s1), row(A[1, 1] * s1)],
                      Tc=[row(B[0] * u), row(B[1] * u)], q_order=[0, 1],
                      q=[row(0.25 * s0 * s0), row(0.05 * s1 * s1)], r=row(0.05 * u * u),
                      store_J_all=False, store_idx_all=False).validate()
    if kind == "attitude":
        s = bb.Solver_attitude()
        s.n_mesh_w, s.n_mesh_t = 400, 120
        return bb.tables.stack_problems(s._axis_descs())
    s = bb.Solver_pos_att()
    s.n_mesh_x, s.n_mesh_v, s.n_mesh_t, s.n_mesh_w = 12, 10, 8, 15
    return s.channel_desc(0)


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "kirk"
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    d = make(kind)
    ok = True
    for part_dim in ((d.D - 1, 0) if kind != "pos_att" else (d.D - 1, 2)):
        if kind == "kirk_host":
            ok = check_stage_host(d, part_dim, rank, world, local) and ok
            continue
        ok = check(d, kind, part_dim, rank, world, local) and ok
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_CHECK", "PASS" if int(t.item()) == 1 else "FAIL", flush=True)
    sys.exit(0 if int(t.item()) == 1 else 1)


def check(d, kind, part_dim, rank, world, local):
    sw = bb.Sweep(d, device=local, part_dim=part_dim, rank=rank, nranks=world)
    ids = [bb.get_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    sw.comm_init(ids[0])
    n_stages = 6
    ok = True
    for kernel in (bb.KERNEL_DIRECT, bb.KERNEL_AUTO, bb.KERNEL_WINDOW):   # WINDOW: strip / tile kernels with peer stores
        sw.set_J(None)
        sw.run(n_stages, kernel=kernel)
        J, idx = sw.get_J(), sw.get_idx()
        ref = cbind.sweep(d, n_stages=n_stages)
        lo, hi = sw.slab[0], sw.slab[1]
        inner = int(np.prod(d.n[:part_dim]))
        n_p = d.n[part_dim]
        Jr = ref["J_last"].reshape(d.P, -1, n_p, inner)[:, :, lo:hi, :].reshape(d.P, -1)
        Ir = ref["idx_last"].reshape(d.P, -1, n_p, inner)[:, :, lo:hi, :].reshape(d.P, -1)
        good = bool(np.array_equal(J, Jr) and np.array_equal(idx, Ir))
        print(f"rank {rank}/{world} {kind} part_dim={part_dim} kernel={sw.last_kernel} slab={sw.slab} "
              f"halo={os.environ.get('BELLMAN_NO_P2P') and 'nccl' or 'auto'} "
              f"{'OK' if good else 'MISMATCH'} exchange_ms={sw.stats()['ms_exchange']:.3f}", flush=True)
        ok = ok and good
    sw.close()
    return ok


def check_stage_host(d, part_dim, rank, world, local):
    """bellman_stage_host on a sharded handle: three stages driven from host arrays (the first from a rough
    J_N, the next two continuing from the device's J, so they read the halo rows the neighbours stored),
    slab-pipelined for row slabs (part_dim 0), the plain sequence for column slabs."""
    sw = bb.Sweep(d, device=local, part_dim=part_dim, rank=rank, nranks=world)
    ids = [bb.get_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    sw.comm_init(ids[0])
    JN = np.random.default_rng(5).normal(size=(1, d.S)) * 3
    ref = cbind.sweep(d, n_stages=3, J_N=JN, keep_all=True)
    lo, hi = sw.slab[0], sw.slab[1]
    inner = int(np.prod(d.n[:part_dim]))
    n_p = d.n[part_dim]
    ok = True
    for k in range(3):
        J, idx = sw.stage_host(JN if k == 0 else None, kernel=bb.KERNEL_WINDOW)
        stage = d.N - 1 - k
        Jr = ref["J_all"][stage - 1].reshape(d.P, -1, n_p, inner)[:, :, lo:hi, :].reshape(d.P, -1)
        Ir = ref["idx_all"][stage - 1].reshape(d.P, -1, n_p, inner)[:, :, lo:hi, :].reshape(d.P, -1)
        good = bool(np.array_equal(J, Jr) and np.array_equal(idx, Ir)) and sw.current_stage == stage
        print(f"rank {rank}/{world} stage_host part_dim={part_dim} stage={stage} kernel={sw.last_kernel} "
              f"launches={sw.stats()['launches']} halo_mode={sw.halo_mode} {'OK' if good else 'MISMATCH'}", flush=True)
        ok = ok and good
    sw.close()
    return ok


if __name__ == "__main__":
    main()
