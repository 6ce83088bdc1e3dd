"""Worker for the multi-GPU parity test: run under
    python -m torch.distributed.run --nproc-per-node N tests/multigpu_worker.py
Every rank owns a contiguous element range, halo traces travel over NCCL; the gathered state
after a few RK steps and the RHS must equal the single-GPU result BITWISE (both owners of a
partition-boundary face evaluate the flux with the same (master, slave) argument order)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "flou.jl_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import flou_b200 as F
    from common import Case, smooth_state

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    cases = [
        (Case(3, (4, 4, 2 * world), 4, eq="euler", op="split", nf="mat", avg="cha"), 6),
        (Case(2, (5, 3 * world), 5, eq="euler", op="split", nf="mat", avg="cha"), 6),
        (Case(2, (6, 2 * world + 1), 4, nodes="GL", eq="adv", op="strong", nf="lxf", avg="std"), 6),
        # row f2: HybridDivOperator (line kernel only) across a partition
        (Case(2, (4, 3 * world), 4, eq="euler", op="hybrid", nf="mat", avg="cha", blend=1.0), 4),
        (Case(3, (3, 3, world), 3, eq="euler", op="split", nf="mat", avg="cha", periodic=[("5", "6")],
              bcs={"1": ("inflow", [1.0, 0.4, 0.0, 0.1, 2.7]), "2": ("outflow", None),
                   "3": ("slip", None), "4": ("slip", None)}), 4),
    ]
    ok = True
    for case, nsteps in cases:
        disc, eq = case.product(rank=rank, nranks=world, device=local)
        ids = [F.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        disc.comm_init(ids[0])
        # global smooth state (every rank builds the same one), local rows for this rank
        full, eq1 = case.product(rank=0, nranks=1, device=local, create=(rank == 0))
        Qg = smooth_state(full.coords(), case.nd, case.eq)
        rows = disc.local_rows()
        Q = np.asfortranarray(Qg[rows])
        dQ = disc.new_state()
        F.rhs(dQ, Q, F.EquationConfig(disc, eq), 0.0)
        dtn = F.get_max_dt(Q, disc, eq, 0.4)       # ncclAllReduce(min) over the ranks (row f1)
        # row f3: monitors are ncclAllReduce(sum) over the ranks; the limiter is element-local
        mon = lim = None
        if case.eq == "euler":
            mon = tuple(F.get_monitor(disc, eq, name)(Q, disc, eq) for name in ("kinetic_energy", "entropy"))
            lim = Q.copy(order="F")
            F.get_limiter(disc, eq, "zhang_shu", 0.95)(lim, disc, eq)
        u = Q.copy(order="F")
        sol, _ = F.timeintegrate(u, disc, eq, F.ORK256(), nsteps * 1e-3, dt=1e-3)
        assert sol is not None
        parts = [None] * world
        dist.all_gather_object(parts, (dQ, sol.u[-1], dtn, mon, lim))
        if rank == 0:
            dQ1 = full.new_state()
            F.rhs(dQ1, Qg, F.EquationConfig(full, eq1), 0.0)
            u1 = Qg.copy(order="F")
            F.timeintegrate(u1, full, eq1, F.ORK256(), nsteps * 1e-3, dt=1e-3)
            dQn = np.concatenate([p[0] for p in parts], axis=0)
            un = np.concatenate([p[1] for p in parts], axis=0)
            dt1 = F.get_max_dt(Qg, full, eq1, 0.4)
            same = (np.array_equal(dQn, dQ1) and np.array_equal(un, u1)
                    and all(p[2] == dt1 for p in parts))
            if case.eq == "euler":
                # the partial sums are added in a different order: 1e-13, identical on all ranks
                for j, name in enumerate(("kinetic_energy", "entropy")):
                    m1 = F.get_monitor(full, eq1, name)(Qg, full, eq1)
                    same = same and all(p[3] == parts[0][3] for p in parts) and abs(parts[0][3][j] / m1 - 1) <= 1e-13
                l1 = Qg.copy(order="F")
                F.get_limiter(full, eq1, "zhang_shu", 0.95)(l1, full, eq1)
                same = same and np.array_equal(np.concatenate([p[4] for p in parts], axis=0), l1) \
                    and not np.array_equal(l1, Qg)
            print(f"[multigpu] {case!r}: ranks={world} bitwise_equal={same} "
                  f"max|d rhs|={np.max(np.abs(dQn - dQ1)):.3e} max|d u|={np.max(np.abs(un - u1)):.3e}",
                  flush=True)
            ok = ok and same
            full.close()
        disc.close()
    flag = [ok]
    dist.broadcast_object_list(flag, src=0)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag[0] else 1)


if __name__ == "__main__":
    main()
