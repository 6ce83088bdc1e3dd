"""N>1 path: (a) GPU test -- element-partitioned run over NCCL equals the single-GPU run
bitwise (needs >= 2 GPUs, skipped otherwise); (b) CPU test -- world_size-2 gloo check that the
halo plans of neighbouring ranks agree slot by slot (host logic of flou_b200_partition_plan)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_partitioned_run_equals_single_gpu_bitwise(gpu):
    if gpu < 2:
        pytest.skip("needs at least 2 GPUs on the box")
    n = 2 if gpu < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", "29517",
           os.path.join(ROOT, "tests", "multigpu_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    sys.stdout.write(out.stdout[-4000:])
    sys.stderr.write(out.stderr[-4000:])
    assert out.returncode == 0
    assert "bitwise_equal=False" not in out.stdout


def _plan_worker(rank, world, port, spec, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "flou.jl_b200"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from common import Case
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        case = Case(*spec["args"], **spec["kw"])
        disc, _ = case.product(rank=rank, nranks=world, create=False)
        plan = disc.partition_plan()
        mine = dict(rank=rank, begin=disc.elem_begin, end=disc.elem_end, peers=plan["peers"],
                    counts=plan["counts"], faces=plan["faces"].tolist(),
                    elemfaces=plan["elemfaces"].tolist(),
                    n_interior=plan["n_interior"], n_boundary=plan["n_boundary"])
        allp = [None] * world
        dist.all_gather_object(allp, mine)
        errors = []
        # ranges tile the global element order
        if allp[0]["begin"] != 0 or any(allp[i]["end"] != allp[i + 1]["begin"] for i in range(world - 1)):
            errors.append("ranges do not tile")
        off = 0
        for peer, cnt in zip(mine["peers"], mine["counts"]):
            theirs = allp[peer]
            # the slots this rank expects from `peer` are, in order, the slots `peer` sends here
            toff = 0
            for p2, c2 in zip(theirs["peers"], theirs["counts"]):
                if p2 == rank:
                    break
                toff += c2
            else:
                errors.append(f"rank {peer} has no slots for rank {rank}")
                continue
            if theirs["faces"][toff:toff + cnt] != mine["faces"][off:off + cnt]:
                errors.append(f"slot order differs between ranks {rank} and {peer}")
            off += cnt
        if mine["n_interior"] + mine["n_boundary"] != disc.elem_end - disc.elem_begin:
            errors.append("interior + boundary != owned elements")
        q.put((rank, errors, mine["n_boundary"], len(mine["faces"])))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("spec", [
    dict(args=(3, (3, 3, 4), 3), kw=dict(eq="euler")),
    dict(args=(2, (4, 5), 4), kw=dict(eq="euler")),
    dict(args=(2, (4, 6), 3), kw=dict(eq="adv", op="strong", nf="lxf", avg="std", nodes="GL",
                                      periodic=[("1", "2")],
                                      bcs={"3": ("outflow", None), "4": ("outflow", None)})),
    dict(args=(1, (9,), 4), kw=dict(eq="euler")),
], ids=lambda s: f"{s['args']}")
def test_halo_plans_agree_between_ranks_gloo(spec):
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (abs(hash(str(spec))) % 300)
    procs = [ctx.Process(target=_plan_worker, args=(r, world, port, spec, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, errors, nb, ng in results:
        assert errors == [], f"rank {rank}: {errors}"
        assert ng > 0 and nb > 0
