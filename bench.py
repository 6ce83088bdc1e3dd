#!/usr/bin/env python
"""Benchmark of the hot path: DOF-updates/s per RK stage (fp64) of the fused DGSEM
RHS + low-storage RK stage kernel, plus the end-to-end figure through the public API.

    python bench.py --gpus N --steps K --warmup W [--workload cfg4|cfg3|cfg2|cfg1]
    python bench.py --impl reference ...      # CPU restatement of Flou's path (oracle)

A "step" is one RK time step (5 ORK256 stages = 5 fused kernel passes) over the whole
mesh.  Default workload: BASELINE.json configs[3] -- 3-D Euler Taylor-Green vortex, 128^3
hexes, p=4, EC split form + matrix dissipation, element-partitioned over the N GPUs
(strong scaling; fits one B200: 3 x 10.5 GB of state).  Inputs are synthetic (the TGV
initial condition of SURVEY.md 8(d)).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "flou.jl_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "DOF-updates/sec per RK stage (fp64)"
UNIT = "DOF-updates/s"

WORKLOADS = {
    # name: (nd, elements per direction, np, equation, description)
    "cfg1": dict(nd=2, n=(32, 32), np=4, eq="adv", dt=1e-3,
                 desc="2D linear advection 32x32 p=3 GLL StrongDiv+LxF ORK256"),
    "cfg2": dict(nd=2, n=(128, 128), np=5, eq="euler", dt=1e-3,
                 desc="2D Euler isentropic vortex 128x128 p=4 GLL SplitDiv(Chandrasekhar)+MatrixDissipation ORK256"),
    "cfg3": dict(nd=3, n=(32, 32, 32), np=4, eq="euler", dt=1e-3,
                 desc="3D Euler Taylor-Green vortex 32^3 p=3 GLL SplitDiv(Chandrasekhar)+MatrixDissipation ORK256"),
    "cfg3b": dict(nd=3, n=(64, 64, 64), np=4, eq="euler", dt=5e-4,
                  desc="3D Euler Taylor-Green vortex 64^3 p=3 GLL SplitDiv(Chandrasekhar)+MatrixDissipation ORK256"),
    "cfg4": dict(nd=3, n=(128, 128, 128), np=5, eq="euler", dt=1e-4,
                 desc="3D Euler Taylor-Green vortex 128^3 p=4 GLL SplitDiv(Chandrasekhar)+MatrixDissipation ORK256"),
    "cfg4s": dict(nd=3, n=(64, 64, 64), np=5, eq="euler", dt=2e-4,
                  desc="3D Euler Taylor-Green vortex 64^3 p=4 GLL SplitDiv(Chandrasekhar)+MatrixDissipation ORK256"),
    "cfg5b": dict(nd=2, n=(73 * 1024,), np=6, eq="euler", dt=5e-6, unstructured=32,
                  desc="2D Euler on the reference's 2D_cylinder quad mesh (73 quads) refined 32x32 (74 752 quads), "
                       "p=5 GLL, SplitDiv(Chandrasekhar)+MatrixDissipation, slip walls + inflow/outflow, ORK256"),
    "cfg5": dict(nd=2, n=(73 * 64,), np=6, eq="euler", dt=2e-5, unstructured=8,
                 desc="2D Euler on the reference's 2D_cylinder quad mesh (73 quads) refined 8x8, p=5 GLL, "
                      "SplitDiv(Chandrasekhar)+MatrixDissipation, slip walls + inflow/outflow, ORK256"),
}
GAMMA = 1.4


# ----------------------------------------------------------------------------- synthetic ICs
def initial_condition(x, w):
    """Gaussian bump (cfg1), isentropic vortex (cfg2), Taylor-Green (3-D)."""
    nd = w["nd"]
    if w["eq"] == "adv":
        return np.exp(-((x[:, 0] - 0.5) ** 2) / (2 * 0.1 ** 2)
                      - ((x[:, 1] - 0.5) ** 2) / (2 * 0.1 ** 2))[:, None]
    Q = np.empty((x.shape[0], nd + 2))
    if nd == 2:
        beta = 5.0
        xr, yr = x[:, 0] - 5.0, x[:, 1] - 5.0
        r2 = xr * xr + yr * yr
        e = np.exp(0.5 * (1 - r2))
        u = 1.0 - beta / (2 * np.pi) * yr * e
        v = 1.0 + beta / (2 * np.pi) * xr * e
        T = 1.0 - (GAMMA - 1) * beta ** 2 / (8 * GAMMA * np.pi ** 2) * np.exp(1 - r2)
        rho = T ** (1 / (GAMMA - 1))
        p = rho * T
        Q[:, 0], Q[:, 1], Q[:, 2] = rho, rho * u, rho * v
        Q[:, 3] = p / (GAMMA - 1) + 0.5 * rho * (u * u + v * v)
        return Q
    M0 = 0.1
    X, Y, Z = x[:, 0], x[:, 1], x[:, 2]
    u = np.sin(X) * np.cos(Y) * np.cos(Z)
    v = -np.cos(X) * np.sin(Y) * np.cos(Z)
    p = 1.0 / (GAMMA * M0 ** 2) + (np.cos(2 * X) + np.cos(2 * Y)) * (np.cos(2 * Z) + 2) / 16
    Q[:, 0], Q[:, 1], Q[:, 2], Q[:, 3] = 1.0, u, v, 0.0
    Q[:, 4] = p / (GAMMA - 1) + 0.5 * (u * u + v * v)
    return Q


def domain(w):
    nd = w["nd"]
    if w["eq"] == "adv":
        return [0.0] * nd, [1.0] * nd
    if nd == 2:
        return [0.0, 0.0], [10.0, 10.0]
    return [0.0] * 3, [2 * np.pi] * 3


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi sampled DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.path = device, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [s.strip() for s in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(names, f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if not sm:      # region shorter than one sampling period: one query right after it
            try:
                o = subprocess.run(["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.Q}",
                                    "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                   timeout=20).stdout
                f = [x.strip() for x in o.strip().split(",")]
                sm.append(float(f[1])); mx.append(float(f[2]))
                for name, val in zip(names, f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
                out["note"] = "timed region shorter than the sampling period; sampled right after it"
            except Exception:
                pass
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["sm_max_mhz"] = float(max(mx))
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload):
    """DRAM bytes per launch of the stage kernel from the committed ncu capture, if any."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        return d.get(workload, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


# ----------------------------------------------------------------------------- CPU baseline
def cpu_sample_case(w):
    """A bounded sample of the same workload for the CPU restatement (same operator, fluxes,
    node count; fewer elements)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from common import Case
    nd = w["nd"]
    n = {2: (48, 48), 3: (16, 16, 16)}[nd] if w["eq"] == "euler" else (32, 32)
    # (config 5: the CPU sample is the same operator/order on a Cartesian quad mesh)
    if w["eq"] == "adv":
        return Case(nd, n, w["np"], nodes="GLL", eq="adv", op="strong", nf="lxf", avg="std",
                    a=(2.0, -1.0, 0.0)), n
    return Case(nd, n, w["np"], nodes="GLL", eq="euler", op="split", nf="mat", avg="cha"), n


def run_cpu(w, steps, warmup):
    """Times the oracle (CPU restatement of Flou's rhs! + ORK256 loop, OpenMP on all host
    threads) on the bounded sample.  Returns (DOF-updates/s per stage, cores, sample text,
    ms per step)."""
    import oracle as O
    case, n = cpu_sample_case(w)
    orc = case.oracle()
    start, finish = domain(w)
    # map the unit-box oracle coordinates onto the workload's domain for the IC
    lo, hi = np.zeros(w["nd"]), np.array([1.0 + 0.5 * d for d in range(w["nd"])])
    x = (orc.coords - lo) / (hi - lo) * (np.array(finish) - np.array(start)) + np.array(start)
    Q = np.asfortranarray(initial_condition(x, w))
    cores = os.cpu_count() or 1
    u = orc.lsrk2n(Q, O.ORK256, 1e-5, warmup) if warmup > 0 else Q
    t = time.perf_counter()
    orc.lsrk2n(u, O.ORK256, 1e-5, steps)
    dt = time.perf_counter() - t
    rate = orc.ndof * 5 * steps / dt
    sample = (f"{'x'.join(map(str, n))} elements p={w['np'] - 1} ({orc.ndof} DOF), "
              f"{steps} ORK256 steps, oracle/pipeline.c OpenMP")
    return rate, cores, sample, dt / steps * 1e3


# ----------------------------------------------------------------------------- main
def main():
    # keep stdout clean for the ONE JSON line: libraries (NCCL banner, torchrun notices) that
    # print to fd 1 are sent to stderr, the JSON goes to the saved descriptor at the end
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        os.write(json_fd, (json.dumps(line) + "\n").encode())

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS))
    ap.add_argument("--e2e-rk-steps", type=int, default=10,
                    help="RK steps per public-API timeintegrate call in the end-to-end leg")
    ap.add_argument("--e2e-calls", type=int, default=2)
    ap.add_argument("--dt", type=float, default=None, help="override the workload's time step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    nstages = 5
    npts = w["np"] ** w["nd"]
    ndof_global = int(np.prod(w["n"])) * npts
    config = {"workload": w["desc"], "elements": list(w["n"]), "p": w["np"] - 1,
              "ndofs": ndof_global, "nv": 1 if w["eq"] == "adv" else w["nd"] + 2,
              "rk": "ORK256 (5 stages)", "partition": f"contiguous element ranges x{world}",
              "l2_policy": "inputs larger than L2 (state >> 126 MB)" if ndof_global * 40 > 2.5e8
              else "state fits L2; timed back-to-back as the RK loop runs it"}

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        steps = max(1, min(args.steps, 5))
        warm = max(0, min(args.warmup, 1))
        rate, cores, sample, ms = run_cpu(w, steps, warm)
        line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT,
                "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": ms,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                 "sample": sample},
                "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return

    # ------------------------------------------------------------------ B200 arm
    import flou_b200 as F
    from flou_b200 import geometry as G
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        gloo = dist.new_group(backend="gloo")

    nd = w["nd"]
    bcs = {}
    if w.get("unstructured"):
        g = np.load(os.path.join(ROOT, "tests", "golden", "cylinder_p5.npz"))
        groups = [(str(n), [int(v) for v in str(e).split(",")])
                  for n, e in zip(g["group_names"], g["group_entities"])]
        raw = F.RawMesh(g["nodes"], g["quads"], g["lines"], g["line_tags"], g["line_entity"], groups)
        mesh = F.UnstructuredMesh(2, raw, refinement=w["unstructured"])
        Qinf = F.vars_prim2cons((1.0, 0.5, 0.0, 1.0), F.EulerEquation(2, GAMMA))
        for name in mesh.bdnames:
            bcs[name] = (F.EulerInflowBC(Qinf) if name == "Left" else
                         F.EulerOutflowBC() if name == "Right" else F.EulerSlipBC())
    else:
        start, finish = domain(w)
        mesh = F.CartesianMesh(nd, start, finish, w["n"])
        mesh.apply_periodicBCs(*[(str(2 * d + 1), str(2 * d + 2)) for d in range(nd)])
    basis = F.LagrangeBasis("GLL", w["np"])
    if w["eq"] == "adv":
        eq = F.LinearAdvection(2.0, -1.0)
        op = F.StrongDivOperator(F.LxF(F.StdAverage(), 1.0))
    else:
        eq = F.EulerEquation(nd, GAMMA)
        op = F.SplitDivOperator(F.ChandrasekharAverage(),
                                F.MatrixDissipation(F.ChandrasekharAverage(), 1.0))
    std = {2: F.StdQuad, 3: F.StdHex}[nd](basis, F.DGSEMrec(basis), eq.nv)
    disc = F.MultielementDisc(mesh, std, eq, op, bcs, rank=rank, nranks=world, device=local_rank)
    if world > 1:
        ids = [F.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0, group=gloo)
        disc.comm_init(ids[0])

    # synthetic initial condition for the owned elements, filled in slabs (host memory)
    Q = disc.new_state()
    verts = mesh.element_vertices()[disc.elem_begin:disc.elem_end]
    chunk = 1 << 15
    for e0 in range(0, verts.shape[0], chunk):
        x = G.element_coords(verts[e0:e0 + chunk], std.xi)
        if w.get("unstructured"):      # uniform farfield + deterministic smooth perturbation
            q0 = np.tile(Qinf, (x.shape[0], 1))
            q0[:, 0] *= 1 + 0.02 * np.sin(3 * x[:, 0]) * np.cos(2 * x[:, 1])
            Q[e0 * npts:e0 * npts + x.shape[0], :] = q0
            continue
        Q[e0 * npts:(e0 + verts[e0:e0 + chunk].shape[0]) * npts, :] = initial_condition(x, w)
    del verts
    pinned = False
    try:
        F.lib().flou_b200_pin_host(Q.ctypes.data, Q.nbytes)
        pinned = True
    except Exception:
        pass
    solver = F.ORK256(williamson_condition=False)
    dt = args.dt if args.dt is not None else w["dt"]

    def barrier():
        disc.synchronize()
        if dist is not None:
            dist.barrier()

    def global_max(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident leg: `value`
    disc.upload(Q)
    F.advance(disc, solver, dt, args.warmup)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = disc.kernel_launches()
    barrier()
    disc.timer_start()
    t0 = time.perf_counter()
    F.advance(disc, solver, dt, args.steps)
    ms_dev = disc.timer_stop()          # CUDA events on the launching stream
    barrier()
    wall = time.perf_counter() - t0
    launches = disc.kernel_launches() - l0
    clocks = sampler.stop() if rank == 0 else None
    ms_dev = global_max(ms_dev)
    if disc.status() & 1 and not os.environ.get("FLOU_BENCH_IGNORE_STATUS"):
        raise SystemExit("bench: state left the admissible set (negative density/pressure)")
    value = ndof_global * nstages * args.steps / (ms_dev * 1e-3)

    # ---- per-kernel split of a stage (CUDA events around each launch, direct launches)
    ksplit = None
    if world == 1:
        disc.profile(True)
        F.advance(disc, solver, dt, 2)
        ms_f, ms_e, npass = disc.profile(False)
        if npass > 0:
            ksplit = {"face_flux_kernel_ms": ms_f / npass, "element_kernel_ms": ms_e / npass,
                      "passes": int(npass)}

    # ---- end-to-end leg through the public API: host buffer in, host buffer out, per call
    e2e = None
    if not args.no_e2e:
        m = args.e2e_rk_steps
        u = Q                                   # alias_u0=True: integrates in place
        F.timeintegrate(u, disc, eq, solver, m * dt, dt=dt, nsteps=m, save_start=False)   # warm-up call
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_calls):
            sol, _ = F.timeintegrate(u, disc, eq, solver, m * dt, dt=dt, nsteps=m, save_start=False)
            if sol is None:
                raise SystemExit("bench: simulation crashed in the end-to-end leg")
        barrier()
        te = global_max(time.perf_counter() - t0)
        e2e = {"value": ndof_global * nstages * m * args.e2e_calls / te, "unit": UNIT,
               "h2d_bytes_per_step": int(ndof_global * eq.nv * 8),
               "d2h_bytes_per_step": int(ndof_global * eq.nv * 8),
               "step": f"one timeintegrate() call = upload + {m} RK steps + download",
               "calls": args.e2e_calls, "pinned_host": pinned, "ms_per_call": te / args.e2e_calls * 1e3}

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant (only) kernel: the fused stage kernel
    peak, peak_src = hbm_peak()
    bytes_per_dof = 32 * eq.nv                      # read u,tmp + write u,tmp
    if w.get("unstructured"):
        bytes_per_dof += 8 * (nd * nd + 1)          # per-node metric + jac (config 5: 168 B)
    ndof_local = disc.ndofs
    stages = nstages * args.steps
    stage_ms = ms_dev / stages                      # whole stage: every kernel of one RK stage
    stage_gbs = ndof_local * bytes_per_dof / (stage_ms * 1e-3) / 1e9
    # dominant kernel = the element kernel (line_kernel_ws: volume + lift + RK update; it moves all
    # of the algorithmic bytes); the face-flux kernel only adds non-algorithmic traffic
    kernel_ms = ksplit["element_kernel_ms"] if ksplit else stage_ms
    achieved = ndof_local * bytes_per_dof / (kernel_ms * 1e-3) / 1e9
    config["launch"] = disc.kernel_info()
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": ncu_traffic(args.workload),
                "kernel": "flou::line_kernel_ws (element kernel of the two-kernel stage)",
                "algorithmic_bytes_per_dof": bytes_per_dof,
                "dofs_per_launch": ndof_local, "avg_launch_ms": kernel_ms,
                "stage_ms": stage_ms, "stage_achieved": stage_gbs, "stage_frac": stage_gbs / peak,
                "kernels_per_stage": ksplit, "peak_source": peak_src,
                "note": "frac = dominant kernel alone; stage_frac = all kernels of an RK stage "
                        "(what `value` is made of)"}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        rate, cores, sample, _ = run_cpu(w, 2, 1)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu, "wall_ms_per_step": wall / args.steps * 1e3}
    emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
