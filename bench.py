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
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,enforced.power.limit")

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
        sm, mx, pw, cap, reasons = [], [], [], [], set()
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
                try:
                    pw.append(float(f[3])); cap.append(float(f[9]))
                except (ValueError, IndexError):
                    pass
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
        if pw and cap:      # board power against its enforced limit: what sw_power_cap refers to
            out["power_w"] = float(np.median(pw))
            out["power_limit_w"] = float(max(cap))
        out["reasons"] = sorted(reasons)
        return out


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def fp64_peak():
    """Measured DFMA rate of the device (profiles/tools/dfma_peak.cu, thread-instructions/s at the
    burst clock) and the fp64 thread-instructions per DOF and stage of the two kernels, counted by
    ncu (smsp__sass_thread_inst_executed_op_{dfma,dmul,dadd}_pred_on).  None if not recorded."""
    try:
        peak = json.load(open(os.path.join(ROOT, "profiles", "fp64_peak.json")))
        cnt = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        return peak, cnt
    except Exception:
        return None, None


def ncu_traffic(workload):
    """DRAM bytes per launch of the stage kernel from the committed ncu capture, if any."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        return d.get(workload, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


# ----------------------------------------------------------------------------- CPU baseline
# Bounded sample of each workload for the CPU legs (config 4 does not
# fit / finish on the host path, "time 32^3 ... at p=4 and state the per-DOF rate").
CPU_SAMPLE = {"cfg1": (32, 32), "cfg2": (128, 128), "cfg3": (32, 32, 32), "cfg3b": (32, 32, 32),
              "cfg4": (32, 32, 32), "cfg4s": (32, 32, 32), "cfg5": (96, 96), "cfg5b": (96, 96)}


def cpu_sample_info(name):
    w = WORKLOADS[name]
    n = CPU_SAMPLE[name]
    ndofs = int(np.prod(n)) * w["np"] ** w["nd"]
    text = f"{'x'.join(map(str, n))} elements p={w['np'] - 1} ({ndofs} DOF), same operator / fluxes / IC"
    if w.get("unstructured"):
        text += " on a Cartesian quad mesh (the CPU legs do not read the unstructured mesh)"
    return {"elements": list(n), "p": w["np"] - 1, "ndofs": ndofs, "what": text}


def cpu_sample_case(name):
    """A bounded sample of the same workload for the CPU restatement (same operator, fluxes,
    node count; fewer elements)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from common import Case
    w = WORKLOADS[name]
    nd, n = w["nd"], CPU_SAMPLE[name]
    if w["eq"] == "adv":
        return Case(nd, n, w["np"], nodes="GLL", eq="adv", op="strong", nf="lxf", avg="std",
                    a=(2.0, -1.0, 0.0)), n
    return Case(nd, n, w["np"], nodes="GLL", eq="euler", op="split", nf="mat", avg="cha"), n


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def find_julia():
    """Plan A of BASELINE.md section 4: a Julia that can load the reference (PATH, or an install
    shipped under baseline/_ref/).  Returns (julia, project) or None."""
    import shutil
    exe = shutil.which("julia")
    for cand in (os.path.join(ROOT, "baseline", "_ref", "julia", "bin", "julia"),):
        if exe is None and os.path.exists(cand):
            exe = cand
    if exe is None:
        return None
    for proj in (os.path.join(ROOT, "baseline", "_ref", "Flou.jl"), "/root/reference"):
        if os.path.exists(os.path.join(proj, "Project.toml")):
            return exe, proj
    return None


def run_julia(name, steps, budget_s):
    """Flou.jl's own multithreaded path (bench/flou_cpu.jl).  None when it cannot run."""
    found = find_julia()
    if not found:
        return None
    exe, proj = found
    w, n = WORKLOADS[name], CPU_SAMPLE[name]
    base = name if name in ("cfg1", "cfg2", "cfg3", "cfg4") else {"cfg3b": "cfg3", "cfg4s": "cfg4"}.get(name)
    if base is None:
        return None
    try:
        out = subprocess.run([exe, "-t", str(host_threads()), f"--project={proj}",
                              os.path.join(ROOT, "bench", "flou_cpu.jl"), base, str(n[0]), str(w["np"]),
                              str(steps)], capture_output=True, text=True, timeout=max(60, budget_s)).stdout
        d = json.loads([l for l in out.splitlines() if l.startswith("{")][-1])
        return d
    except Exception:
        return None


def run_cpu(name, steps, warmup, budget_s=150.0):
    """Times the CPU path on the bounded sample with ALL host threads: Flou.jl itself when a Julia
    is present (kind "reference"), else the oracle (CPU restatement of Flou's rhs! + ORK256 loop,
    OpenMP; kind "port").  Stops early when the time budget is spent; returns a dict with the
    rate (DOF-updates/s per stage), the threads actually used and the steps actually timed."""
    w = WORKLOADS[name]
    info = cpu_sample_info(name)
    jl = run_julia(name, steps, budget_s)
    if jl is not None:
        return {"rate": jl["rate"], "cores": int(jl["threads"]), "kind": "reference", "steps": int(jl["steps"]),
                "warmup": 1, "ms_per_step": jl["ms_per_step"],
                "sample": info["what"] + f", {jl['steps']} ORK256 steps, Flou.jl (julia -t {jl['threads']}, bench/flou_cpu.jl)"}
    import oracle as O
    case, n = cpu_sample_case(name)
    orc = case.oracle()
    start, finish = domain(w)
    # map the unit-box oracle coordinates onto the workload's domain for the IC
    lo, hi = np.zeros(w["nd"]), np.array([1.0 + 0.5 * d for d in range(w["nd"])])
    x = (orc.coords - lo) / (hi - lo) * (np.array(finish) - np.array(start)) + np.array(start)
    Q = np.asfortranarray(initial_condition(x, w))
    # torchrun exports OMP_NUM_THREADS=1 to its workers: set the thread count explicitly and
    # report what OpenMP will really use
    O.set_threads(host_threads())
    cores = O.max_threads()
    t0 = time.perf_counter()
    u = Q
    for _ in range(warmup):
        u = orc.lsrk2n(u, O.ORK256, 1e-5, 1)
    per = (time.perf_counter() - t0) / warmup if warmup > 0 else None
    done, t = 0, time.perf_counter()
    while done < steps:
        u = orc.lsrk2n(u, O.ORK256, 1e-5, 1)
        done += 1
        el = time.perf_counter() - t
        if done < steps and (time.perf_counter() - t0) + el / done > budget_s:
            break
    dt = time.perf_counter() - t
    return {"rate": orc.ndof * 5 * done / dt, "cores": cores, "kind": "port", "steps": done,
            "warmup": warmup, "ms_per_step": dt / done * 1e3,
            "sample": info["what"] + f", {done} ORK256 steps, oracle/pipeline.c (C restatement of the "
                                     f"reference path, OpenMP x{cores}; no Julia on this box)"}


# ----------------------------------------------------------------------------- self-checks
def parity_check(device=0):
    """--check (default on, rank 0 at N=1): BASELINE config 3 at size (32^3, p=3, EC split form +
    matrix dissipation) through the public API against the oracle: RHS and the state after 2
    ORK256 steps.  Returns the dict printed as `parity` (and `parity_err` = the larger error)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import flou_b200 as F
    import oracle as O
    from common import Case, random_state, relerr, smooth_state
    O.set_threads(host_threads())
    case = Case(3, (32, 32, 32), 4, nodes="GLL", eq="euler", op="split", nf="mat", avg="cha")
    orc = case.oracle()
    disc, eq = case.product(device=device)
    Q = np.asfortranarray(0.9 * smooth_state(orc.coords, 3, "euler") + 0.1 * random_state(orc.ndof, 3, "euler", amp=0.3))
    dQ = disc.new_state()
    F.rhs(dQ, Q, F.EquationConfig(disc, eq), 0.0)
    e_rhs = relerr(dQ, orc.rhs(Q))
    nsteps, dt = 2, 2e-5
    ref = orc.lsrk2n(Q, O.ORK256, dt, nsteps)
    u = Q.copy(order="F")
    sol, _ = F.timeintegrate(u, disc, eq, F.ORK256(), nsteps * dt, dt=dt)
    e_state = relerr(sol.u[-1], ref) if sol is not None else float("inf")
    disc.close()
    return {"case": "config 3 at size: 3D Euler 32^3 p=3, SplitDiv(Chandrasekhar)+MatrixDissipation, "
                    "0.9 smooth + 0.1 random state (seed 20230917)",
            "against": "oracle/pipeline.c (CPU restatement of rhs! + ORK256)",
            "rhs_err": e_rhs, "rhs_tol": 1e-12, "state_err": e_state, "state_steps": nsteps,
            "state_tol": 1e-10, "ok": bool(e_rhs <= 1e-12 and e_state <= 1e-10)}


def multi_gpu_check(dist, gloo, rank, world, local_rank, npn):
    """N > 1: a small mesh of the workload's kernel instance, element-partitioned over the N ranks
    with NCCL halo exchange, RHS + 4 RK steps; rank 0 repeats it unpartitioned on its own GPU and
    compares the gathered result BITWISE.  Returns (flag, detail) on rank 0, (None, None) elsewhere."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import flou_b200 as F
    from common import Case, random_state, smooth_state
    case = Case(3, (6, 6, 4 * world), npn, nodes="GLL", eq="euler", op="split", nf="mat", avg="cha")
    disc, eq = case.product(rank=rank, nranks=world, device=local_rank)
    ids = [F.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0, group=gloo)
    disc.comm_init(ids[0])
    full, eq1 = case.product(rank=0, nranks=1, device=local_rank, create=(rank == 0))
    Qg = np.asfortranarray(0.9 * smooth_state(full.coords(), 3, "euler")
                           + 0.1 * random_state(full.coords().shape[0], 3, "euler", amp=0.3))
    Q = np.asfortranarray(Qg[disc.local_rows()])
    dQ = disc.new_state()
    F.rhs(dQ, Q, F.EquationConfig(disc, eq), 0.0)
    u = Q.copy(order="F")
    nsteps, dt = 4, 1e-4
    sol, _ = F.timeintegrate(u, disc, eq, F.ORK256(), nsteps * dt, dt=dt)
    parts = [None] * world
    dist.all_gather_object(parts, (dQ, None if sol is None else sol.u[-1]), group=gloo)
    disc.close()
    if rank != 0:
        return None, None
    dQ1 = full.new_state()
    F.rhs(dQ1, Qg, F.EquationConfig(full, eq1), 0.0)
    u1 = Qg.copy(order="F")
    F.timeintegrate(u1, full, eq1, F.ORK256(), nsteps * dt, dt=dt)
    full.close()
    if any(p[1] is None for p in parts):
        return False, "partitioned run crashed"
    dQn = np.concatenate([p[0] for p in parts], axis=0)
    un = np.concatenate([p[1] for p in parts], axis=0)
    same = bool(np.array_equal(dQn, dQ1) and np.array_equal(un, u1))
    detail = (f"3D Euler {case.n} p={npn - 1} over {world} ranks vs 1 rank: rhs max|d|="
              f"{np.max(np.abs(dQn - dQ1)):.1e}, state after {nsteps} steps max|d|={np.max(np.abs(un - u1)):.1e}")
    return same, detail


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
    ap.add_argument("--no-check", action="store_true",
                    help="skip the self-checks (parity vs the oracle at N=1, multi_gpu_bitwise at N>1)")
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
              # the CPU legs (cpu_baseline here, --impl reference) time this bounded sample of the
              # workload, not the full mesh
              "cpu_sample": cpu_sample_info(args.workload),
              "l2_policy": "inputs larger than L2 (state >> 126 MB)" if ndof_global * 40 > 2.5e8
              else "state fits L2; timed back-to-back as the RK loop runs it"}

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        r = run_cpu(args.workload, max(1, args.steps), max(0, min(args.warmup, 1)))
        rate = r["rate"]
        line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT,
                "n_gpus": args.gpus, "steps": r["steps"], "warmup": r["warmup"],
                "ms_per_step": r["ms_per_step"],
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": rate, "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                                 "sample": r["sample"]},
                "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0},
                "gpu_launches": 0,
                "note": "per-DOF rate of the CPU path on config.cpu_sample (bounded sample of the workload); "
                        "steps = steps actually timed inside the time budget"}
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
               "note": f"the host<->device copies of a call are amortised over {m} RK steps ({m * nstages} stages); "
                       "the upload is not overlapped with the first stage",
               "calls": args.e2e_calls, "pinned_host": pinned, "ms_per_call": te / args.e2e_calls * 1e3}

    ndof_local = disc.ndofs
    launch_info = disc.kernel_info()
    # ---- self-checks (outside every timed region)
    mg_flag = mg_detail = None
    if world > 1 and not args.no_check and w["nd"] == 3 and w["eq"] == "euler" and not w.get("unstructured"):
        disc.close()
        mg_flag, mg_detail = multi_gpu_check(dist, gloo, rank, world, local_rank, w["np"])
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
    stages = nstages * args.steps
    stage_ms = ms_dev / stages                      # whole stage: every kernel of one RK stage
    stage_gbs = ndof_local * bytes_per_dof / (stage_ms * 1e-3) / 1e9
    # dominant kernel = the element kernel (line_kernel_ws: volume + lift + RK update; it moves all
    # of the algorithmic bytes); the face-flux kernel only adds non-algorithmic traffic
    kernel_ms = ksplit["element_kernel_ms"] if ksplit else stage_ms
    achieved = ndof_local * bytes_per_dof / (kernel_ms * 1e-3) / 1e9
    config["launch"] = launch_info
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak,
                # per-launch DRAM bytes from the committed ncu capture of the single-GPU run; a rank of
                # a partitioned run moves 1/N of it plus the halo, not re-captured: null
                "traffic": ncu_traffic(args.workload) if world == 1 else None,
                "kernel": "flou::line_kernel_ws (element kernel of the two-kernel stage)",
                "algorithmic_bytes_per_dof": bytes_per_dof,
                "dofs_per_launch": ndof_local, "avg_launch_ms": kernel_ms,
                "stage_ms": stage_ms, "stage_achieved": stage_gbs, "stage_frac": stage_gbs / peak,
                "kernels_per_stage": ksplit, "peak_source": peak_src,
                "note": "frac = dominant kernel alone; stage_frac = all kernels of an RK stage "
                        "(what `value` is made of)"}

    # second roof: the EC split-form kernel sits near the fp64/HBM ridge
    fpk, cnt = fp64_peak()
    per_dof = (cnt or {}).get(args.workload, {}).get("fp64_thread_instr_per_dof")
    if fpk and per_dof:
        rate = ndof_local / (stage_ms * 1e-3)
        roofline["fp64_frac"] = per_dof["stage"] * rate / fpk["dfma_per_s"]
        roofline["fp64"] = {"thread_instr_per_dof": per_dof, "peak_dfma_per_s": fpk["dfma_per_s"],
                            "peak_source": "measured: profiles/fp64_peak.json (profiles/tools/dfma_peak.cu, "
                                           f"{fpk['dfma_per_clk_per_sm']:.1f} DFMA/clk/SM at {fpk['sm_mhz_during']} MHz)",
                            "note": "fraction of the measured DFMA issue rate the whole stage sustains; the "
                                    "power cap lowers the SM clock under this kernel (see clocks)"}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        r = run_cpu(args.workload, 2, 1, budget_s=60.0)
        cpu = {"value": r["rate"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}

    parity = None
    if world == 1 and not args.no_check:
        disc.close()
        parity = parity_check(local_rank)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu, "wall_ms_per_step": wall / args.steps * 1e3}
    if parity is not None:
        line["parity"] = parity
        line["parity_err"] = max(parity["rhs_err"], parity["state_err"])
    if mg_flag is not None:
        line["multi_gpu_bitwise"] = mg_flag
        line["multi_gpu_check"] = mg_detail
    emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
