#!/usr/bin/env python
"""bench.py -- IK steps/s for the BASELINE headline workload (Unitree G1, 3 FrameTasks + PostureTask +
configuration/velocity limits; BASELINE.json configs[2]), one solve_ik step over a batch per "step".

  python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun, one rank per GPU)
  python bench.py --impl reference ...                      CPU arm: the fp64 oracle port on the host cores, SAME workload,
                                                            batch and seeds as the GPU arm (rank 0 alone under torchrun)

One JSON line on stdout (rank 0).  Fields follow the driver contract; in addition
  roofline     K1 through the API (bik_fk_jac: FK + dense Jacobian sweep): algorithmic bytes / CUDA-event time vs measured HBM peak
  roofline_k2  K2 (QP assembly + exact active-set solve): algorithmic FLOPs / time vs the MEASURED fp64 FMA rate of the CUDA cores
  step_kernels the two kernels of a step in situ (K1 with the packed hand-off, K2 with the fused integrate)
  cpu_baseline oracle port timed on this box's host cores (N = 1 only)
  e2e          same metric through the host-buffer C-ABI entry (bik_step_host), copies inside the timing
  per_config   BASELINE configs 2, 4, 5 at their batch sizes, and the headline workload in the few-active-bounds regime (N = 1)
  strong       N > 1: BASELINE's literal multi-GPU config (65 536 instances GLOBAL) + the optional all-gather of dq on NCCL
"""

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

from mink_b200._abi import spec_from_workload  # noqa: E402
from mink_b200.workloads import WORKLOADS, load_flat, make_inputs, task_frames  # noqa: E402

METRIC = "ik_steps_per_s"
UNIT = "IK steps/s"
SEED0 = 1000   # rank r of the GPU arm draws its shard with seed SEED0 + r; the reference arm draws the same shards


def k1_bytes_per_instance(fm, spec, posture_batched):
    """reads nq + 7F + P*nq + 3C, writes 6F*nv + 6F + nv*P + C(3nv+3), fp32."""
    F, P, C = spec.nframe, spec.nposture, spec.ncom
    reads = fm.nq + 7 * F + (P * fm.nq if posture_batched else 0) + 3 * C
    writes = 6 * F * fm.nv + 6 * F + P * fm.nv + C * (3 * fm.nv + 3)
    return 4 * (reads + writes)


def k2_flops_per_instance(fm, spec, mean_iters):
    """k*nv*(nv+1) assembly + (n_iter)(nv^3/3 + 2 nv^2) factor+solves."""
    nv, k = fm.nv, spec.nrows
    return k * nv * (nv + 1) + mean_iters * (nv ** 3 / 3.0 + 2.0 * nv * nv)


def bench_config(name, wl, B, world):
    """The `config` object: identical for the GPU arm and the reference arm."""
    tasks = f"{len(wl['frames'])} FrameTasks" + (" + PostureTask" if wl.get("posture") else "") + (" + ComTask" if wl.get("com") else "")
    return {"workload": f"{name}: {tasks} + limits {[l['kind'] for l in wl['limits']]} (BASELINE.json configs[2])" if name == "g1"
            else f"{name}: {tasks} + limits {[l['kind'] for l in wl['limits']]}",
            "batch_per_gpu": B, "global_batch": world * B, "dt": wl["dt"], "damping": wl["damping"],
            "inputs": f"make_inputs(seed={SEED0} + rank), q in the inner 80 % of the joint ranges, targets = FK(q + N(0, 0.1^2)): ~10 active bounds per instance",
            "step": "solve_ik + integrate: check_limits + FK/Jacobians (K1) + QP assemble/solve + integrate (K2), every step from the same q0",
            "l2": "256 MB flush between timed steps", "parallelism": f"dp{world} (independent instances, no collective in the step)"}


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 200 ms while the timed region runs."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def host_cores() -> dict:
    """Cores this process may really use: the scheduler affinity capped by the cgroup CPU quota (a container on a 128-thread
    host with cpu.max = 8 cores runs 128 OpenMP threads 16x slower than 8)."""
    try:
        aff = len(os.sched_getaffinity(0))
    except AttributeError:
        aff = os.cpu_count() or 1
    quota = None
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(path).read().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    quota = float(txt[0]) / float(txt[1])
            else:
                q = float(txt[0])
                if q > 0:
                    quota = q / float(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            break
        except (OSError, ValueError, IndexError):
            continue
    cores = max(1, min(aff, int(math.ceil(quota)) if quota else aff))
    return {"cores": cores, "affinity": aff, "cgroup_quota": quota}


def pin_openmp(cores: int):
    """Before libgomp starts: one thread per usable core, pinned (the oracle library is loaded after this)."""
    os.environ["OMP_NUM_THREADS"] = str(cores)
    os.environ.setdefault("OMP_PROC_BIND", "close")
    os.environ.setdefault("OMP_PLACES", "cores")
    os.environ.setdefault("OMP_DYNAMIC", "false")


def cpu_arm(name, wl, fm, spec, shards, steps, warmup, cores, budget_s):
    """The reference's algorithm on the host cores (oracle/ik_oracle.c: fp64 FK/Jacobians/Lie algebra + Goldfarb-Idnani QP,
    OpenMP over instances) on the given input shards.  Every step solves + integrates the sample from the same q0; the
    sample is the whole batch unless that would exceed `budget_s` for the run.  Returns the cpu_baseline object."""
    from oracle.ikoracle import Oracle

    orc = Oracle(fm.to_blob(), spec, fm.nq, fm.nv)
    cat = lambda k: None if shards[0].get(k) is None else np.concatenate([s[k] for s in shards], axis=0)
    q, ft, ct = cat("q"), cat("frame_targets"), cat("com_target")
    pt = shards[0]["posture_target"]
    total = q.shape[0]

    def run(n):
        t0 = time.perf_counter()
        out = orc.step(q[:n], None if ft is None else ft[:n], pt, None if ct is None else ct[:n], dt=wl["dt"], damping=wl["damping"],
                       nsteps=1, integrate=True, nthreads=cores)
        return time.perf_counter() - t0, out

    probe_n = min(total, 4096)
    run(probe_n)                                   # loads the library, spins the thread team up
    t_probe, _ = run(probe_n)
    rate = probe_n / max(t_probe, 1e-9)
    n = int(min(total, max(probe_n, rate * budget_s / max(steps + warmup, 1))))
    for _ in range(warmup):
        run(n)
    times, out = [], None
    for _ in range(steps):
        t, out = run(n)
        times.append(t)
    med = statistics.median(times)
    dq_ref, _, st_ref, nact = out
    return {"value": n / med, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} of {total} instances (global batch, seeds {SEED0}..{SEED0 + len(shards) - 1}) x {steps} steps, median step time; "
                      "oracle/ik_oracle.c (fp64, OpenMP, threads pinned)",
            "step_s": {"median": med, "min": min(times), "max": max(times)},
            "mean_active_constraints": float(nact.mean()), "flagged": int((st_ref != 0).sum())}, dq_ref, n


def oracle_fk(fm, spec, wl):
    from oracle.ikoracle import Oracle

    orc = Oracle(fm.to_blob(), spec, fm.nq, fm.nv)
    frames = task_frames(wl, fm)
    return lambda qq: orc.fk(qq, frames)


def run_reference(args, rank, world):
    """CPU arm: oracle port (oracle/ik_oracle.c) = the reference's algorithm restated in C, because the reference itself
    (Python + mujoco + qpsolvers) cannot be installed here.  Same workload, batch and seeds as the GPU arm."""
    if rank != 0:
        return
    hc = host_cores()
    pin_openmp(hc["cores"])
    name = args.workload
    wl = WORKLOADS[name]
    fm = load_flat(wl["robot"])
    spec = spec_from_workload(fm, wl)
    B = args.batch_per_gpu
    fk = oracle_fk(fm, spec, wl)
    shards = [make_inputs(fm, wl, B, fk, seed=SEED0 + r) for r in range(world)]
    steps = max(args.steps, 1)
    cpu, _, n = cpu_arm(name, wl, fm, spec, shards, steps, args.warmup, hc["cores"], budget_s=150.0)
    cpu["host"] = hc
    value = cpu["value"]
    line = {"metric": METRIC, "value": value, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * world * B / value, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": bench_config(name, wl, B, world),
            "cpu_baseline": cpu,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def measure_e2e(prob, inp, fm, B, dt_, damping, args, world, allmax, barrier):
    """Same metric through the host-buffer C-ABI entry (bik_step_host): H2D + kernels + D2H inside the timing.
    Inputs live in PINNED host memory; every step copies q + targets up and dq + integrated q + status down.
    Every rank runs its own shard at the same time (they share the host's PCIe/memory system); the job time is
    the max over ranks of the summed host wall clock of the calls (bik_step_host synchronises before returning)."""
    import torch

    def pinned(a, dtype=torch.float32):
        t = torch.empty(a.shape, dtype=dtype, pin_memory=True)
        t.copy_(torch.as_tensor(np.ascontiguousarray(a)).to(dtype))
        return t.numpy()

    hq0 = pinned(inp["q"])
    hq = pinned(inp["q"])
    hft, hpt = pinned(inp["frame_targets"]), pinned(inp["posture_target"])
    hct = None if inp.get("com_target") is None else pinned(inp["com_target"])
    hdq = pinned(np.zeros((B, fm.nv), np.float32))
    hst = pinned(np.zeros(B, np.int32), torch.int32)
    for _ in range(3):
        hq[:] = hq0
        prob.step_host(hq, hft, hpt, hct, dt=dt_, damping=damping, nsteps=1, integrate=True, out_dq=hdq, out_status=hst)
    n_e2e = max(3, min(args.steps, 20))
    te, up, down = 0.0, 0, 0
    barrier()
    for _ in range(n_e2e):
        hq[:] = hq0                       # host-side reset, not timed
        t0 = time.perf_counter()
        _, _, _, up, down = prob.step_host(hq, hft, hpt, hct, dt=dt_, damping=damping, nsteps=1, integrate=True, out_dq=hdq,
                                           out_status=hst)
        te += time.perf_counter() - t0
    assert not hst.any()
    te = allmax(te)
    ms = 1e3 * te / n_e2e
    return {"value": world * B * n_e2e / te, "unit": UNIT, "h2d_bytes_per_step": up * world, "d2h_bytes_per_step": down * world,
            "ms_per_step": ms, "pcie_gbs_up_plus_down_per_gpu": (up + down) / (ms * 1e-3) / 1e9,
            "note": "bik_step_host (C ABI, pinned host buffers; whole-K2-wave chunks through an upload stream, two compute streams in turn and a download stream), every rank on its "
                    "shard at the same time; host wall clock per call, max over ranks; bytes are whole-job"}, hdq.copy()


def timed_steps(torch, prob, q, q0, tensors, dt_, damping, dq, status, flush, n):
    """n solve_ik + integrate steps from q0, L2 flushed before each, CUDA events around each -> list of ms."""
    ft, pt, ct = tensors
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for s in range(n):
        q.copy_(q0)
        flush.fill_(float(s))
        ev[s][0].record()
        prob.step(q, ft, pt, ct, dt=dt_, damping=damping, nsteps=1, integrate=True, dq=dq, status=status)
        ev[s][1].record()
    torch.cuda.synchronize()
    return [a.elapsed_time(b) for a, b in ev]


def config_row(torch, model_cache, name, B, sigma, flush, seed=SEED0, oracle_sample=2048):
    """One per_config row: BASELINE config `name` at batch B -- step time, mapping, K2 iterations, parity on a sample."""
    from mink_b200.engine import DeviceModel, Problem

    wl = WORKLOADS[name]
    fm = load_flat(wl["robot"])
    spec = spec_from_workload(fm, wl)
    dev = flush.device
    if wl["robot"] not in model_cache:
        model_cache[wl["robot"]] = DeviceModel(fm, device=dev.index)
    model = model_cache[wl["robot"]]
    prob = Problem(model, spec)
    frames = task_frames(wl, fm)

    def fk(qq):
        poses, com = model.fk(qq, frames, want_com=spec.ncom > 0)
        return poses.cpu().numpy().astype(np.float64), (com.cpu().numpy().astype(np.float64) if com is not None else None)

    inp = make_inputs(fm, wl, B, fk, seed=seed, sigma=sigma)
    # fp32-representable inputs: the oracle below sees exactly what the device sees (identical inputs, north_star)
    for k in ("q", "frame_targets", "posture_target", "com_target"):
        if inp.get(k) is not None:
            inp[k] = np.asarray(inp[k], np.float32).astype(np.float64)
    f32 = lambda a: None if a is None else torch.tensor(np.ascontiguousarray(a), dtype=torch.float32, device=dev)
    q0, ft, pt, ct = f32(inp["q"]), f32(inp["frame_targets"]), f32(inp["posture_target"]), f32(inp.get("com_target"))
    q = q0.clone()
    dq = torch.empty((B, fm.nv), device=dev, dtype=torch.float32)
    status = torch.empty(B, device=dev, dtype=torch.int32)
    timed_steps(torch, prob, q, q0, (ft, pt, ct), wl["dt"], wl["damping"], dq, status, flush, 3)
    ms = statistics.median(timed_steps(torch, prob, q, q0, (ft, pt, ct), wl["dt"], wl["damping"], dq, status, flush, 10))
    flagged = int(((status & ~1) != 0).sum())
    J, e, ep, Gc, hc = prob.fk_jac(q0, ft, pt, ct, dt=wl["dt"])
    _, _, it = prob.solve(q0, J, e, ep, Gc, hc, wl["dt"], wl["damping"], return_iters=True)
    row = {"batch": B, "value": B / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "sigma": sigma, "mapping": prob.describe(wl["damping"]),
           "k2_factorisations_mean": float(it.float().mean()), "k2_factorisations_max": int(it.max()), "flagged": flagged}
    try:
        from oracle.ikoracle import Oracle

        orc = Oracle(fm.to_blob(), spec, fm.nq, fm.nv)
        n = min(B, oracle_sample)
        ctn = None if inp.get("com_target") is None else inp["com_target"][:n]
        dq_ref, _, st_ref, nact = orc.step(inp["q"][:n], inp["frame_targets"][:n], inp["posture_target"], ctn, dt=wl["dt"], damping=wl["damping"],
                                           nsteps=1, integrate=False, nthreads=host_cores()["cores"])
        ok = st_ref == 0
        row["max_abs_dq_err_vs_oracle"] = float(np.abs(dq[:n].cpu().numpy().astype(np.float64) - dq_ref)[ok].max())
        row["oracle_sample"] = int(ok.sum())
        row["mean_active_constraints"] = float(nact[ok].mean())
    except Exception as exc:   # the oracle is a checker: its absence must not hide the timing
        row["oracle_error"] = f"{type(exc).__name__}: {exc}"
    prob.close()
    return row


def latency_row(torch, model_cache, name, flush, T=200):
    """BASELINE configs[0]: ONE instance (the reference's interactive use) -- per-step latency, the regime where launch overhead
    and not throughput decides.  Device time per step inside one bik_step(nsteps=T) call, and host wall time of one-step calls."""
    from mink_b200.engine import DeviceModel, Problem

    wl = WORKLOADS[name]
    fm = load_flat(wl["robot"])
    spec = spec_from_workload(fm, wl)
    dev = flush.device
    if wl["robot"] not in model_cache:
        model_cache[wl["robot"]] = DeviceModel(fm, device=dev.index)
    model = model_cache[wl["robot"]]
    prob = Problem(model, spec)
    frames = task_frames(wl, fm)

    def fk(qq):
        poses, com = model.fk(qq, frames, want_com=spec.ncom > 0)
        return poses.cpu().numpy().astype(np.float64), (com.cpu().numpy().astype(np.float64) if com is not None else None)

    inp = make_inputs(fm, wl, 1, fk, seed=SEED0)
    f32 = lambda a: None if a is None else torch.tensor(np.ascontiguousarray(a), dtype=torch.float32, device=dev)
    q0, ft, pt, ct = f32(inp["q"]), f32(inp["frame_targets"]), f32(inp["posture_target"]), f32(inp.get("com_target"))
    q = q0.clone()
    dq = torch.empty((1, fm.nv), device=dev, dtype=torch.float32)
    status = torch.empty(1, device=dev, dtype=torch.int32)
    kw = dict(dt=wl["dt"], damping=wl["damping"], integrate=True, dq=dq, status=status)
    dev_ms = []
    for rep in range(5):
        q.copy_(q0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        prob.step(q, ft, pt, ct, nsteps=T, **kw)
        b.record()
        torch.cuda.synchronize()
        dev_ms.append(a.elapsed_time(b))
    walls = []
    for rep in range(300):
        t0 = time.perf_counter()
        prob.step(q, ft, pt, ct, nsteps=1, **kw)
        torch.cuda.synchronize()
        walls.append(time.perf_counter() - t0)
    prob.close()
    return {"batch": 1, "timesteps": T, "us_per_step_device": 1e3 * statistics.median(dev_ms[1:]) / T,
            "us_per_call_host_wall": 1e6 * statistics.median(walls[50:]), "mapping": None,
            "note": "device: T steps inside one bik_step call (2 launches per step, no host round trip); host wall: one-step calls "
                    "through the Python engine including the synchronize"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="g1", choices=list(WORKLOADS))
    ap.add_argument("--batch-per-gpu", type=int, default=65536)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip per_config / rollout / strong-scaling extras")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    hostc = host_cores()
    if world == 1:
        pin_openmp(hostc["cores"])   # the cpu_baseline leg loads the oracle library later; torchrun (N > 1) exports OMP_NUM_THREADS=1

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the product path)")
    # BIK_BENCH_BACKEND=gloo is a plumbing check for boxes with fewer GPUs than ranks (ranks share devices, timing
    # reductions travel over gloo); the driver's runs use NCCL, one GPU per rank.
    backend = os.environ.get("BIK_BENCH_BACKEND", "nccl")
    dev_index = local_rank if backend == "nccl" else local_rank % torch.cuda.device_count()
    torch.cuda.set_device(dev_index)
    dev = torch.device("cuda", dev_index)
    if world > 1:
        if backend == "nccl":
            dist.init_process_group("nccl", device_id=dev)
        else:
            dist.init_process_group(backend)

    def allmax(x: float) -> float:
        """max over ranks of a host scalar (device tensor over NCCL, host tensor over gloo)"""
        if world == 1:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device=dev if backend == "nccl" else "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    from mink_b200 import _lib
    from mink_b200.engine import DeviceModel, Problem

    name = args.workload
    wl = WORKLOADS[name]
    fm = load_flat(wl["robot"])
    spec = spec_from_workload(fm, wl)
    model = DeviceModel(fm, device=dev_index)
    prob = Problem(model, spec)
    frames = task_frames(wl, fm)
    B = args.batch_per_gpu

    def fk(qq):
        poses, com = model.fk(qq, frames, want_com=spec.ncom > 0)
        return poses.cpu().numpy().astype(np.float64), (com.cpu().numpy().astype(np.float64) if com is not None else None)

    inp = make_inputs(fm, wl, B, fk, seed=SEED0 + rank)
    f32 = lambda a: None if a is None else torch.tensor(np.ascontiguousarray(a), dtype=torch.float32, device=dev)
    q0, ft, pt, ct = f32(inp["q"]), f32(inp["frame_targets"]), f32(inp["posture_target"]), f32(inp.get("com_target"))
    q = q0.clone()
    dq = torch.empty((B, fm.nv), device=dev, dtype=torch.float32)
    status = torch.empty(B, device=dev, dtype=torch.int32)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev, dtype=torch.float32)  # 256 MB > 126 MB L2
    dt_, damping = wl["dt"], wl["damping"]

    def one_step():
        prob.step(q, ft, pt, ct, dt=dt_, damping=damping, nsteps=1, integrate=True, dq=dq, status=status)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        q.copy_(q0)
        one_step()
    barrier()
    sampler = ClockSampler(dev_index)
    if rank == 0:
        sampler.start()
    # ---- timed region: exactly K steps, each from the same q0 with L2 flushed in between ----
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    wall0 = time.perf_counter()
    for s in range(args.steps):
        q.copy_(q0)
        flush.fill_(float(s))
        ev[s][0].record()
        one_step()
        ev[s][1].record()
    barrier()
    wall = time.perf_counter() - wall0
    ms = [a.elapsed_time(b) for a, b in ev]
    total_s = allmax(sum(ms)) * 1e-3
    clocks = sampler.stop() if rank == 0 else None
    assert int(status.max()) == 0, "status flags set during the bench"
    value = world * B * args.steps / total_s
    dq_step = dq.clone()

    # ---- per-kernel timing (same inputs, L2 flushed, CUDA events behind a device-side spin: no host latency inside) ----
    def timed(fn, n=8, pre=None):
        out = []
        for _ in range(n + 2):
            if pre:
                pre()
            flush.fill_(1.0)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda._sleep(2_000_000)
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            out.append(a.elapsed_time(b))
        return statistics.median(out[2:])

    J, e, ep, Gc, hc = prob.fk_jac(q0, ft, pt, ct, dt=dt_)
    k1_s = timed(lambda: prob.fk_jac(q0, ft, pt, ct, dt=dt_)) * 1e-3
    k2_s = timed(lambda: prob.solve(q0, J, e, ep, Gc, hc, dt_, damping)) * 1e-3
    # the step's own two kernels (BIK_STEP_PHASE is a measurement aid of libbik: 1 = only K1, 2 = only K2 on the rows K1 left)
    def phase_ms(ph):
        os.environ["BIK_STEP_PHASE"] = str(ph)
        try:
            return timed(one_step, pre=lambda: q.copy_(q0))
        finally:
            os.environ["BIK_STEP_PHASE"] = "0"
    q.copy_(q0); one_step()
    k1_step_ms, k2_step_ms = phase_ms(1), phase_ms(2)

    extras = not args.no_extras
    rollout = None
    if extras:
        # ---- SURVEY 8(d) second regime: T = 100 timesteps with targets held, q integrated on the device --------
        T = 100
        q.copy_(q0)
        prob.step(q, ft, pt, ct, dt=dt_, damping=damping, nsteps=2, integrate=True, dq=dq, status=status)   # warm
        q.copy_(q0)
        flush.fill_(3.0)
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        r0.record()
        prob.step(q, ft, pt, ct, dt=dt_, damping=damping, nsteps=T, integrate=True, dq=dq, status=status)
        r1.record()
        barrier()
        roll_ms = allmax(r0.elapsed_time(r1))
        rollout = {"timesteps": T, "value": world * B * T / (roll_ms * 1e-3), "unit": UNIT,
                   "ms_per_timestep": roll_ms / T,
                   "instances_flagged": int((status != 0).sum().item()),   # status bits OR-ed over the 100 steps (rank 0's shard)
                   "note": "one bik_step call with nsteps=100 from q0, targets held: instances converge, bounds deactivate"}

    e2e, dq_e2e = measure_e2e(prob, inp, fm, B, dt_, damping, args, world, allmax, barrier)
    e2e["max_abs_dq_diff_vs_device_buffers"] = float(np.abs(dq_e2e - dq_step.cpu().numpy()).max())

    # ---- N > 1: BASELINE's literal multi-GPU configuration (65 536 instances GLOBAL) + the optional all-gather of dq ----
    strong = None
    if world > 1 and extras:
        from mink_b200.distributed import all_gather_rows, shard_bounds

        G_total = 65536
        lo, hi = shard_bounds(G_total, rank, world)
        n = hi - lo
        qs0, fts, cts = q0[:n].contiguous(), ft[:n].contiguous(), (None if ct is None else ct[:n].contiguous())
        qs, dqs, sts = qs0.clone(), torch.empty((n, fm.nv), device=dev), torch.empty(n, device=dev, dtype=torch.int32)
        timed_steps(torch, prob, qs, qs0, (fts, pt, cts), dt_, damping, dqs, sts, flush, 3)
        barrier()
        ms_s = timed_steps(torch, prob, qs, qs0, (fts, pt, cts), dt_, damping, dqs, sts, flush, 20)
        step_ms = allmax(statistics.median(ms_s))
        ag = []
        for _ in range(12):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            a.record(); full = all_gather_rows(dqs, G_total); b.record()
            torch.cuda.synchronize()
            ag.append(a.elapsed_time(b))
        ag_ms = allmax(statistics.median(ag[2:]))
        assert full.shape == (G_total, fm.nv)
        strong = {"global_batch": G_total, "batch_per_gpu": n, "ms_per_step": step_ms, "value": G_total / (step_ms * 1e-3), "unit": UNIT,
                  "allgather_dq_ms": ag_ms, "allgather_bytes_total": G_total * fm.nv * 4, "backend": backend,
                  "value_with_allgather_every_step": G_total / ((step_ms + ag_ms) * 1e-3),
                  "note": "BASELINE configs[2] as written: 65 536 G1 instances over N GPUs (strong scaling: median step time, max over ranks); "
                          "all-gather = mink_b200.distributed.all_gather_rows(dq) on NCCL, only needed when a caller wants the whole batch on every rank"}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    peaks_path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        hbm_peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json"
    else:
        hbm_peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    bytes_k1 = k1_bytes_per_instance(fm, spec, posture_batched=False) * B
    achieved = bytes_k1 / k1_s / 1e9
    traffic, traffic_src = None, None   # DRAM bytes per launch from the committed ncu --set full capture of this workload (never measured under the bench)
    tpath = os.path.join(REPO, "profiles", "r2_traffic.json")
    tj = json.load(open(tpath)) if os.path.exists(tpath) else {}
    if B == 65536 and name == "g1" and "k1_dense" in tj:
        traffic = tj["k1_dense"]["dram_read_bytes"] + tj["k1_dense"]["dram_write_bytes"]
        traffic_src = tj.get("source")
    roofline = {"kernel": "k1_kernel<float,4,dense> through bik_fk_jac (FK + dense Jacobian sweep, what Task.compute_jacobian returns)",
                "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic,
                "traffic_source": traffic_src, "peak_source": peak_src,
                "bytes_per_instance": bytes_k1 // B, "ms": k1_s * 1e3, "share_of_api_pair": k1_s / (k1_s + k2_s)}
    cpu = None
    if not args.no_cpu_baseline and world == 1:   # the CPU baseline is an N=1 figure (rank 0 owns every host core there)
        cpu, dq_ref, n_cpu = cpu_arm(name, wl, fm, spec, [inp], 3, 1, hostc["cores"], budget_s=25.0)
        cpu["host"] = hostc
        # parity on the same sample (reported, asserted in tests): the oracle saw the fp64 inputs, the device their fp32 rounding
        cpu["max_abs_dq_err_vs_oracle"] = float(np.abs(dq_step[:n_cpu].cpu().numpy() - dq_ref).max())
    pyref = os.path.join(REPO, "profiles", "r2_python_reference_cpu.json")
    python_reference = json.load(open(pyref)) if os.path.exists(pyref) else None
    prec = os.environ.get("BIK_SOLVE_PRECISION", "f64")
    prec = "f32" if prec in ("f32", "float") else "f64"
    peak64, peak32 = _lib.C.c_double(0), _lib.C.c_double(0)
    _lib.check(_lib.load().bik_measure_fma_peak(dev_index, 1, 5, _lib.C.byref(peak64)))
    _lib.check(_lib.load().bik_measure_fma_peak(dev_index, 0, 5, _lib.C.byref(peak32)))
    k2_peak = peak64.value if prec == "f64" else peak32.value
    mapping = prob.describe(damping)        # which K1 lane group / K2 path this problem runs on (bik_problem_describe)
    k2_name = "K2 " + mapping.split("k2: ")[-1] + " (QP assembly + exact active-set solve, packed Cholesky)"
    _, _, it_dev = prob.solve(q0, J, e, ep, Gc, hc, dt_, damping, return_iters=True)
    mean_iters_dev = float(it_dev.float().mean())
    # k nv (nv+1) + n_iter (nv^3/3 + 2 nv^2) with the MEASURED factorisations per instance.  It is the dense nv x nv
    # count: the kernel does less arithmetic than that (decoupled and unbounded dofs are eliminated), so `frac` is an upper bound
    # on the FMA-pipe share and mostly shows that K2 is latency / issue bound, not FLOP bound.
    flops_k2 = k2_flops_per_instance(fm, spec, mean_iters_dev) * B
    roofline_k2 = {"kernel": k2_name, "bound": "fma", "achieved": flops_k2 / (k2_step_ms * 1e-3) / 1e12,
                   "peak": k2_peak, "unit": "TFLOP/s", "frac": flops_k2 / (k2_step_ms * 1e-3) / 1e12 / k2_peak,
                   "peak_source": f"measured here: bik_measure_fma_peak (16 FMA chains/thread, all SMs): fp64 {peak64.value:.1f}, fp32 {peak32.value:.1f} TFLOP/s",
                   "flops_per_instance": flops_k2 / B, "flops_note": "dense nv x nv count of SURVEY 8(d) at the measured iteration count",
                   "measured_iterations_mean": mean_iters_dev, "mapping": mapping, "ms": k2_step_ms,
                   "ms_through_bik_solve_dense_rows": k2_s * 1e3}
    step_kernels = {"k1_packed_ms": k1_step_ms, "k2_fused_integrate_ms": k2_step_ms, "k2_share_of_step": k2_step_ms / (k1_step_ms + k2_step_ms),
                    "packed_bytes_per_instance": 4 * int(mapping.split("packed=")[1].split(";")[0]),
                    "note": "the two launches of bik_step timed alone (BIK_STEP_PHASE): K1 writes only the non-zero Jacobian columns "
                            "(792 B instead of 3 168 B per G1 instance) and checks the limits; K2 also integrates q"}
    per_config = None
    if extras and world == 1:
        cache = {wl["robot"]: model}
        per_config = {}
        for cname, cB, sg in (("ur5e_dls", 4096, 0.1), ("shadow", 16384, 0.1), ("spot", 32768, 0.1), ("g1_full", 4096, 0.1), ("ur5e_wall", 16384, 0.1),
                               ("h1", 4096, 0.1), ("aloha", 16384, 0.1), ("aloha_coll", 1024, 0.1)):   # beyond BASELINE: more of the reference's examples
            try:
                per_config[cname] = config_row(torch, cache, cname, cB, sg, flush)
            except Exception as exc:
                per_config[cname] = {"error": f"{type(exc).__name__}: {exc}"}
        try:
            per_config["ur5e_single_instance"] = latency_row(torch, cache, "ur5e", flush)
        except Exception as exc:
            per_config["ur5e_single_instance"] = {"error": f"{type(exc).__name__}: {exc}"}
        try:
            per_config["g1_few_active_bounds"] = config_row(torch, cache, "g1", 65536, 0.01, flush)   # delta ~ N(0, 0.01^2)
        except Exception as exc:
            per_config["g1_few_active_bounds"] = {"error": f"{type(exc).__name__}: {exc}"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * total_s / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 FK/Jacobian + " + prec + " QP", "data": "synthetic",
            "config": bench_config(name, wl, B, world),
            "gpu_launches": 2 * args.steps, "clocks": clocks, "roofline": roofline, "roofline_k2": roofline_k2, "step_kernels": step_kernels,
            "cpu_baseline": cpu, "e2e": e2e, "rollout_T100": rollout, "per_config": per_config, "strong": strong,
            "python_reference_on_shims": python_reference, "wall_s_timed_region": wall}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
