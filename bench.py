#!/usr/bin/env python
"""bench.py -- IK steps/s for the BASELINE headline workload (Unitree G1, 3 FrameTasks + PostureTask +
configuration/velocity limits), one solve_ik step over a batch per "step".

  python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun, one rank per GPU)
  python bench.py --impl reference ...                      CPU arm: the fp64 oracle port on all host cores

One JSON line on stdout (rank 0).  Fields follow the driver contract; in addition
  roofline     K1 (FK + Jacobian sweep): algorithmic bytes / CUDA-event time vs measured HBM peak
  roofline_k2  K2 (batched factorisation / active set): algorithmic FLOPs / time vs fp64-FMA peak
  cpu_baseline oracle port timed on this box's host cores on a bounded sample
  e2e          same metric through the host-buffer C-ABI entry (bik_step_host), copies inside the timing
"""

import argparse
import json
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


def host_threads() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def run_reference(args, rank, world):
    """CPU arm: oracle port (oracle/ik_oracle.c, fp64, OpenMP over instances) = the reference's algorithm
    restated in C, because the reference itself (Python + mujoco + qpsolvers) cannot be installed here."""
    if rank != 0:
        return
    from oracle.ikoracle import Oracle, num_threads

    wl = WORKLOADS[args.workload]
    fm = load_flat(wl["robot"])
    spec = spec_from_workload(fm, wl)
    orc = Oracle(fm.to_blob(), spec, fm.nq, fm.nv)
    frames = task_frames(wl, fm)
    B = args.cpu_sample
    inp = make_inputs(fm, wl, B, lambda qq: orc.fk(qq, frames), seed=0)
    threads = host_threads()   # torchrun exports OMP_NUM_THREADS=1: ask for every core we may run on explicitly

    def step():
        return orc.step(inp["q"], inp["frame_targets"], inp["posture_target"], inp.get("com_target"), dt=wl["dt"],
                        damping=wl["damping"], nsteps=1, integrate=True, nthreads=threads)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = B * args.steps / dt
    line = {"metric": METRIC, "value": value, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {len(wl['frames'])} FrameTasks + PostureTask + limits {[l['kind'] for l in wl['limits']]}",
                       "batch_per_step": B, "dt": wl["dt"], "damping": wl["damping"]},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{B} instances x {args.steps} steps, oracle/ik_oracle.c (fp64, OpenMP, Goldfarb-Idnani QP)"},
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
    for _ in range(2):
        hq[:] = hq0
        prob.step_host(hq, hft, hpt, hct, dt=dt_, damping=damping, nsteps=1, integrate=True, out_dq=hdq, out_status=hst)
    n_e2e = max(3, min(args.steps, 10))
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
    return {"value": world * B * n_e2e / te, "unit": UNIT, "h2d_bytes_per_step": up * world, "d2h_bytes_per_step": down * world,
            "note": "bik_step_host (C ABI, pinned host buffers), every rank on its shard at the same time; host wall clock "
                    "per call, max over ranks; bytes are whole-job"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="g1", choices=list(WORKLOADS))
    ap.add_argument("--batch-per-gpu", type=int, default=65536)
    ap.add_argument("--cpu-sample", type=int, default=16384, help="instances per CPU-baseline step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

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

    from mink_b200.engine import DeviceModel, Problem

    wl = WORKLOADS[args.workload]
    fm = load_flat(wl["robot"])
    spec = spec_from_workload(fm, wl)
    model = DeviceModel(fm, device=dev_index)
    prob = Problem(model, spec)
    frames = task_frames(wl, fm)
    B = args.batch_per_gpu

    def fk(qq):
        poses, com = model.fk(qq, frames, want_com=spec.ncom > 0)
        return poses.cpu().numpy().astype(np.float64), (com.cpu().numpy().astype(np.float64) if com is not None else None)

    inp = make_inputs(fm, wl, B, fk, seed=1000 + rank)
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

    # ---- per-kernel timing for the rooflines (same inputs, L2 flushed, CUDA events) ----------------
    k1_ms, k2_ms = [], []
    J, e, ep, Gc, hc = prob.fk_jac(q0, ft, pt, ct, dt=dt_)
    for s in range(max(5, min(args.steps, 10))):
        flush.fill_(1.0)
        a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        torch.cuda._sleep(2_000_000)   # ~1 ms of device spin: the host enqueues both launches behind it, so the events
        a.record()                     # bracket kernel time only (no host launch latency between them)
        J, e, ep, Gc, hc = prob.fk_jac(q0, ft, pt, ct, dt=dt_)
        b.record()
        prob.solve(q0, J, e, ep, Gc, hc, dt_, damping)
        c.record()
        torch.cuda.synchronize()
        k1_ms.append(a.elapsed_time(b)); k2_ms.append(b.elapsed_time(c))
    k1_s, k2_s = statistics.median(k1_ms) * 1e-3, statistics.median(k2_ms) * 1e-3

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

    e2e = measure_e2e(prob, inp, fm, B, dt_, damping, args, world, allmax, barrier)

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
    traffic = None   # DRAM bytes per launch from the committed ncu --set full capture of this workload (never measured under the bench)
    tpath = os.path.join(REPO, "profiles", "r1_traffic.json")
    if os.path.exists(tpath) and B == 65536 and args.workload == "g1":
        tj = json.load(open(tpath))["k1_kernel"]
        traffic = tj["dram_read_bytes"] + tj["dram_write_bytes"]
    roofline = {"kernel": "k1_kernel (FK + Jacobian sweep)", "bound": "hbm", "achieved": achieved, "peak": hbm_peak,
                "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                "bytes_per_instance": bytes_k1 // B, "ms": k1_s * 1e3, "share_of_step": k1_s / (k1_s + k2_s)}
    # K2: algorithmic FLOPs need the mean active-set iteration count -> measured by the oracle on a sample
    mean_iters = None
    cpu = None
    if not args.no_cpu_baseline and world == 1:   # the CPU baseline is an N=1 figure (rank 0 owns every host core there)
        from oracle.ikoracle import Oracle, num_threads

        orc = Oracle(fm.to_blob(), spec, fm.nq, fm.nv)
        Bc = min(args.cpu_sample, B)
        sl = slice(0, Bc)
        ctn = None if inp.get("com_target") is None else inp["com_target"][sl]
        t0 = time.perf_counter()
        dq_ref, _, st_ref, nact = orc.step(inp["q"][sl], inp["frame_targets"][sl], inp["posture_target"], ctn, dt=dt_, damping=damping,
                                           nsteps=1, integrate=True, nthreads=host_threads())
        tc = time.perf_counter() - t0
        cpu = {"value": Bc / tc, "unit": UNIT, "cores": host_threads(), "kind": "port",
               "sample": f"first {Bc} instances of rank 0's batch, 1 step, oracle/ik_oracle.c (fp64, OpenMP)",
               "mean_active_constraints": float(nact.mean())}
        # parity on the same sample (reported, asserted in tests)
        q.copy_(q0)
        one_step()
        torch.cuda.synchronize()
        cpu["max_abs_dq_err_vs_oracle"] = float(np.abs(dq[:Bc].cpu().numpy() - dq_ref).max())
    prec = os.environ.get("BIK_SOLVE_PRECISION", "f64")
    prec = "f32" if prec in ("f32", "float") else "f64"
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    fma_per_clk_sm = 64 if prec == "f64" else 128
    k2_peak = 148 * fma_per_clk_sm * 2 * sm_mhz * 1e6 / 1e12
    mapping = prob.describe(damping)        # which K1 lane group / K2 path this problem runs on (bik_problem_describe)
    k2_name = "K2 " + mapping.split("k2: ")[-1] + " (QP assembly + exact active-set solve, packed Cholesky)"
    _, _, it_dev = prob.solve(q0, J, e, ep, Gc, hc, dt_, damping, return_iters=True)
    mean_iters_dev = float(it_dev.float().mean())
    # k nv (nv+1) + n_iter (nv^3/3 + 2 nv^2) with the MEASURED factorisations per instance.  It is the dense nv x nv
    # count: the kernel does less arithmetic than that (decoupled and unbounded dofs are eliminated), so `frac` is an upper bound
    # on the FMA-pipe share and mostly shows that K2 is latency / issue bound, not FLOP bound.
    flops_k2 = k2_flops_per_instance(fm, spec, mean_iters_dev) * B
    roofline_k2 = {"kernel": k2_name, "bound": "fma", "achieved": flops_k2 / k2_s / 1e12,
                   "peak": k2_peak, "unit": "TFLOP/s", "frac": flops_k2 / k2_s / 1e12 / k2_peak,
                   "peak_source": f"148 SM x {fma_per_clk_sm} FMA/clk x 2 x {sm_mhz:.0f} MHz (nominal CUDA-core rate at the sampled clock)",
                   "flops_per_instance": flops_k2 / B, "flops_note": "dense nv x nv count of SURVEY 8(d) at the measured iteration count",
                   "measured_iterations_mean": mean_iters_dev, "mapping": mapping, "ms": k2_s * 1e3,
                   "share_of_step": k2_s / (k1_s + k2_s)}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * total_s / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 FK/Jacobian + " + prec + " QP", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {len(wl['frames'])} FrameTasks + PostureTask + limits {[l['kind'] for l in wl['limits']]}"
                                   " (BASELINE.json configs[2])",
                       "batch_per_gpu": B, "global_batch": world * B, "dt": dt_, "damping": damping,
                       "step": "check_limits + FK/Jacobian (K1) + QP assemble/solve (K2) + integrate, every step from the same q0",
                       "l2": "256 MB flush between timed steps", "parallelism": f"dp{world} (independent instances, no collective)"},
            "gpu_launches": 2 * args.steps, "clocks": clocks, "roofline": roofline, "roofline_k2": roofline_k2,
            "cpu_baseline": cpu, "e2e": e2e, "rollout_T100": rollout, "wall_s_timed_region": wall}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
