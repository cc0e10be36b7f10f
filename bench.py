#!/usr/bin/env python
"""bench.py -- throughput of the per-k-mode Boltzmann path on B200 (BASELINE.json metric:
LCDM C_l sets/sec at lmax 2500, with Boltzmann k-modes/sec alongside).

A "step" is one pass of the hot path (evolve -> spline-in-k -> LOS projection -> C_l) over one
batch of `--batch` synthetic LCDM parameter sets per GPU (planck_2018.ini cosmology box,
massless neutrinos, T/E/lensing-potential sources, lmax 2500: BASELINE.json configs[1] without the
host-side lensing of C_l, which SURVEY.md 8(f) lists as "next").  Parameter sets shard whole
across GPUs with no data-path collective ("scaling": "weak").

  value   device time (CUDA events on the library's stream), boundary tables resident in HBM
  e2e     the same batch through the C-ABI with HOST buffers: per step every cosmology's
          boundary tables are copied host->device from page-locked memory, and LinearSrc-free
          results (iCl) are read back; wall clock bracketed by device synchronisation
  roofline      dominant kernel (the ODE kernel): algorithmic FP64 flops / its CUDA-event time
                against the FP64 FMA peak measured live; the LOS kernel's HBM figure is beside it
  cpu_baseline  the oracle (CPU restatement, OpenMP over k-modes) on this box's host cores

`--impl reference` times the CPU implementation of the path (the oracle port: the reference is
Fortran and cannot be built in this image) on the same workload."""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "lcdm_cl_sets_per_sec_lmax2500"
UNIT = "C_l sets/s"
WORKLOAD = "planck2018_lcdm_box_massless_nu_T+E+phi_lmax2500_keta6250_lensed"


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stop_flag, self.th = index, [], False, None

    def _run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def start(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()

    def stop(self):
        self.stop_flag = True
        if self.th:
            self.th.join(timeout=6)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 6:
                continue
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except ValueError:
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def build_batch(n, seed):
    from camb_b200.host import HostState
    from camb_b200.workloads import synthetic_lcdm_batch
    ps = synthetic_lcdm_batch(n, seed=seed, lmax=2500, lensing=True, massless=True)
    return [HostState(p) for p in ps]


def oracle_cl_sets(hss, nthreads, budget_s, min_sets=1):
    """Time the oracle chain on a bounded sample of the batch; returns (sets, seconds, k_modes)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    t0 = time.perf_counter()
    n = modes = 0
    from camb_b200 import lensing as L
    from camb_b200.workloads import synthetic_highl_template

    class OracleLens:          # the CPU restatement of the theta loops behind the same host code
        def lens_contrib(self, plan, spectra, out=None):
            for i, sp in enumerate(spectra):
                out[i] = O.lens_contrib(plan, sp, nthreads=nthreads)
            return list(out)

    plan, template = L.plan_for(hss[0]), synthetic_highl_template()
    for hs in hss:
        icl, _ = O.cls_from_params(hs, nthreads=nthreads)
        L.lens_cls_from_icl(OracleLens(), plan, [hs], [icl], template, nthreads=1)
        n += 1
        modes += len(hs.source_q())
        if time.perf_counter() - t0 > budget_s and n >= min_sets:
            break
    return n, time.perf_counter() - t0, modes


def run_reference(args, rank, world):
    """--impl reference: the CPU implementation of the path on the host cores (rank 0 only)."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_step = max(1, min(args.batch, 4))
    hss = build_batch(per_step, seed=20261017)
    for hs in hss:
        hs.bessel()
    for _ in range(args.warmup):
        oracle_cl_sets(hss, cores, 1e9)
    t0 = time.perf_counter()
    sets = modes = 0
    for _ in range(args.steps):
        n, _, m = oracle_cl_sets(hss, cores, 1e9)
        sets += n
        modes += m
    dt = time.perf_counter() - t0
    v = sets / dt
    sample = f"{per_step} of the batch's parameter sets per step, OpenMP over k-modes, {cores} threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sets_per_step": per_step,
                   "note": "CPU restatement of the reference path (oracle port); the Fortran reference cannot be built in this image"},
        "k_modes_per_sec": modes / dt,
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def los_gather_evals(hs):
    """Number of (q, l, tau) terms loop B evaluates for one cosmology: N_eval = sum_q sum_{l <= llmax(q)} (n2 - n1 + 1)
    with the limits of DoFlatIntegration (cmbmain.f90:1702-1724).  Each term gathers one 32-byte Bessel record, so
    32 * N_eval is the L1/L2 sector traffic that actually bounds the kernel (SURVEY.md 8(d) 'gather bytes')."""
    m = hs.model
    tau, tau0, ls = np.asarray(hs.tau), m.tau0, np.asarray(hs.ls, dtype=np.float64)
    boost = m.AccuracyBoost * m.BessIntBoost
    total = 0
    for q in np.asarray(hs.q_int):
        llmax = int(round(q * tau0))
        llmax = 17 if llmax < 15 else int(round(q * tau0 + 6 * np.pi))
        l = ls[ls <= llmax]
        if not len(l):
            continue
        tmin = np.maximum(tau[1], tau0 - 80 * l * boost / q)
        tmax = tau0 - np.maximum(0.0, l - 4.2 * l ** (1 / 3.0)) / q
        ok = tmax >= tau[1]
        n1 = np.searchsorted(tau, tmin[ok], side="right")
        n2 = np.searchsorted(tau, tmax[ok], side="right")
        if m.n_src > 2:   # the phi column's DoInt cut removes whole (q, l) pairs at high q
            qmax_int = np.maximum(850.0, l[ok]) * 3 * boost / tau0 * 1.2
            keep = q < qmax_int
            n1, n2 = n1[keep], n2[keep]
        total += int(np.clip(n2 - n1 + 1, 0, None).sum())
    return total


def model_h2d_bytes(hs):
    m = hs.model
    return 8 * (m.nthermo * (16 + 7) + 2 * m.n_tau) + 8 * max(0, m.num_transfer_redshifts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=env_int("CAMB_B200_BATCH", 64), help="parameter sets per GPU per step")
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU work for the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from camb_b200.gpu import GpuContext
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; camb_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = GpuContext(local)
    B = args.batch
    hss = build_batch(B, seed=20261017 + 1000 * rank)       # every rank its own parameter sets
    ctx.set_bessel(hss[0])
    for hs in hss:
        ctx.pin_host_tables(hs)
    dms = [ctx.create_model(hs) for hs in hss]
    n_modes = sum(len(hs.source_q()) for hs in hss)

    # lens_Cls: the batch is projected against one Bessel table (hss[0]'s l sampling), so one plan; the high-L template is
    # synthetic here (the reference's data file does not travel), it only shapes the tail above Max_l
    from camb_b200 import lensing as L
    from camb_b200.workloads import synthetic_highl_template
    plan, template = L.plan_for(hss[0]), synthetic_highl_template()
    host_threads = max(1, min(16, (os.cpu_count() or 1) // world))     # the ranks of one box share its host cores

    def step(want_stats=False):
        _, _, st = ctx.evolve(hss, dms, want_sources=False, want_stats=want_stats)
        t_ev = ctx.timings_ms()["evolve"]
        ctx.project(hss, dms, want_delta=False)
        icl = ctx.cl(hss, dms)
        tm = ctx.timings_ms()
        _, lensed = L.lens_cls_from_icl(ctx, plan, hss, icl, template, ls_owner=hss[0], reuse_buffers=True, nthreads=host_threads)
        tm["lens"] = ctx.last_lens_ms()
        return lensed, st, t_ev, tm

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    fp64_peak = ctx.fp64_peak_tflops()
    for _ in range(args.warmup):
        icl, st, _, _ = step(True)
    # algorithmic work of the ODE kernel from the integrator's own counters
    evals = sum(int(s[:, 0].sum()) for s in st)
    neq_evals = sum(int(s[:, 2].sum()) for s in st)
    attempts = sum(int(s[:, 3].sum()) for s in st)
    neq_attempts = neq_evals / 6.0                       # 6 integrator derivs calls per attempt (FSAL)
    ode_flops = 5.0 * neq_evals + 200.0 * evals + 70.0 * neq_attempts
    launches0 = ctx.launch_count()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    ctx.timer_start()
    t_ev = t_los = t_spl = t_cl = t_lens = 0.0
    for _ in range(args.steps):
        icl, _, tev, tm = step(False)
        t_ev += tev; t_los += tm["los"]; t_spl += tm["spline_k"]; t_cl += tm["cl"]; t_lens += tm["lens"]
    icl = icl.copy()       # the lensed C_l of the resident path (the host buffers are reused by the e2e steps below)
    ms = ctx.timer_stop()
    barrier()
    clocks = sampler.stop()
    launches = ctx.launch_count() - launches0
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * B * args.steps / (ms_max * 1e-3)

    # ---- e2e: host buffers -> C-ABI -> host results, copies inside the timed region ----
    h2d = sum(model_h2d_bytes(hs) for hs in hss) + sum(8 * (3 * len(hs.q_int) + len(hs.source_q())) for hs in hss)
    h2d += B * 8 * 4 * (plan.lmax + 1)
    d2h = sum(8 * len(hs.ls) * 6 for hs in hss) + B * 8 * 4 * (plan.lmax_lensed + 1)

    def e2e_step():
        d = [ctx.create_model(hs) for hs in hss]
        ctx.evolve(hss, d, want_sources=False, want_stats=False)
        ctx.project(hss, d, want_delta=False)
        _, out = L.lens_cls_from_icl(ctx, plan, hss, ctx.cl(hss, d), template, ls_owner=hss[0], reuse_buffers=True, nthreads=host_threads)
        for x in d:
            ctx.destroy_model(x)
        return out

    ctx.set_option("async_model_upload", 1)   # the host tables outlive the step (as CAMBdata's allocatables outlive cmbmain)
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = e2e_step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    assert np.array_equal(out[0], icl[0]) and np.all(icl[:, 2:, 0] > 0), "e2e path and resident path disagree"
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / float(t.item())

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        ev_ms = t_ev / args.steps
        los_ms = t_los / args.steps
        ach = ode_flops / (ev_ms * 1e-3) / 1e12
        # LOS: compulsory bytes per launch = Bessel table + per cosmology (Src + dd + Delta out)
        b = hss[0].bessel()
        los_bytes = 32.0 * (b.nx - 1) * b.n_l + sum(
            8.0 * (2 * len(hs.source_q()) * hs.n_src * hs.n_tau + hs.n_src * len(hs.ls) * len(hs.q_int)) for hs in hss)
        gather_bytes = 32.0 * sum(los_gather_evals(hs) for hs in hss)
        traffic = {}
        try:   # DRAM bytes per launch of the two kernels from the committed `ncu --set full` capture of this command
            traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sets_per_gpu_per_step": B, "k_modes_per_set": n_modes / B,
                       "lensing": "lens_Cls default method (CorrFuncFullSkyApodized), %d angles x l <= %d, lensed output to l = %d; "
                                  "synthetic high-L template" % (plan.c.npoints - 1, plan.lmax, plan.lmax_lensed),
                       "l2": "inputs larger than L2 (per-step working set %.0f MB)" % (
                           (los_bytes + sum(model_h2d_bytes(hs) for hs in hss)) / 1e6)},
            "k_modes_per_sec": world * n_modes * args.steps / (ms_max * 1e-3),
            # BASELINE.json's second metric: Boltzmann k-modes/s of stages 1-2 alone (this rank's evolve stage, CUDA events)
            "k_modes_per_sec_evolve_stage": world * n_modes / (t_ev / args.steps * 1e-3),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "fp64", "kernel": "cb_evolve_kernel", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": ach / fp64_peak if fp64_peak else None, "traffic": traffic.get("cb_evolve_kernel_dram_bytes_per_launch"),
                         "peak_source": "DFMA microbenchmark measured in this run (MEASURED_PEAKS.json has no FP64 entry)",
                         "ms_per_launch": ev_ms, "share_of_step": t_ev / (t_ev + t_los + t_spl + t_cl + t_lens),
                         "derivs_evals_per_launch": evals, "algorithmic_flops_per_launch": ode_flops},
            "roofline_los": {"bound": "hbm", "kernel": "cb_los_kernel", "achieved": los_bytes / (los_ms * 1e-3) / 1e9,
                             "peak": hbm_peak, "unit": "GB/s", "frac": los_bytes / (los_ms * 1e-3) / 1e9 / hbm_peak,
                             "traffic": traffic.get("cb_los_kernel_dram_bytes_per_launch"),
                             "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650",
                             "gather_bytes_per_launch": gather_bytes, "l2_gather_gbs": gather_bytes / (los_ms * 1e-3) / 1e9,
                             "note": "working set is L2-resident: the kernel is bound by the 32-byte Bessel-record gathers through L1/L2 "
                                     "(l2_gather_gbs), not by HBM",
                             "ms_per_launch": los_ms, "share_of_step": t_los / (t_ev + t_los + t_spl + t_cl + t_lens)},
            "stage_ms_per_step": {"evolve": ev_ms, "spline_k": t_spl / args.steps, "los": los_ms, "cl": t_cl / args.steps,
                                  "lens": t_lens / args.steps},
        }
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            n, secs, m = oracle_cl_sets(hss, cores, args.cpu_budget)
            line["cpu_baseline"] = {"value": n / secs, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"first {n} parameter sets of the batch, {secs:.1f} s, OpenMP over k-modes",
                                    "k_modes_per_sec": m / secs}
        print(json.dumps(line))
    for dm in dms:
        ctx.destroy_model(dm)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
