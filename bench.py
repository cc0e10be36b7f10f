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
WORKLOAD = "planck2018_lcdm_box_massless_nu_T+E+phi_lensed_to_lmax2500 (Max_l 2700, k_eta_max 6750: set_for_lmax)"

# --workload: the default is the configuration BASELINE.json's metric is quoted on (configs[1] with its "massless neutrinos"
# override); the others are BASELINE.json's configs[0] and configs[4] as SURVEY.md 8(d) specifies them, and the latency of
# one cosmology.  Every workload is a batch of `--batch` synthetic parameter sets per GPU per step.
WORKLOADS = {
    "lcdm_massless": dict(name=WORKLOAD, lensing=True),
    # the massive-neutrino workloads run the general (shared-memory) evolve kernel, whose per-set time is best at 64 sets per step
    "config1_massive": dict(name="params_ini_box_one_massive_nu_T+E_lmax2500_keta6250_unlensed", lensing=False, batch=64),
    "config5_mnu_w0wa": dict(name="planck2018_box_mnu+w0wa_ppf_T+E_lmax2500_keta6250_unlensed", lensing=False, batch=64),
    "single_cosmology": dict(name="planck_2018_massless_nu_T+E+phi_lmax2500_lensed_one_cosmology", lensing=True),
}


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


def build_batch(n, seed, workload="lcdm_massless"):
    from camb_b200.host import HostState
    from camb_b200.workloads import planck2018_params, synthetic_config1_batch, synthetic_lcdm_batch
    if workload == "config1_massive":
        ps = synthetic_config1_batch(n, seed=seed)
    elif workload == "config5_mnu_w0wa":
        ps = synthetic_lcdm_batch(n, seed=seed, lmax=2500, lensing=False, extended=True)
    elif workload == "single_cosmology":
        ps = [planck2018_params(lmax=2500, lensing=True, massless=True, lensed_to_lmax=True) for _ in range(n)]
    else:
        ps = synthetic_lcdm_batch(n, seed=seed, lmax=2500, lensing=True, massless=True, lensed_to_lmax=True)
    return [HostState(p) for p in ps]


def oracle_cl_sets(hss, nthreads, budget_s, min_sets=1, lensing=True, per_set_threads=False):
    """Time the CPU restatement of the path on a bounded sample of the batch; returns (sets, seconds, k_modes).
    Default: whole parameter sets in parallel, one thread each (a thread pool over ctypes calls, which release the
    GIL) -- the way a batch runs on CPU cores; `per_set_threads`: one set at a time with OpenMP over its k-modes (the
    reference's own single-cosmology parallelisation, cmbmain.f90:205-209)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import concurrent.futures as cf
    import oracle_lib as O
    from camb_b200 import lensing as L
    O.use_fast(True)          # -O3 -march=native build, as the reference's Makefile (fortran/Makefile:103)
    template = L.default_highl_template()
    plan = L.plan_for(hss[0]) if lensing else None
    inner = nthreads if per_set_threads else 1

    class OracleLens:          # the CPU restatement of the theta loops behind the same host code
        def lens_contrib(self, plan, spectra, out=None):
            for i, sp in enumerate(spectra):
                out[i] = O.lens_contrib(plan, sp, nthreads=inner)
            return list(out)

    def one(hs):
        icl, _ = O.cls_from_params(hs, nthreads=inner)
        if lensing:
            L.lens_cls_from_icl(OracleLens(), plan, [hs], [icl], template, nthreads=1)
        return len(hs.source_q())

    t0 = time.perf_counter()
    n = modes = 0
    if per_set_threads:
        for hs in hss:
            modes += one(hs)
            n += 1
            if time.perf_counter() - t0 > budget_s and n >= min_sets:
                break
    else:
        with cf.ThreadPoolExecutor(max_workers=nthreads) as ex:
            i = 0
            while i < len(hss):
                chunk = hss[i:i + nthreads]
                modes += sum(ex.map(one, chunk))
                n += len(chunk)
                i += len(chunk)
                if time.perf_counter() - t0 > budget_s and n >= min_sets:
                    break
    return n, time.perf_counter() - t0, modes


def run_reference(args, rank, world):
    """--impl reference: the CPU implementation of the path on the host cores (rank 0 only)."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    wl = WORKLOADS[args.workload]
    per_step = 1 if args.workload == "single_cosmology" else max(1, min(args.batch, 2 * cores))
    hss = build_batch(per_step, seed=20261017, workload=args.workload)
    for hs in hss:
        hs.bessel()
    single = args.workload == "single_cosmology"
    for _ in range(args.warmup):
        oracle_cl_sets(hss, cores, 1e9, lensing=wl["lensing"], per_set_threads=single)
    t0 = time.perf_counter()
    sets = modes = 0
    for _ in range(args.steps):
        n, _, m = oracle_cl_sets(hss, cores, 1e9, lensing=wl["lensing"], per_set_threads=single)
        sets += n
        modes += m
    dt = time.perf_counter() - t0
    v = sets / dt
    sample = (f"{per_step} of the batch's parameter sets per step, " +
              ("OpenMP over the k-modes of the one cosmology" if single else "one thread per parameter set") +
              f", {cores} threads, oracle built -O3 -march=native")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["name"], "sets_per_step": per_step,
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


def model_h2d_bytes(hs, thermo_on_device=False):
    m = hs.model
    per_knot = 5 if thermo_on_device else 16 + 7   # cs2, dotmu, emmu, adot, a/tau | the Horner records + seven tables
    return 8 * (m.nthermo * per_knot + 2 * m.n_tau) + 8 * max(0, m.num_transfer_redshifts)


def one_cosmology_latency(ctx, reps=5):
    """Wall-clock latency of ONE planck_2018 cosmology (massless neutrinos, lensed) through the C-ABI with host buffers:
    model upload, evolve, project, C_l, lens_Cls.  Best evolve path per stage (the library chooses the kernel by the
    number of modes in the call)."""
    import torch
    from camb_b200 import lensing as L
    from camb_b200.host import HostState
    from camb_b200.workloads import planck2018_params
    hs = HostState(planck2018_params(lmax=2500, lensing=True, massless=True, lensed_to_lmax=True))
    ctx.set_bessel(hs)
    plan, template = L.plan_for(hs), L.default_highl_template()
    out = {}
    for label, opt in (("register_lane_groups", 1), ("mode_per_lane_tmem", 3), ("general_32_lanes", 0)):
        ctx.set_option("use_reg_kernel", opt)
        ctx.set_option("reg_min_modes", 0)          # force the named kernel (the library's default routes few modes to the general one)
        best = 1e30
        for _ in range(reps + 1):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            dm = ctx.create_model(hs)
            ctx.evolve([hs], [dm], want_sources=False, want_stats=False)
            ctx.project([hs], [dm], want_delta=False)
            icl = ctx.cl([hs], [dm])
            L.lens_cls_from_icl(ctx, plan, [hs], icl, template, nthreads=1)
            ctx.destroy_model(dm)
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        out[label] = 1e3 * best
    ctx.set_option("use_reg_kernel", 1)
    ctx.set_option("reg_min_modes", -1)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="lcdm_massless", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=env_int("CAMB_B200_BATCH", 0),
                    help="parameter sets per GPU per step (default: 128, or the workload's own default)")
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU work for the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-latency", action="store_true", help="skip the one-cosmology latency measurement")
    ap.add_argument("--host-thermo-tables", action="store_true",
                    help="e2e leg: upload the host-packed thermo tables (3.3 MB per set) instead of deriving them on the device")
    args = ap.parse_args()
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.warmup < 3:
        args.warmup = 3
    wl = WORKLOADS[args.workload]
    if args.batch <= 0:
        args.batch = wl.get("batch", 128)
    if args.workload == "single_cosmology":
        args.batch = 1
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
        # NCCL prints its version banner with a plain printf when the communicator is created: stdout carries ONE JSON line, so
        # file descriptor 1 points at stderr until the first collective has run
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    ctx = GpuContext(local)
    B = args.batch
    lensing = wl["lensing"]
    hss = build_batch(B, seed=20261017 + 1000 * rank, workload=args.workload)       # every rank its own parameter sets
    ctx.set_bessel(hss[0])
    for hs in hss:
        ctx.pin_host_tables(hs)
    dms = [ctx.create_model(hs) for hs in hss]
    n_modes = sum(len(hs.source_q()) for hs in hss)
    lean = all(hs.model.Num_Nu_massive == 0 and hs.model.de_is_cosmological_constant for hs in hss)

    # Loops B and C run once per group of sets with the same l sampling, each group against its own (resident) Bessel
    # table: every set is projected at ITS l samples, as the reference would (the default batch has two groups, 63 + 1).
    # lens_Cls likewise per group; the high-L template is the reference's own table (package data).
    from camb_b200 import lensing as L
    from camb_b200.gpu import group_by_l_sampling
    groups = group_by_l_sampling(hss)
    for g in groups:
        ctx.set_bessel(hss[g[0]])            # upload every table once, outside the timed region (cosmology independent)
    plans = [L.plan_for(hss[g[0]]) if lensing else None for g in groups]
    plan = plans[0]
    template = L.default_highl_template() if lensing else None
    host_threads = max(1, min(16, (os.cpu_count() or 1) // world))     # the ranks of one box share its host cores

    def project_cl_lens(models, tm, delta_out=None, while_projecting=None):
        """loops B, C and lens_Cls for the whole batch, group by group; returns the per-set result arrays.
        `while_projecting`: host work done after the first group's projection has been queued (it runs under that kernel)."""
        res = [None] * B
        for g, pl in zip(groups, plans):
            hg, dg = [hss[i] for i in g], [models[i] for i in g]
            ctx.set_bessel(hg[0])
            ctx.project(hg, dg, want_delta=delta_out is not None, out=[delta_out[i] for i in g] if delta_out is not None else None)
            if while_projecting is not None:
                while_projecting()
                while_projecting = None
            icl = ctx.cl(hg, dg)          # the first blocking call of the step: every stage timer before it has landed
            t2 = ctx.timings_ms()
            tm["spline_k"] += t2["spline_k"]; tm["los"] += t2["los"]; tm["cl"] += t2["cl"]
            if tm.get("evolve") is None:
                tm["evolve"] = t2["evolve"]
            out = icl
            if lensing:
                _, out = L.lens_cls_from_icl(ctx, pl, hg, icl, template, reuse_buffers=True, nthreads=host_threads)
                tm["lens"] += ctx.last_lens_ms()
            for i, x in zip(g, out):
                res[i] = np.array(x)
        return res

    def step(want_stats=False):
        _, _, st = ctx.evolve(hss, dms, want_sources=False, want_stats=want_stats)
        tm = dict(evolve=None, spline_k=0.0, los=0.0, cl=0.0, lens=0.0)   # cb_evolve / cb_project only queue their work ("async_calls")
        res = project_cl_lens(dms, tm)
        return res, st, tm["evolve"], tm

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    fp64_peak = ctx.fp64_peak_tflops()
    ctx.set_option("async_calls", 1)          # calls without host outputs return when queued; errors surface at the next blocking call
    for _ in range(args.warmup):
        icl, st, _, _ = step(True)
    # algorithmic work of the ODE kernel from the integrator's own counters
    evals = sum(int(s[:, 0].sum()) for s in st)
    neq_evals = sum(int(s[:, 2].sum()) for s in st)
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
    icl = [a.copy() for a in icl]       # the result of the resident path
    ms = ctx.timer_stop()
    barrier()
    clocks = sampler.stop()
    launches = ctx.launch_count() - launches0
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * B * args.steps / (ms_max * 1e-3)

    # ---- e2e: host buffers -> C-ABI -> host results, copies inside the timed region.  What the drop-in API hands back
    # is read back every step: iCl AND Delta_p_l_k (results.py:541 reads it from the host), plus the lensed C_l.
    tod = not args.host_thermo_tables    # e2e: the thermo derivative tables and Horner records are derived on the device (8(f) rank 1)
    h2d = sum(model_h2d_bytes(hs, tod) for hs in hss) + sum(8 * (3 * len(hs.q_int) + len(hs.source_q())) for hs in hss)
    d2h = sum(8 * len(hs.ls) * (6 if hs.n_src > 2 else 3) for hs in hss)
    d2h += sum(8 * hs.n_src * len(hs.ls) * len(hs.q_int) for hs in hss)
    if lensing:
        h2d += B * 8 * 4 * (plan.lmax + 1)
        d2h += B * 8 * 4 * (plan.lmax_lensed + 1)
    # two sets of page-locked result arrays: the device->host copy of step i's Delta_p_l_k (the bulk of the read-back) is queued
    # behind its kernel ("async_outputs") and completes under step i+1's evolve kernels; it is waited for, and step i's
    # device buffers released, before step i+1 projects.  Every step's copies are inside the timed region.
    delta_bufs = [[np.zeros((len(hs.q_int), len(hs.ls), hs.n_src)) for hs in hss] for _ in range(2)]
    for bufs in delta_bufs:
        for a in bufs:
            ctx.pin_array(a)
    pending = []     # device models of the previous step, alive until its copies have landed

    upcoming = []    # device models of the next step: their host tables are handed to the copy engine while this step projects

    def upload_next():
        upcoming[:] = [ctx.create_model(hs, thermo_on_device=tod) for hs in hss]

    def e2e_step(i):
        if not upcoming:
            upload_next()
        d = list(upcoming)
        upcoming.clear()
        ctx.evolve(hss, d, want_sources=False, want_stats=False)
        ctx.wait_outputs()                      # step i-1's Delta_p_l_k is on the host now
        for x in pending:
            ctx.destroy_model(x)
        pending[:] = d
        # every step uploads one batch of inputs inside the timed region: step i+1's, queued behind step i's projection kernel
        return project_cl_lens(d, dict(spline_k=0.0, los=0.0, cl=0.0, lens=0.0), delta_out=delta_bufs[i % 2], while_projecting=upload_next)

    def e2e_drain():
        ctx.wait_outputs()
        for x in pending + upcoming:
            ctx.destroy_model(x)
        pending.clear()
        upcoming.clear()

    ctx.set_option("async_model_upload", 1)   # the host tables outlive the step (as CAMBdata's allocatables outlive cmbmain)
    ctx.set_option("async_outputs", 1)
    for i in range(max(4, args.warmup)):      # both sets of result arrays written, and the device memory pool at its steady size
        e2e_step(i)                           # (three batches of models are alive at the peak), before the clock starts
    e2e_drain()
    barrier()
    t0 = time.perf_counter()
    e2e_marks = [t0]
    for i in range(args.steps):
        out = e2e_step(i)
        e2e_marks.append(time.perf_counter())
    e2e_drain()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    e2e_marks.append(t0 + dt)
    ctx.set_option("async_outputs", 0)
    delta_host = delta_bufs[(args.steps - 1) % 2]
    assert all(np.array_equal(a, b) for a, b in zip(out, icl)) and np.all(np.abs(delta_host[0]).max(axis=(0, 2)) > 0), \
        "e2e path and resident path disagree"
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / float(t.item())

    latency = None
    if rank == 0 and not args.no_latency:
        latency = one_cosmology_latency(ctx)

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
        tot = t_ev + t_los + t_spl + t_cl + t_lens
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["name"], "sets_per_gpu_per_step": B, "k_modes_per_set": n_modes / B,
                       "l_sampling_groups": [len(g) for g in groups],
                       "lensing": ("lens_Cls default method (CorrFuncFullSkyApodized), %d angles x l <= %d, lensed output to l = %d; "
                                   "the reference's high-L template" % (plan.c.npoints - 1, plan.lmax, plan.lmax_lensed)) if lensing else "none",
                       "l2": "inputs larger than L2 (per-step working set %.0f MB)" % (
                           (los_bytes + sum(model_h2d_bytes(hs) for hs in hss)) / 1e6)},
            "k_modes_per_sec": world * n_modes * args.steps / (ms_max * 1e-3),
            # BASELINE.json's second metric: Boltzmann k-modes/s of stages 1-2 alone (this rank's evolve stage, CUDA events)
            "k_modes_per_sec_evolve_stage": world * n_modes / (t_ev / args.steps * 1e-3),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "returns": "iCl, Delta_p_l_k" + (", lensed C_l" if lensing else ""),
                    "thermo_tables": "derived on the device from five base arrays" if tod else "host-packed, uploaded",
                    "step_wall_ms": [round(1e3 * (b - a), 2) for a, b in zip(e2e_marks[:-1], e2e_marks[1:])],   # the last entry is the drain
                    "pipelining": "the device->host copy of a step's Delta_p_l_k overlaps the next step's evolve kernels "
                                  "(cb_set_option async_outputs, cb_wait_outputs); all copies inside the timed region"},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "fp64", "kernel": "cb_evolve_reg_kernel (three static layouts)" if lean else "cb_evolve_kernel",
                         "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": ach / fp64_peak if fp64_peak else None, "traffic": None,
                         "peak_source": "DFMA microbenchmark measured in this run (MEASURED_PEAKS.json has no FP64 entry)",
                         "ms_per_launch": ev_ms, "share_of_step": t_ev / tot,
                         "derivs_evals_per_launch": evals, "algorithmic_flops_per_launch": ode_flops},
            "roofline_los": {"bound": "hbm", "kernel": "cb_los_kernel", "achieved": los_bytes / (los_ms * 1e-3) / 1e9,
                             "peak": hbm_peak, "unit": "GB/s", "frac": los_bytes / (los_ms * 1e-3) / 1e9 / hbm_peak,
                             "traffic": None,
                             "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650",
                             "record_bytes_per_launch": gather_bytes, "l2_gbs": gather_bytes / (los_ms * 1e-3) / 1e9,
                             "note": "the working set is L2-resident: every 32-byte Bessel record comes from L1/L2 (record_bytes_per_launch = "
                                     "32 B x the (q, l, tau) terms of the reference's own limits; l2_gbs = that stream per second); ncu: L2 "
                                     "at 20 % of its throughput, the warps wait on the L1-miss turnaround (DESIGN.md 4.3); HBM is not the resource",
                             "ms_per_launch": los_ms, "share_of_step": t_los / tot},
            "stage_ms_per_step": {"evolve": ev_ms, "spline_k": t_spl / args.steps, "los": los_ms, "cl": t_cl / args.steps,
                                  "lens": t_lens / args.steps},
        }
        if latency:
            line["latency_ms"] = {"one_planck2018_cosmology_1gpu": min(latency.values()), "by_evolve_path": latency,
                                  "note": "wall clock through the C-ABI with host buffers, best of 5"}
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            sample_sets = hss[:max(1, min(B, 2 * cores))]
            for hs in sample_sets:
                hs.bessel()        # cosmology-independent table, cached across calls by the reference too (bessels.f90:83-103): not timed
            oracle_cl_sets(sample_sets[:cores], cores, 1e9, lensing=lensing, per_set_threads=(B == 1))   # one untimed pass (warm caches)
            n, secs, m = oracle_cl_sets(sample_sets, cores, args.cpu_budget, lensing=lensing, per_set_threads=(B == 1))
            line["cpu_baseline"] = {"value": n / secs, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"first {n} parameter sets of the batch, {secs:.1f} s, one thread per parameter set, "
                                              "oracle built -O3 -march=native",
                                    "k_modes_per_sec": m / secs}
            n2, secs2, _ = oracle_cl_sets(hss[:4], cores, min(args.cpu_budget, 6.0), lensing=lensing, per_set_threads=True)
            line["cpu_baseline"]["per_cosmology_openmp_value"] = n2 / secs2   # the reference's own parallelisation: one set at a time
        print(json.dumps(line))
    for dm in dms:
        ctx.destroy_model(dm)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
