#!/usr/bin/env python
"""bench.py — sigma builds/s of the determinant-CI sigma vector build sigma = H C (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cas16] [--impl ours|reference]

A "step" is one CISigmaBuilder.Hamiltonian(C, sigma) on synthetic 8-fold symmetric integrals
(forte2_b200/synthetic.py, variant S, seed 20261017) with C resident in HBM.  Workload at every N:
CAS(16e,16o) singlet, 165 636 900 determinants (C = 1.33 GB, far larger than the 126 MB L2, so no L2
flush is needed between steps).  For N > 1 the CI vector is sharded by alpha-string rows
(strong scaling: the same sigma build is split over N GPUs; forte2_b200/dist.py picks the rows form --
all-to-all of a column slab of C, reduce-scatter of the opposite-spin term -- when the gather-form
kernel serves the space, else the full form with one NCCL all-gather of C per step).

The JSON line carries, beside the driver contract's keys:
  roofline     the opposite-spin kernel (t3g_kernel; t3_kernel where the gather form does not fit) against the FP64
               tensor peak measured in this run with a cuBLAS DGEMM (MEASURED_PEAKS.json has no FP64 entry)
  davidson     (N = 1) time of one Davidson-Liu iteration of a 3-root solve on the same workload: 3 sigma
               builds + the device solver's streaming passes (forte2_b200/davidson.py)
  cpu_baseline the reference's own C++ CISigmaBuilder::Hamiltonian (oracle/_ref) timed on the host cores
               on a bounded sample, scaled to the workload by the reference's own GEMM flop count; the sigma of
               that sample build (CAS(14,14), 11 778 624 determinants: BASELINE.json configs[1]) is compared with
               this package's sigma of the same vector on the GPU: cpu_baseline.parity_rel_l2_vs_gpu
  e2e          the same metric through the numpy-facing call (host buffers in, host buffers out)
--impl reference times that CPU implementation alone, in the same units.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {  # name -> (norb, na, nb)
    "cas12": (12, 6, 6), "cas14": (14, 7, 7), "cas16": (16, 8, 8), "cas18": (18, 9, 9),
}


def comb(n, k):
    from math import comb as c
    return c(n, k)


def hz_gemm_flops(n, na, nb):
    """flops of the reference's Harrison-Zarrabian opposite-spin GEMM, 2 n^4 K1a K1b
    (forte2/ci/ci_sigma_builder_harrison_zarrabian.cc:264-265): the scaling law for the CPU sample."""
    return 2.0 * n ** 4 * comb(n, na - 1) * comb(n, nb - 1)


class quiet_stdout:
    """The reference prints through std::cout even at log level 0 (helpers/logger.h:68-69); keep the
    bench output to the one JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self._saved = os.dup(1)
        self._null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self._null, 1)

    def __exit__(self, *a):
        os.dup2(self._saved, 1)
        os.close(self._null)
        os.close(self._saved)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        self.index, self.samples, self._stop, self._thr = index, [], threading.Event(), None

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._thr = threading.Thread(target=self._run, daemon=True)
        self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._thr.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx.append(float(s[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def measure_fp64_peak(torch, n=8192, reps=5):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    torch.matmul(a, b)
    torch.cuda.synchronize()
    best = float("inf")
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b
    torch.cuda.empty_cache()
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


def reference_sample(workload, budget_s, steps, warmup, threads=None, check=None):
    """Time the reference's own CPU builder (oracle/_ref, else the port) on a bounded sample and scale
    to the workload.  Returns (builds_per_s_scaled, info dict).  check(space, E, H, V, c, sigma_ref), if given,
    is called with the sample's last build and its return value is reported as info["parity_rel_l2_vs_gpu"]."""
    import numpy as np

    from forte2_b200 import synthetic
    from oracle import ref

    n, na, nb = WORKLOADS[workload]
    cores = threads or os.cpu_count() or 1
    os.environ["FORTE_NUM_THREADS_OVERRIDE"] = str(cores)
    os.environ.setdefault("OPENBLAS_NUM_THREADS", str(min(cores, 64)))
    if not ref.available():
        # restatement: scalar Slater-rules loop, one thread, CAS(8,8)
        from oracle import checker, strings_port
        sn, sa, sb = 8, 4, 4
        space = (sa, sb, 0, [[0] * sn], [], [])
        dets = strings_port.determinants(*space)
        E, H, V = synthetic.integrals(sn, "S")
        c = synthetic.trial_vector(len(dets))
        times = []
        for i in range(warmup + steps):
            t = time.perf_counter()
            checker.port_sigma(dets, E, H, V, c)
            if i >= warmup:
                times.append(time.perf_counter() - t)
        per = sum(times) / len(times)
        scale = hz_gemm_flops(n, na, nb) / hz_gemm_flops(sn, sa, sb)
        return 1.0 / (per * scale), {
            "kind": "port", "cores": 1, "seconds_per_sample_build": per,
            "sample": f"oracle/sigma_port.c on CAS({sa + sb},{sn}) ({len(dets)} dets), scaled x{scale:.3g} by the "
                      f"reference's HZ GEMM flop count to {workload}"}
    # probe with CAS(12,12) and pick the largest sample that fits the budget
    def run(sn, reps_warm, reps):
        sa = sb = sn // 2
        s = ref.RefCIStrings(sa, sb, 0, [[0] * sn], [], [])
        E, H, V = synthetic.integrals(sn, "S")
        b = ref.RefCISigmaBuilder(s, E, H, V, 0)
        b.set_algorithm("hz")  # forte2's Python default (forte2/base_classes/params.py:174-176)
        b.set_memory(16384)
        c = synthetic.trial_vector(s.ndet)
        sg = np.zeros_like(c)
        times = []
        for i in range(reps_warm + reps):
            t = time.perf_counter()
            b.Hamiltonian(c, sg)
            if i >= reps_warm:
                times.append(time.perf_counter() - t)
        parity.clear()
        if check is not None and reps > 0:
            try:
                parity.append(check((sa, sb, 0, [[0] * sn], [], []), E, H, V, c, sg))
            except Exception as e:  # the parity figure must never cost the bench line
                parity.append(f"not computed: {e}"[:200])
        return sum(times) / len(times), s.ndet

    parity = []
    probe, _ = run(12, 0, 1)
    est14 = probe * hz_gemm_flops(14, 7, 7) / hz_gemm_flops(12, 6, 6)
    sn = 14 if (n >= 14 and est14 * (steps + warmup) <= budget_s) else min(12, n)
    per, ndet = run(sn, warmup, steps)
    scale = hz_gemm_flops(n, na, nb) / hz_gemm_flops(sn, sn // 2, sn // 2)
    info = {
        "kind": "reference", "cores": cores, "seconds_per_sample_build": per,
        "sample": f"reference CISigmaBuilder::Hamiltonian (hz, oracle/_ref, FORTE_NUM_THREADS_OVERRIDE={cores}, "
                  f"OPENBLAS_NUM_THREADS={os.environ['OPENBLAS_NUM_THREADS']}) on CAS({sn},{sn}) ({ndet} dets), "
                  f"{steps} build(s); scaled x{scale:.4g} by its HZ GEMM flop count 2 n^4 K1a K1b to {workload}"}
    if parity:
        info["parity_rel_l2_vs_gpu"] = parity[0]
        info["parity_note"] = (f"sigma of this package on the GPU against the reference's sigma of the sample build, "
                               f"CAS({sn},{sn}), same integrals and trial vector; the bar is 1e-12")
    return 1.0 / (per * scale), info


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n, na, nb = WORKLOADS[args.workload]
    with quiet_stdout():
        value, info = reference_sample(args.workload, budget_s=240.0, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": "sigma_builds_per_s", "value": value, "unit": "sigma builds/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / value,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.workload, args.gpus),
        "cpu_baseline": {"value": value, "unit": "sigma builds/s", **info},
        "e2e": {"value": value, "unit": "sigma builds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def ncu_traffic(workload, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed
    `ncu --set full` summary of this workload (profiles/, written by scripts/ncu_summary.py), or None."""
    import glob
    import re
    best = None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", f"r*_{kernel}_{workload}.md"))):
        text = open(path).read()
        rd = re.search(r"\| dram__bytes_read\.sum \| ([0-9.]+) \| (\w+) \|", text)
        wr = re.search(r"\| dram__bytes_write\.sum \| ([0-9.]+) \| (\w+) \|", text)
        if rd and wr:
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
            best = {"bytes": float(rd.group(1)) * scale[rd.group(2)] + float(wr.group(1)) * scale[wr.group(2)],
                    "source": os.path.relpath(path, ROOT)}
    return best


def workload_config(workload, gpus):
    n, na, nb = WORKLOADS[workload]
    ndet = comb(n, na) * comb(n, nb)
    return {"workload": f"CASCI({na + nb}e,{n}o) singlet sigma build, {ndet} determinants, synthetic 8-fold symmetric "
                        f"integrals (variant S, seed 20261017)",
            "ndet": ndet, "norb": n, "na": na, "nb": nb,
            "partition": "single GPU" if gpus == 1 else f"alpha-string rows over {gpus} GPUs, NCCL",
            "l2_policy": "inputs (C, sigma: %.2f GB each) larger than L2; no flush" % (ndet * 8 / 1e9)}


def davidson_leg(torch, strings, builder, E, H, V, nroot=3, iters=5):
    """Device Davidson-Liu (forte2_b200/davidson.py) on the bench workload: `iters` iterations of an
    nroot-root solve from the reference's guess (forte2/ci/ci.py:411-452); iteration 0 builds
    guess_per_root * nroot sigma vectors, every later one nroot.  Times are CUDA-event times per iteration."""
    from forte2_b200.ci_solver import CISolver
    from forte2_b200.davidson import DavidsonLiuParams
    solver = CISolver(strings, E, H, V, nroot=nroot, davidson_liu_params=DavidsonLiuParams(maxiter=iters),
                      log_level=0, builder=builder)
    solver.run()
    ds = solver.eigensolver
    steady = [t for t, h in zip(ds.iter_times, ds.history) if h[5] == nroot]
    steady.sort()
    med = steady[len(steady) // 2] if steady else None
    out = {"nroot": nroot, "iterations_run": len(ds.iter_times),
           "iter_ms_all": [round(t * 1e3, 3) for t in ds.iter_times],
           "sigma_builds_per_iter": [h[5] for h in ds.history],
           "iter_ms_steady": med * 1e3 if med else None,
           "solver_launches": int(ds.k.launches),
           "note": "steady = iterations that build nroot sigma vectors; solver overhead = iter - nroot * sigma"}
    del solver, ds
    torch.cuda.empty_cache()
    return out


def run_ours(args):
    import numpy as np
    import torch

    import forte2_b200 as f2
    from forte2_b200 import synthetic

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    n, na, nb = WORKLOADS[args.workload]
    E, H, V = synthetic.integrals(n, "S")
    strings = f2.CIStrings(na, nb, 0, [[0] * n], [], [])
    ndet = strings.ndet
    builder = f2.CISigmaBuilder(strings, E, H, V, 0, device=local_rank)
    builder.set_memory(96 * 1024)
    gen = torch.Generator(device=dev)
    gen.manual_seed(20261018)

    if world == 1:
        c = torch.randn(ndet, dtype=torch.float64, device=dev, generator=gen)
        c /= c.norm()
        sigma = torch.empty_like(c)

        def step():
            builder.Hamiltonian(c, sigma)
    else:
        from forte2_b200.dist import ShardedSigma
        sharded = ShardedSigma(builder, rank, world, form=args.form)
        c_loc = torch.randn(sharded.local_size, dtype=torch.float64, device=dev, generator=gen)
        sigma_loc = torch.empty_like(c_loc)

        def step():
            sharded.Hamiltonian(c_loc, sigma_loc)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    launches_per_step = builder.launch_count()
    t3_ms = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        barrier()
        e0.record()
        for _ in range(args.steps):
            step()
            # per-term CUDA-event times are read back after the loop for the last step only; the
            # library records them on the launching stream every build
        e1.record()
        barrier()
    ms_total = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    last = builder.last_build_ms()  # [H0, same-spin alpha, same-spin beta, opposite spin, total]
    phases = None
    if world > 1:  # every rank takes part: one extra build with a CUDA event at each phase boundary
        phases = [(name, round(ms, 3)) for name, ms in sharded.profile_phases(c_loc, sigma_loc)]
    ms_per_step = ms_total / args.steps
    value = 1e3 / ms_per_step

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (opposite-spin, FP64 tensor pipe) ----
    peak = measure_fp64_peak(torch)
    shape = builder.t3_shape()
    frac_work = 1.0 / world  # each rank contracts its own rows
    useful_flops = 2.0 * shape["M"] * shape["k_cols"] * frac_work
    executed_flops = 2.0 * shape["M_pad"] * shape["k_cols"] * frac_work
    survey_flops = 2.0 * (n * (n + 1) / 2) ** 2 * ndet * frac_work
    t3_s = last[3] * 1e-3
    roofline = {
        "bound": "tensor",
        "kernel": ("t3g_kernel (opposite-spin, gather form, DMMA)"
                   if shape["form"] == "gather" and (world == 1 or sharded.form == "rows")
                   else "t3_kernel (opposite-spin, scatter form, DMMA)"),
        "achieved": useful_flops / t3_s / 1e12, "peak": peak, "unit": "TFLOP/s",
        "frac": useful_flops / t3_s / 1e12 / peak,
        "peak_source": "cuBLAS DGEMM 8192^3 measured in this run (MEASURED_PEAKS.json has no FP64 entry)",
        "flops_per_launch": useful_flops, "executed_flops_per_launch": executed_flops,
        "executed_tflops": executed_flops / t3_s / 1e12,
        "survey_F_sigma3_flops": survey_flops, "survey_F_sigma3_tflops_equiv": survey_flops / t3_s / 1e12,
        "kernel_ms": last[3], "kernel_share_of_step": last[3] / last[4] if last[4] else None,
        "traffic": None,
    }
    if world == 1 and shape["form"] == "gather":
        tr = ncu_traffic(args.workload, "t3g")
        if tr:
            roofline["traffic"] = tr["bytes"]
            roofline["traffic_source"] = tr["source"] + " (ncu --set full, one launch)"
            roofline["algorithmic_hbm_bytes"] = 16.0 * ndet
    hbm_peak = None
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        hbm_peak = 6650.0
    same_spin_ms = last[1] + last[2]
    # flops of the same-spin contraction actually executed: per spin sum_K m_K^2 * columns * 2 with
    # m = C(n - N + 2, 2) parents per (N-2)-electron string (CAS), one-electron part folded in
    def ss_flops(ne, other_strings):
        if ne < 2:
            return 2.0 * n * n * comb(n, ne) * other_strings if ne == 1 else 0.0
        m = comb(n - ne + 2, 2)
        return 2.0 * comb(n, ne - 2) * m * m * other_strings
    ss_fl = (ss_flops(na, comb(n, nb)) + ss_flops(nb, comb(n, na))) * frac_work
    # FP64 reductions into sigma issued by the two same-spin kernels, against the measured L2 atomic rate
    ss_red = sum(comb(n, ne - 2) * comb(n - ne + 2, 2) * o for ne, o in ((na, comb(n, nb)), (nb, comb(n, na)))
                 if ne >= 2) * frac_work
    same_spin = {
        "bound": "hbm", "kernels": "ssg_kernel (alpha) + transposes + ssg_kernel (beta), one-electron folded",
        "algorithmic_bytes": 2 * 24.0 * ndet * frac_work, "ms": same_spin_ms,
        "achieved": (2 * 24.0 * ndet * frac_work) / (same_spin_ms * 1e-3) / 1e9 if same_spin_ms else None,
        "peak": hbm_peak, "unit": "GB/s",
        "frac": ((2 * 24.0 * ndet * frac_work) / (same_spin_ms * 1e-3) / 1e9 / hbm_peak) if same_spin_ms else None,
        "note": "not HBM-bound at this size (DESIGN.md section 3): the binding resources are the FP64 tensor pipe "
                "and the L2 atomic units",
        "tensor_tflops": ss_fl / (same_spin_ms * 1e-3) / 1e12 if same_spin_ms else None,
        "tensor_frac": ss_fl / (same_spin_ms * 1e-3) / 1e12 / peak if same_spin_ms else None,
        "reductions_per_s": ss_red / (same_spin_ms * 1e-3) if same_spin_ms else None,
        "reduction_rate_frac": (ss_red / (same_spin_ms * 1e-3) / 345e9) if same_spin_ms else None,
        "reduction_rate_peak": "345e9 FP64 red.global.add per s, scripts/micro/red_bench.cu (profiles/r01u_red_bench.log)",
    }

    # ---- end to end through the numpy-facing call (host buffers, H2D + D2H inside) ----
    e2e = None
    if world == 1 and args.no_e2e:
        e2e = {"value": None, "unit": "sigma builds/s", "h2d_bytes_per_step": ndet * 8, "d2h_bytes_per_step": ndet * 8,
               "note": "skipped (--no-e2e)"}
    elif world == 1:
        c_host = c.cpu().numpy()
        s_host = np.empty_like(c_host)
        builder.Hamiltonian(c_host, s_host)  # warm-up: allocates the pinned staging buffers
        reps = max(1, min(args.steps, 3))
        t0 = time.perf_counter()
        for _ in range(reps):
            builder.Hamiltonian(c_host, s_host)
        dt = (time.perf_counter() - t0) / reps
        e2e = {"value": 1.0 / dt, "unit": "sigma builds/s", "h2d_bytes_per_step": ndet * 8,
               "d2h_bytes_per_step": ndet * 8, "ms_per_step": dt * 1e3, "steps": reps,
               "api": "CISigmaBuilder.Hamiltonian(numpy, numpy) -> f2_sigma_apply_host"}
    else:
        e2e = {"value": value, "unit": "sigma builds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
               "note": "sharded vectors live on the GPUs; no host path at N>1"}
        comm_ms = ms_per_step - last[4]

    # ---- one Davidson-Liu iteration of a 3-root solve (N = 1): the caller on either side of sigma ----
    davidson = None
    if world == 1 and not args.no_davidson:
        try:
            davidson = davidson_leg(torch, strings, builder, E, H, V)
        except torch.cuda.OutOfMemoryError as e:  # the 3-root subspace needs 27 vectors
            davidson = {"error": f"out of memory: {e}"[:200]}

    # ---- CPU baseline beside it (rank 0, N = 1 only) ----
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        def gpu_parity(space, Es, Hs, Vs, cs, sigma_ref):
            # BASELINE.json configs[1] when the sample is CAS(14,14): the reference's own sigma vs ours, full size
            st = f2.CIStrings(*space)
            bs = f2.CISigmaBuilder(st, Es, Hs, Vs, 0, device=local_rank)
            got = np.empty_like(cs)
            bs.Hamiltonian(cs, got)
            return float(np.linalg.norm(got - sigma_ref) / np.linalg.norm(sigma_ref))

        with quiet_stdout():
            v, info = reference_sample(args.workload, budget_s=30.0, steps=1, warmup=0, check=gpu_parity)
        cpu = {"value": v, "unit": "sigma builds/s", **info}

    line = {
        "metric": "sigma_builds_per_s", "value": value, "unit": "sigma builds/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.workload, world),
        "term_ms_last_step": {"H0": last[0], "same_spin_alpha": last[1], "same_spin_beta": last[2],
                              "opposite_spin": last[3], "total": last[4]},
        "roofline": roofline, "roofline_same_spin": same_spin, "cpu_baseline": cpu, "e2e": e2e,
        "davidson": davidson,
        "sharded": None if world == 1 else {
            "form": sharded.form, "local_compute_ms_rank0": last[4], "exchange_ms_rank0": comm_ms,
            "phases_ms_rank0_one_extra_build": phases,
            "note": "exchange = step time minus rank 0's kernel time: the NCCL collectives of "
                    "forte2_b200/dist.py and their pack/unpack copies"},
        "gpu_launches": int(launches_per_step * args.steps), "clocks": clocks.summary(),
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="cas16", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-davidson", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (needs 2 x 8 ndet bytes of pinned memory)")
    ap.add_argument("--form", default="auto", choices=["auto", "full", "rows"],
                    help="N > 1: communication form of forte2_b200/dist.py")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
