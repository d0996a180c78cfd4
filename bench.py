#!/usr/bin/env python
"""bench.py — sigma builds/s of the determinant-CI sigma vector build sigma = H C (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cas16] [--impl ours|reference] [--legs a,b]

A "step" is one CISigmaBuilder.Hamiltonian(C, sigma) on synthetic 8-fold symmetric integrals
(forte2_b200/synthetic.py, variant S, seed 20261017) with C resident in HBM.  Headline workload at every N:
CAS(16e,16o) singlet, 165 636 900 determinants (BASELINE.json configs[2]; C = 1.33 GB, far larger than the 126 MB
L2, so no L2 flush is needed between steps).  For N > 1 the CI vector is sharded by alpha-string rows
(strong scaling: the same sigma build is split over N GPUs; forte2_b200/dist.py picks the rows form --
all-to-all of a column slab of C, reduce-scatter of the opposite-spin term -- when a gather-form
kernel serves the space, else the full form with one NCCL all-gather of C per step).

The JSON line carries, beside the driver contract's keys:
  parity_sampled_rel  EVERY run: relative L2 error of the sigma just timed at ~2000 seeded determinants against
               sigma_I = sum_J <I|H|J> C_J rebuilt on the host from Slater rules (oracle/sampled.py; sharded: the
               coefficients are fetched from the owning ranks)
  roofline     the opposite-spin kernel against the FP64 tensor peak measured in this run with a cuBLAS DGEMM
               (MEASURED_PEAKS.json has no FP64 entry)
  default_memory  (N = 1) the same step under set_memory(1024), forte2's default ci_builder_memory
               (forte2/base_classes/params.py:149-178); the headline figure uses a budget that holds the
               transposed beta scratch in one slab
  rdm          (N = 1) device time of the 1- and 2-RDM spin blocks of one vector, tensor roofline of the opposite-spin block
  davidson     (N = 1) one Davidson-Liu iteration of a 3-root solve on the same workload: 3 sigma builds + the
               device solver's streaming passes, with the HBM roofline of those passes (davidson.roofline)
  legs         (N = 8, or --legs) BASELINE.json configs[3] and [4] -- CAS(18e,18o), 2 363 904 400 determinants, and
               the 3-GAS / 20-orbital space, 1 865 504 274 determinants in 225 blocks -- sharded over the GPUs: ms per
               sigma build (max over ranks), per-kernel roofline, sampled parity, ms per Davidson iteration
  cpu_baseline the reference's own C++ CISigmaBuilder::Hamiltonian (oracle/_ref) on the host cores: the full
               CAS(16,16) build measured by `--impl reference` on this box when that arm ran first (build/ref_arm_*.json),
               else a bounded CAS(14,14) sample scaled by the reference's GEMM flop count and labelled extrapolated;
               in both cases the sigma of the CAS(14,14) sample build (BASELINE.json configs[1]) is compared with this
               package's sigma of the same vector on the GPU: cpu_baseline.parity_rel_l2_vs_gpu
  e2e          the same metric through the numpy-facing call (host buffers in, host buffers out)
--impl reference times ONE full build of the workload by that CPU implementation (all host threads), whatever
--steps says (steps_measured = 1): a CAS(16,16) build takes minutes; only if a probe predicts more than 25 minutes
does it fall back to the largest smaller CAS and say so (config.workload names the size actually timed).
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

def _cas(n):
    return (n // 2, n // 2, 0, [[0] * n], [], [])


# name -> CIStrings arguments (na, nb, symmetry, orbital_symmetry, gas_min, gas_max).  gas20 is BASELINE.json
# configs[4] as made concrete in SURVEY.md section 8d; gas20twin its one-GPU validation size.
SPACES = {"cas12": _cas(12), "cas14": _cas(14), "cas16": _cas(16), "cas18": _cas(18),
          "gas20": (10, 10, 0, [[0] * 6, [0] * 8, [0] * 6], [8, 0, 0], [12, 16, 4]),
          "gas20twin": (10, 10, 0, [[0] * 6, [0] * 8, [0] * 6], [10, 0, 0], [12, 16, 2])}
WORKLOADS = {k: (sum(len(g) for g in v[3]), v[0], v[1]) for k, v in SPACES.items()}  # name -> (norb, na, nb)
NDET = {"gas20": 1865504274, "gas20twin": 22065484}
LABEL = {"gas20": "GAS/ORMAS CI, 3 GAS spaces (6, 8, 6 orbitals), 10 alpha + 10 beta electrons, gas_min [8,0,0], "
                  "gas_max [12,16,4], 225 determinant blocks",
         "gas20twin": "GAS/ORMAS CI, 3 GAS spaces (6, 8, 6 orbitals), gas_min [10,0,0], gas_max [12,16,2], 36 blocks"}
REF_CACHE = os.path.join(ROOT, "build")


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


def _ref_threads(threads=None):
    cores = threads or os.cpu_count() or 1
    os.environ["FORTE_NUM_THREADS_OVERRIDE"] = str(cores)
    os.environ.setdefault("OPENBLAS_NUM_THREADS", str(min(cores, 64)))
    return cores


def _ref_build_time(sn, reps_warm, reps, check=None):
    """seconds per CISigmaBuilder::Hamiltonian of the reference (oracle/_ref, hz) on CAS(sn, sn), mean of `reps`;
    returns (seconds, ndet, parity or None)"""
    import numpy as np

    from forte2_b200 import synthetic
    from oracle import ref
    sa = sb = sn // 2
    space = (sa, sb, 0, [[0] * sn], [], [])
    s = ref.RefCIStrings(*space)
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
    parity = None
    if check is not None:
        try:
            parity = check(space, E, H, V, c, sg)
        except Exception as e:  # the parity figure must never cost the bench line
            parity = f"not computed: {e}"[:200]
    return sum(times) / len(times), s.ndet, parity


def _port_sample(workload, steps, warmup):
    """oracle/_ref absent: the plain-C restatement (scalar Slater-rules loop, one thread) on CAS(8,8), extrapolated"""
    from forte2_b200 import synthetic
    from oracle import checker, strings_port
    n, na, nb = WORKLOADS[workload]
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
        "kind": "port", "cores": 1, "seconds_per_sample_build": per, "extrapolated": True,
        "sample": f"oracle/sigma_port.c on CAS({sa + sb},{sn}) ({len(dets)} dets), scaled x{scale:.3g} by the "
                  f"reference's HZ GEMM flop count to {workload}"}


def _ref_cache_path(workload):
    return os.path.join(REF_CACHE, f"ref_arm_{workload}.json")


def reference_full(workload, limit_s=1500.0):
    """The --impl reference arm: ONE full build of the workload by the reference's CPU implementation.  A CAS(12,12)
    probe scaled by the HZ GEMM flop count (times 1.7: the law under-predicts CAS(12) -> CAS(14) by that much)
    decides whether the full build fits `limit_s`; if not, the largest CAS that does is timed instead and the
    workload figure is an extrapolation, labelled so.  Returns (builds/s, info, workload actually timed)."""
    from oracle import ref
    if workload not in ("cas12", "cas14", "cas16", "cas18") or not ref.available():
        v, info = _port_sample(workload, 1, 0) if not ref.available() else (None, None)
        if v is None:
            raise SystemExit("the reference arm times CAS workloads only")
        return v, info, "cas8"
    n, na, nb = WORKLOADS[workload]
    cores = _ref_threads()
    probe, _, _ = _ref_build_time(12, 0, 1)
    sizes = [m for m in (18, 16, 14, 12) if m <= n]
    timed = 12
    for m in sizes:
        est = probe * hz_gemm_flops(m, m // 2, m // 2) / hz_gemm_flops(12, 6, 6) * (1.7 if m > 12 else 1.0)
        if est <= limit_s and m <= 16:  # CAS(18,18): C + sigma + TR/TL alone are 76 GB and hours of CPU time
            timed = m
            break
    per, ndet, _ = (probe, comb(12, 6) ** 2, None) if timed == 12 else _ref_build_time(timed, 0, 1)
    scale = hz_gemm_flops(n, na, nb) / hz_gemm_flops(timed, timed // 2, timed // 2)
    info = {
        "kind": "reference", "cores": cores, "seconds_per_sample_build": per, "steps_measured": 1,
        "extrapolated": timed != n,
        "sample": f"reference CISigmaBuilder::Hamiltonian (hz, oracle/_ref, FORTE_NUM_THREADS_OVERRIDE={cores}, "
                  f"OPENBLAS_NUM_THREADS={os.environ['OPENBLAS_NUM_THREADS']}, set_memory 16 GiB) on CAS({timed},{timed}) "
                  f"({ndet} dets), ONE full build" + ("" if timed == n else
                  f"; scaled x{scale:.4g} by its HZ GEMM flop count 2 n^4 K1a K1b to {workload} (EXTRAPOLATED)"),
        "probe_cas12_seconds": probe}
    if timed == n:
        try:
            os.makedirs(REF_CACHE, exist_ok=True)
            with open(_ref_cache_path(workload), "w") as f:
                json.dump({"workload": workload, "seconds": per, "cores": cores, "when": time.time(),
                           "host": os.uname().nodename}, f)
        except OSError:
            pass
    return 1.0 / (per * scale), info, f"cas{timed}"


def cpu_baseline_sample(workload, check=None):
    """cpu_baseline of this package's arm: a bounded sample of the reference's builder (one CAS(14,14) build, about
    15 s on 16 threads; its sigma is what `check` compares with the GPU), and, when `--impl reference` measured the
    full workload on this box just before (the driver runs that arm first), that measurement as the value."""
    from oracle import ref
    if not ref.available():
        return _port_sample(workload, 1, 0)
    n, na, nb = WORKLOADS[workload]
    cores = _ref_threads()
    sn = min(14, n)
    per, ndet, parity = _ref_build_time(sn, 0, 1, check)
    scale = hz_gemm_flops(n, na, nb) / hz_gemm_flops(sn, sn // 2, sn // 2)
    info = {"kind": "reference", "cores": cores, "seconds_per_sample_build": per,
            "sample": f"reference CISigmaBuilder::Hamiltonian (hz, oracle/_ref, FORTE_NUM_THREADS_OVERRIDE={cores}, "
                      f"OPENBLAS_NUM_THREADS={os.environ['OPENBLAS_NUM_THREADS']}) on CAS({sn},{sn}) ({ndet} dets), 1 build"}
    if parity is not None:
        info["parity_rel_l2_vs_gpu"] = parity
        info["parity_note"] = (f"sigma of this package on the GPU against the reference's sigma of the sample build, "
                               f"CAS({sn},{sn}), same integrals and trial vector; the bar is 1e-12")
    value = 1.0 / (per * scale)
    info["extrapolated"] = sn != n
    try:
        cached = json.load(open(_ref_cache_path(workload)))
        if cached["cores"] == cores and cached["host"] == os.uname().nodename and time.time() - cached["when"] < 6 * 3600:
            value = 1.0 / cached["seconds"]
            info["extrapolated"] = False
            info["seconds_per_workload_build_measured"] = cached["seconds"]
            info["sample"] += (f"; value = ONE full {workload} build measured by `bench.py --impl reference` on this box "
                               f"{time.time() - cached['when']:.0f} s earlier ({cached['seconds']:.1f} s per build)")
            info["flop_law_prediction_seconds"] = per * scale
    except (OSError, KeyError, ValueError):
        pass
    if info["extrapolated"]:
        info["sample"] += (f"; value scaled x{scale:.4g} by the HZ GEMM flop count 2 n^4 K1a K1b to {workload} "
                           f"(EXTRAPOLATED: no full-size measurement on this box)")
    return value, info


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    with quiet_stdout():
        value, info, timed = reference_full(args.workload)
    cfg = workload_config(args.workload, args.gpus)
    if timed != args.workload:
        cfg["workload"] += f" -- TIMED: {timed}, extrapolated to this size"
    line = {
        "impl": "reference", "metric": "sigma_builds_per_s", "value": value, "unit": "sigma builds/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "steps_measured": 1,
        "ms_per_step": 1e3 / value,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
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
    ndet = NDET.get(workload, comb(n, na) * comb(n, nb))
    what = LABEL.get(workload, f"CASCI({na + nb}e,{n}o) singlet")
    return {"workload": f"{what} sigma build, {ndet} determinants, synthetic 8-fold symmetric "
                        f"integrals (variant S, seed 20261017)",
            "ndet": ndet, "norb": n, "na": na, "nb": nb,
            "partition": "single GPU" if gpus == 1 else f"alpha-string rows over {gpus} GPUs, NCCL",
            "l2_policy": "inputs (C, sigma: %.2f GB each) larger than L2; no flush" % (ndet * 8 / 1e9)}


def hbm_peak_gbs():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "MEASURED_PEAKS.json"
    except Exception:
        return 6650.0, "fallback stated in B200_PROFILING.md (MEASURED_PEAKS.json absent)"


def davidson_leg(torch, strings, builder, E, H, V, sigma_ms, nroot=3, iters=6):
    """Device Davidson-Liu (forte2_b200/davidson.py) on the bench workload: `iters` iterations of an
    nroot-root solve from the reference's guess (forte2/ci/ci.py:411-452); iteration 0 builds
    guess_per_root * nroot sigma vectors, every later one nroot.  Times are CUDA-event times per iteration.
    roofline: the solver's own streaming passes (everything in an iteration that is not a sigma build) against HBM:
    bytes = 8 x (vector rows read + written by the f2_dl_* kernels), counted call by call."""
    from forte2_b200.ci_solver import CISolver
    from forte2_b200.davidson import DavidsonLiuParams
    solver = CISolver(strings, E, H, V, nroot=nroot, davidson_liu_params=DavidsonLiuParams(maxiter=iters),
                      log_level=0, builder=builder)
    solver.run()
    ds = solver.eigensolver
    steady = sorted((t, b) for t, b, h in zip(ds.iter_times, ds.iter_bytes, ds.history) if h[5] == nroot)
    med, med_bytes = steady[len(steady) // 2] if steady else (None, None)
    peak, src = hbm_peak_gbs()
    out = {"nroot": nroot, "iterations_run": len(ds.iter_times),
           "iter_ms_all": [round(t * 1e3, 3) for t in ds.iter_times],
           "sigma_builds_per_iter": [h[5] for h in ds.history],
           "iter_ms_steady": med * 1e3 if med else None,
           "solver_launches": int(ds.k.launches),
           "note": "steady = iterations that build nroot sigma vectors; solver overhead = iter - nroot * sigma"}
    if med:
        solver_ms = med * 1e3 - nroot * sigma_ms
        out["solver_ms_per_iter"] = solver_ms
        out["roofline"] = {"bound": "hbm", "kernels": "f2_dl_* streaming passes of one iteration (subspace matrix, "
                           "residual + preconditioner + first projection, block Gram-Schmidt, collapse when due)",
                           "bytes_per_iter": med_bytes, "ms": solver_ms,
                           "achieved": med_bytes / (solver_ms * 1e-3) / 1e9 if solver_ms > 0 else None,
                           "peak": peak, "peak_source": src, "unit": "GB/s",
                           "frac": med_bytes / (solver_ms * 1e-3) / 1e9 / peak if solver_ms > 0 else None}
    del solver, ds
    torch.cuda.empty_cache()
    return out


def rdm_leg(torch, strings, builder, peak_tflops, memory_mb=1024):
    """1- and 2-RDMs of one vector of the bench workload (what an MCSCF macro-iteration asks for after the solve:
    forte2/base_classes/ci_base.py:104-133): device time of every spin block through f2_sigma_rdm (CUDA events, result
    left on the device), and the tensor roofline of the dominant one (ab_2rdm, rdm_ab_kernel: 2 n^2 K_a nbs^2 nas ... flops
    executed = 2 x n^2 x (single replacements per alpha string) per determinant)."""
    from forte2_b200 import _lib, synthetic
    L = _lib.lib()
    n, na = strings.norb, strings.na
    stream = torch.cuda.current_stream().cuda_stream
    builder.set_memory(memory_mb)
    c = torch.from_numpy(synthetic.trial_vector(strings.ndet)).cuda()
    kinds = {"a_1rdm": (0, n), "b_1rdm": (1, n), "aa_2rdm": (3, n * (n - 1) // 2), "bb_2rdm": (4, n * (n - 1) // 2),
             "ab_2rdm": (5, n * n)}
    ms, traces = {}, {}
    for name, (kind, M) in kinds.items():
        out = torch.empty(M * M, dtype=torch.float64, device="cuda")
        _lib.check(L.f2_sigma_rdm(builder._h, kind, c.data_ptr(), c.data_ptr(), out.data_ptr(), stream))  # warm-up, tables
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(L.f2_sigma_rdm(builder._h, kind, c.data_ptr(), c.data_ptr(), out.data_ptr(), stream))
        e1.record()
        torch.cuda.synchronize()
        ms[name] = e0.elapsed_time(e1)
        traces[name] = float(out.view(M, M).diagonal().sum())
    nb = strings.nb
    want = {"a_1rdm": na, "b_1rdm": nb, "aa_2rdm": na * (na - 1) / 2, "bb_2rdm": nb * (nb - 1) / 2, "ab_2rdm": na * nb}
    norm2 = float(c @ c)
    trace_err = max(abs(traces[k] / norm2 - want[k]) / max(want[k], 1) for k in kinds)
    shape = builder.t3_shape()
    ab_flops = 2.0 * n * n * shape["k_cols"]  # n^2 (v y) rows x every (single replacement of Ja, beta string) pair
    ab_tf = ab_flops / (ms["ab_2rdm"] * 1e-3) / 1e12
    total = sum(ms.values())
    del c
    torch.cuda.empty_cache()
    return {"ms": {k: round(v, 3) for k, v in ms.items()}, "sf_1rdm_plus_sf_2rdm_ms": round(total, 3),
            "set_memory_mb": memory_mb, "trace_rel_err": trace_err,
            "roofline": {"bound": "tensor", "kernel": "rdm_ab_kernel (opposite-spin 2-RDM, gather form, DMMA)",
                         "flops_per_launch": ab_flops, "ms": ms["ab_2rdm"], "achieved": ab_tf, "peak": peak_tflops,
                         "unit": "TFLOP/s", "frac": ab_tf / peak_tflops if peak_tflops else None},
            "note": "traces: tr(1-RDM) = N_spin, tr(aa) = N(N-1)/2, sum_uv ab[u,v,u,v] = Na Nb, relative to <c|c>"}


class Ctx:
    pass


def measure_space(ctx, name, steps, warmup, memory_mb, form="auto", nsamp=2000, parity=True, clocks=False):
    """One sigma-build leg: set up the space, time `steps` builds after `warmup` (CUDA events, barrier and
    synchronize on both sides, max over ranks), read the per-term kernel times of the last build, check the last
    sigma at sampled determinants.  Returns (result dict, live objects)."""
    import forte2_b200 as f2
    from forte2_b200 import synthetic
    torch, dist, rank, world, dev = ctx.torch, ctx.dist, ctx.rank, ctx.world, ctx.dev
    space = SPACES[name]
    n, na, nb = WORKLOADS[name]
    t0 = time.perf_counter()
    E, H, V = synthetic.integrals(n, "S")
    strings = f2.CIStrings(*space)
    t_strings = time.perf_counter() - t0
    builder = f2.CISigmaBuilder(strings, E, H, V, 0, device=ctx.local_rank)
    builder.set_memory(memory_mb)
    ndet = strings.ndet
    gen = torch.Generator(device=dev)
    gen.manual_seed(20261018 + rank)
    sharded = None
    if world == 1:
        c = torch.randn(ndet, dtype=torch.float64, device=dev, generator=gen)
        c /= c.norm()
        sigma = torch.empty_like(c)

        def step():
            builder.Hamiltonian(c, sigma)
    else:
        from forte2_b200.dist import ShardedSigma
        sharded = ShardedSigma(builder, rank, world, form=form)
        c = torch.randn(sharded.local_size, dtype=torch.float64, device=dev, generator=gen)
        c /= float(ndet) ** 0.5
        sigma = torch.empty_like(c)

        def step():
            sharded.Hamiltonian(c, sigma)
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        step()
    barrier()
    launches_per_step = builder.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(ctx.local_rank) if clocks else None
    if sampler:
        sampler.__enter__()
    barrier()
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    barrier()
    if sampler:
        sampler.__exit__()
    ms_total = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    last = builder.last_build_ms()  # [H0, same-spin alpha, same-spin beta, opposite spin, total]
    phases = None
    if world > 1:  # every rank takes part: one extra build with a CUDA event at each phase boundary
        phases = [(nm, round(ms, 3)) for nm, ms in sharded.profile_phases(c, sigma)]
    ms_per_step = ms_total / steps
    lib_ms = lib_err = None
    if world > 1 and sharded.form == "rows" and sharded._lib is None:
        # the same build with the collectives issued by the C library (f2_dist_init / f2_dist_apply, csrc/dist_nccl.inc)
        try:
            from forte2_b200.dist import LibraryShardedSigma
            libsh = LibraryShardedSigma(builder, rank, world)
            sigma2 = torch.empty_like(sigma)
            for _ in range(max(warmup, 1)):
                libsh.Hamiltonian(c, sigma2)
            barrier()
            e0.record()
            for _ in range(steps):
                libsh.Hamiltonian(c, sigma2)
            e1.record()
            barrier()
            t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            lib_ms = float(t.item()) / steps
            sharded.Hamiltonian(c, sigma)
            lib_err = float((sigma2 - sigma).norm() / sigma.norm())
            libsh.close()
            del sigma2
        except Exception as e:  # noqa: BLE001
            lib_err = f"library NCCL path failed: {e}"[:200]
    res = {"name": name, "ndet": ndet, "nblocks": strings.nblocks, "setup_s": round(setup_s, 2),
           "strings_and_lists_s": round(t_strings, 2), "ms_per_step": ms_per_step, "value": 1e3 / ms_per_step,
           "steps": steps, "warmup": warmup, "set_memory_mb": memory_mb,
           "term_ms_last_step": {"H0": last[0], "same_spin_alpha": last[1], "same_spin_beta": last[2],
                                 "opposite_spin": last[3], "total": last[4]},
           "launches_per_step": int(launches_per_step),
           "peak_mem_gb": round(torch.cuda.max_memory_allocated(dev) / 1e9, 1)}
    if sampler:
        res["clocks"] = sampler.summary()
    if world > 1:
        res["sharded"] = {"form": sharded.form,
                          "collectives": "library (f2_dist_apply)" if sharded._lib is not None else "python (forte2_b200/dist.py)",
                          "local_compute_ms_rank0": last[4],
                          "exchange_ms_rank0": ms_per_step - last[4], "phases_ms_rank0_one_extra_build": phases,
                          "library_nccl_ms_per_step": lib_ms, "library_nccl_rel_diff": lib_err,
                          "note": "exchange = step time minus rank 0's kernel time: the NCCL collectives of "
                                  "forte2_b200/dist.py and their pack/unpack copies"}
    if parity:
        # the checker: sampled Slater rules on the host against the sigma just timed (oracle/ is test infrastructure,
        # used here as the checker only)
        from oracle import sampled
        tp = time.perf_counter()
        if world == 1:
            err, ns = sampled.sampled_rel_error(space, E, H, V, c, sigma, nsamp=nsamp)
        else:
            err, ns = sampled.sharded_rel_error(space, E, H, V, sharded.layouts, rank, c, sigma,
                                                nsamp_per_rank=max(64, nsamp // world))
        res["parity_sampled_rel"] = err
        res["parity_sampled_n"] = ns
        res["parity_sampled_s"] = round(time.perf_counter() - tp, 1)
    live = dict(strings=strings, builder=builder, c=c, sigma=sigma, sharded=sharded, ints=(E, H, V), step=step)
    return res, live


def kernel_rooflines(name, builder, world, last, peak_tflops, sharded_form=None):
    """the opposite-spin kernel against the FP64 tensor peak and the same-spin pair against HBM (and, where the
    closed form exists, against the tensor peak and the L2 reduction rate that actually bound it)"""
    n, na, nb = WORKLOADS[name]
    ndet = NDET.get(name, comb(n, na) * comb(n, nb))
    shape = builder.t3_shape()
    frac_work = 1.0 / world  # each rank contracts its own rows
    useful = 2.0 * shape["M"] * shape["k_cols"] * frac_work
    executed = 2.0 * shape["M_pad"] * shape["k_cols"] * frac_work
    survey = 2.0 * (n * (n + 1) / 2) ** 2 * ndet * frac_work
    t3_s = last[3] * 1e-3
    gather = shape["form"] == "gather" and (world == 1 or sharded_form == "rows")
    kern = ("t3g_kernel" if name.startswith("cas") else "t3gp_kernel") if gather else "t3_kernel"
    roof = {
        "bound": "tensor",
        "kernel": f"{kern} (opposite-spin, {'gather' if gather else 'scatter'} form, DMMA)",
        "achieved": useful / t3_s / 1e12, "peak": peak_tflops, "unit": "TFLOP/s",
        "frac": useful / t3_s / 1e12 / peak_tflops,
        "peak_source": "cuBLAS DGEMM 8192^3 measured in this run (MEASURED_PEAKS.json has no FP64 entry)",
        "flops_per_launch": useful, "executed_flops_per_launch": executed,
        "executed_tflops": executed / t3_s / 1e12,
        "survey_F_sigma3_flops": survey, "survey_F_sigma3_tflops_equiv": survey / t3_s / 1e12,
        "kernel_ms": last[3], "kernel_share_of_step": last[3] / last[4] if last[4] else None,
        "traffic": None,
    }
    if world == 1 and gather:
        tr = ncu_traffic(name, "t3g")
        if tr:
            roof["traffic"] = tr["bytes"]
            roof["traffic_source"] = tr["source"] + " (ncu --set full, one launch)"
            roof["algorithmic_hbm_bytes"] = 16.0 * ndet
    hbm, _ = hbm_peak_gbs()
    ss_ms = last[1] + last[2]
    ss_bytes = 2 * 24.0 * ndet * frac_work
    same = {"bound": "hbm", "kernels": "ssg_kernel (alpha) + ssg_kernel (beta), one-electron folded",
            "algorithmic_bytes": ss_bytes, "ms": ss_ms,
            "achieved": ss_bytes / (ss_ms * 1e-3) / 1e9 if ss_ms else None, "peak": hbm, "unit": "GB/s",
            "frac": ss_bytes / (ss_ms * 1e-3) / 1e9 / hbm if ss_ms else None,
            "note": "not HBM-bound at this size (DESIGN.md section 3): the binding resources are the FP64 tensor pipe "
                    "and the L2 atomic units"}
    if name.startswith("cas") and ss_ms:
        # flops of the same-spin contraction actually executed: per spin sum_K m_K^2 * columns * 2 with
        # m = C(n - N + 2, 2) parents per (N-2)-electron string, one-electron part folded in
        def ss_flops(ne, other):
            if ne < 2:
                return 2.0 * n * n * comb(n, ne) * other if ne == 1 else 0.0
            m = comb(n - ne + 2, 2)
            return 2.0 * comb(n, ne - 2) * m * m * other
        fl = (ss_flops(na, comb(n, nb)) + ss_flops(nb, comb(n, na))) * frac_work
        red = sum(comb(n, ne - 2) * comb(n - ne + 2, 2) * o for ne, o in ((na, comb(n, nb)), (nb, comb(n, na)))
                  if ne >= 2) * frac_work
        same.update({"tensor_tflops": fl / (ss_ms * 1e-3) / 1e12, "tensor_frac": fl / (ss_ms * 1e-3) / 1e12 / peak_tflops,
                     "reductions_per_s": red / (ss_ms * 1e-3), "reduction_rate_frac": red / (ss_ms * 1e-3) / 345e9,
                     "reduction_rate_peak": "345e9 FP64 red.global.add per s, scripts/micro/red_bench.cu "
                                            "(profiles/r01u_red_bench.log)"})
    return roof, same


def sharded_davidson(ctx, live, nroot=1, iters=3, form="auto"):
    from forte2_b200.ci_solver import ShardedCISolver
    from forte2_b200.davidson import DavidsonLiuParams
    E, H, V = live["ints"]
    solver = ShardedCISolver(live["strings"], E, H, V, nroot=nroot, davidson_liu_params=DavidsonLiuParams(maxiter=iters),
                             log_level=0, device=ctx.local_rank, builder=live["builder"], form=form)
    solver.sharded = live["sharded"]
    try:
        solver.run()
    except Exception as e:  # maxiter reached is not an error for a timing run
        if "converge" not in str(e).lower():
            raise
    ds = solver.eigensolver
    out = {"nroot": nroot, "iterations_run": len(ds.iter_times),
           "iter_ms_all": [round(t * 1e3, 2) for t in ds.iter_times],
           "sigma_builds_per_iter": [h[5] for h in ds.history],
           "iter_ms_steady": round(sorted(ds.iter_times[1:] or ds.iter_times)[0] * 1e3, 2)}
    del solver, ds
    ctx.torch.cuda.empty_cache()
    return out


def run_ours(args):
    import numpy as np
    import torch

    import forte2_b200 as f2

    ctx = Ctx()
    ctx.torch = torch
    ctx.world = world = int(os.environ.get("WORLD_SIZE", "1"))
    ctx.rank = rank = int(os.environ.get("RANK", "0"))
    ctx.local_rank = local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    ctx.dev = dev = torch.device("cuda", local_rank)
    ctx.dist = dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        ctx.dist = dist

    main_mb = 96 * 1024
    res, live = measure_space(ctx, args.workload, args.steps, args.warmup, main_mb, form=args.form,
                              parity=not args.no_parity, clocks=True)
    builder, strings, c, sigma = live["builder"], live["strings"], live["c"], live["sigma"]
    E, H, V = live["ints"]
    n, na, nb = WORKLOADS[args.workload]
    ndet = strings.ndet
    last = [res["term_ms_last_step"][k] for k in ("H0", "same_spin_alpha", "same_spin_beta", "opposite_spin", "total")]
    ms_per_step, value = res["ms_per_step"], res["value"]

    # ---- forte2's default scratch budget (N = 1): the beta same-spin term in several transposed row slabs
    default_memory = None
    if world == 1 and not args.no_default_memory:
        builder.set_memory(1024)
        reps = max(2, min(args.steps, 5))
        live["step"]()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            live["step"]()
        e1.record()
        torch.cuda.synchronize()
        lm = builder.last_build_ms()
        default_memory = {"set_memory_mb": 1024, "ms_per_step": e0.elapsed_time(e1) / reps, "steps": reps,
                          "same_spin_beta_ms": lm[2], "launches_per_step": builder.launch_count(),
                          "note": "forte2 passes ci_builder_memory = 1024 MB (forte2/base_classes/params.py:149-178); "
                                  f"the headline step uses set_memory({main_mb})"}
        builder.set_memory(main_mb)

    peak = None
    roofline = same_spin = None
    if rank == 0:
        peak = measure_fp64_peak(torch)
        roofline, same_spin = kernel_rooflines(args.workload, builder, world, last, peak,
                                               live["sharded"].form if live["sharded"] else None)

    # ---- end to end through the numpy-facing call (host buffers, H2D + D2H inside) ----
    e2e = None
    if world == 1 and args.no_e2e:
        e2e = {"value": None, "unit": "sigma builds/s", "h2d_bytes_per_step": ndet * 8, "d2h_bytes_per_step": ndet * 8,
               "note": "skipped (--no-e2e)"}
    elif world == 1:
        c_host = c.cpu().numpy()
        s_host = np.empty_like(c_host)
        for _ in range(3):  # warm-up: pinned staging buffers, streams, the copy pool's threads, first-touch of s_host
            builder.Hamiltonian(c_host, s_host)
        reps = max(1, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(reps):
            builder.Hamiltonian(c_host, s_host)
        dt = (time.perf_counter() - t0) / reps
        e2e = {"value": 1.0 / dt, "unit": "sigma builds/s", "h2d_bytes_per_step": ndet * 8,
               "d2h_bytes_per_step": ndet * 8, "ms_per_step": dt * 1e3, "steps": reps,
               "api": "CISigmaBuilder.Hamiltonian(numpy, numpy) -> f2_sigma_apply_host"}
        del c_host, s_host
    else:
        try:
            # N > 1: every rank's shard from and to pinned host memory through ShardedSigma.Hamiltonian_host (H2D of the
            # shard, sharded build, D2H of the result inside the timed call); CUDA events, max over ranks
            sharded = live["sharded"]
            c_pin = torch.empty(max(sharded.local_size, 1), dtype=torch.float64, pin_memory=True)[: sharded.local_size]
            s_pin = torch.empty(max(sharded.local_size, 1), dtype=torch.float64, pin_memory=True)[: sharded.local_size]
            c_pin.copy_(c)
            for _ in range(3):
                sharded.Hamiltonian_host(c_pin, s_pin)
            reps = max(1, min(args.steps, 5))
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                sharded.Hamiltonian_host(c_pin, s_pin)
            e1.record()
            dist.barrier()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1), float((s_pin.to(dev) - sigma).norm() ** 2), float(sigma.norm() ** 2)],
                             dtype=torch.float64, device=dev)
            tmax = t[:1].clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            dt = float(tmax.item()) / reps
            e2e = {"value": 1e3 / dt, "unit": "sigma builds/s", "h2d_bytes_per_step": ndet * 8, "d2h_bytes_per_step": ndet * 8,
                   "ms_per_step": dt, "steps": reps, "rel_diff_vs_device_path": float((t[1] / t[2]).sqrt().item()),
                   "api": "ShardedSigma.Hamiltonian_host(pinned shard, pinned shard): H2D + sharded build + D2H on every rank"}
            del c_pin, s_pin
        except Exception as e:  # noqa: BLE001  (the line must still be printed; every rank takes the same branch or fails alike)
            e2e = {"value": value, "unit": "sigma builds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                   "note": f"host path at N>1 failed ({e}); value repeats the device-resident figure"[:300]}

    # ---- one Davidson-Liu iteration of a 3-root solve (N = 1): the caller on either side of sigma ----
    davidson = None
    if world == 1 and not args.no_davidson:
        del c, sigma
        live["c"] = live["sigma"] = live["step"] = None
        torch.cuda.empty_cache()
        try:
            davidson = davidson_leg(torch, strings, builder, E, H, V, ms_per_step)
        except torch.cuda.OutOfMemoryError as e:  # the 3-root subspace needs 27 vectors
            davidson = {"error": f"out of memory: {e}"[:200]}

    # ---- 1- and 2-RDMs of one vector (N = 1): the other caller after the solve ----
    rdm = None
    if world == 1 and not args.no_davidson:
        try:
            rdm = rdm_leg(torch, strings, builder, peak)
        except Exception as e:  # noqa: BLE001 -- the leg is informative, the headline stands without it
            rdm = {"error": str(e)[:200]}

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
            v, info = cpu_baseline_sample(args.workload, check=gpu_parity)
        cpu = {"value": v, "unit": "sigma builds/s", **info}

    # ---- BASELINE.json configs[3] and [4], sharded (N = 8 by default) ----
    legs = None
    names = [x for x in (args.legs.split(",") if args.legs not in ("auto", "none") else
                         (["cas18", "gas20"] if (args.legs == "auto" and world == 8) else [])) if x]
    if names:
        legs = {}
        live.clear()
        del builder, strings
        torch.cuda.empty_cache()
        for name in names:
            torch.cuda.reset_peak_memory_stats(dev)
            # spaces with several determinant blocks: the first four builds of a builder take turns through the shapes of
            # t3gp_kernel (launch_t3gm) - they are warm-up, not measurement
            r, lv = measure_space(ctx, name, 3, 4 if name.startswith("gas") else 1, 64 * 1024, form=args.form, nsamp=1024,
                                  parity=not args.no_parity)
            lastn = [r["term_ms_last_step"][k] for k in ("H0", "same_spin_alpha", "same_spin_beta", "opposite_spin", "total")]
            if rank == 0:
                r["roofline"], r["roofline_same_spin"] = kernel_rooflines(
                    name, lv["builder"], world, lastn, peak, lv["sharded"].form if lv["sharded"] else None)
            r["config"] = workload_config(name, world)
            if world > 1 and not args.no_davidson:
                lv["c"] = lv["sigma"] = lv["step"] = None
                torch.cuda.empty_cache()
                r["davidson"] = sharded_davidson(ctx, lv, nroot=1, iters=3, form=args.form)
                r["davidson"]["note"] = ("one-root Davidson-Liu on the sharded vectors (ShardedCISolver); iter_ms_steady = "
                                         "fastest iteration after the first; one sigma build per iteration")
            r["peak_mem_gb"] = round(torch.cuda.max_memory_allocated(dev) / 1e9, 1)
            legs[name] = r
            lv.clear()
            torch.cuda.empty_cache()

    if rank == 0:
        line = {
            "metric": "sigma_builds_per_s", "value": value, "unit": "sigma builds/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.workload, world),
            "parity_sampled_rel": res.get("parity_sampled_rel"), "parity_sampled_n": res.get("parity_sampled_n"),
            "parity_note": "relative L2 error of the timed sigma at seeded sample determinants against Slater rules on "
                           "the host (oracle/sampled.py); the bar is 1e-12",
            "term_ms_last_step": res["term_ms_last_step"],
            "roofline": roofline, "roofline_same_spin": same_spin, "default_memory": default_memory,
            "cpu_baseline": cpu, "e2e": e2e, "davidson": davidson, "rdm": rdm,
            "sharded": res.get("sharded"), "legs": legs,
            "setup_s": res["setup_s"], "peak_mem_gb": res["peak_mem_gb"],
            "gpu_launches": int(res["launches_per_step"] * args.steps), "clocks": res.get("clocks"),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="cas16", choices=sorted(SPACES))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-davidson", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-default-memory", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (needs 2 x 8 ndet bytes of pinned memory)")
    ap.add_argument("--legs", default="auto",
                    help="extra sharded legs after the headline workload: auto (cas18,gas20 at 8 GPUs, none otherwise), "
                         "none, or a comma list of workloads")
    ap.add_argument("--form", default="auto", choices=["auto", "full", "rows"],
                    help="N > 1: communication form of forte2_b200/dist.py")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
