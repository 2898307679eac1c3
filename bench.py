#!/usr/bin/env python
"""bench.py -- ADMM iterations/s of the B200 path on BASELINE.json's headline configuration.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload cfg5|cfg2|cfg3]

Workload at N = 1 (default): configs[4], the 1M-variable sparse QP the metric's target is quoted on
(n = 1e6, m = 2e6, d = 5e-6, same recipe as GenerateRandomQP(randomQp) with the bounds centred on A x* so
that the QP is feasible -- see quadraticprogramsolver_b200/problems.py::config_sparse).  A *step* is one
pass of the hot path over that QP: ADMM iterations 1..ITERS of SolveQuadraticProgram! from x = 0
(ITERS = 100, four convergence checks), with the reference's default settings and Jacobi-PCG.

  value : ADMM iterations/s, problem resident in HBM, device time (CUDA events inside qpb200_solve)
  e2e   : same metric through the public call SolveQuadraticProgram(P, q, A, l, u) with HOST buffers:
          qpb200_create (conversion + upload) + qpb200_solve (+ x, z, y download) + destroy, wall clock
  roofline : the persistent ADMM kernel (the one launch of a step): algorithmic bytes of the matrix and
          vector passes it executed / its CUDA-event duration, against the measured HBM
          peak of MEASURED_PEAKS.json; plus the stand-alone SpMV numbers
  cpu_baseline : oracle/qp_oracle.c (a compiled restatement -- Julia is unavailable) on all host
          cores, same QP, same settings, a time-bounded sample of the same iterations

`--impl reference` times that CPU restatement alone (kind "port").
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

ITERS = 100                     # ADMM iterations per step
CPU_SAMPLE_SECONDS = 20.0       # bounded CPU sample


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_fp64_peak():
    """FP64 FMA peak for the dense-batch roofline: measured live by scripts/microbench/fp64_peak_bench (built by
    scripts/microbench/Makefile, travels to the GPU box) when present, else the committed measurement, else nominal."""
    exe = os.path.join(ROOT, "scripts", "microbench", "fp64_peak_bench")
    try:
        out = subprocess.run([exe], capture_output=True, text=True, timeout=60, check=True).stdout
        rec = json.loads(out.strip().splitlines()[-1])
        return float(rec["dfma_tflops"]), f"measured in this run (scripts/microbench/fp64_peak_bench: DFMA {rec['dfma_tflops']} TF/s, DMMA {rec['dmma_tflops']} TF/s)"
    except Exception:
        pass
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "r2_fp64_peak.json")))
        return float(rec["dfma_tflops"]), "profiles/r2_fp64_peak.json (scripts/microbench/fp64_peak_bench on a B200 of this pool)"
    except Exception:
        return 40.0, "nominal B200 FP64 (no measurement available)"


def load_traffic(workload, world):
    """dram__bytes_read.sum + dram__bytes_write.sum of the timed kernel (one launch = one step), from the committed ncu
    capture of this very launch (profiles/r2_traffic.json, written by scripts/ncu_traffic.py); None when no capture of
    the workload exists -- never a typed-in constant."""
    path = os.path.join(ROOT, "profiles", "r2_traffic.json")
    try:
        rec = json.load(open(path))[f"{workload}_n{world}"]
        return float(rec["dram_bytes_per_launch"]), f"profiles/r2_traffic.json ({rec.get('git', '?')}, {rec.get('what', '')})"
    except Exception:
        return None, None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.thread, self.gpu_index = [], None, None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu_index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max((float(r[3]) for r in self.rows if len(r) > 3 and r[3].replace('.', '', 1).isdigit()),
                                   default=None),
                "samples": len(sm), "reasons": sorted(reasons)}


def make_workload(name, scale):
    from workloads import problems
    t0 = time.time()
    if name == "cfg5":
        P, q, A, l, u = problems.config_cfg5(seed=1234, scale=scale)
        desc = f"cfg5 sparse QP n={P.shape[0]} m={A.shape[0]} nnz(P)={P.nnz} nnz(A)={A.nnz} (randomQp recipe, d=5/n, feasible bounds)"
    elif name == "cfg4":
        P, q, A, l, u = problems.config_cfg4(seed=1234, scale=scale)
        desc = (f"cfg4 constrained least squares (README form) as QP: n={P.shape[0]} m={A.shape[0]} nnz(P=A'A)={P.nnz} "
                f"nnz([B;D])={A.nnz}; A_ls {4 * P.shape[0]}x{P.shape[0]}, B {P.shape[0] // 2} rows (<= c), D {P.shape[0] // 10} rows (= e), 5 nnz/row")
    elif name == "banded":
        P, q, A, l, u = problems.config_banded()
        desc = (f"banded QP n={P.shape[0]} m={A.shape[0]} nnz(P)={P.nnz} nnz(A)={A.nnz}: cfg5's sizes and non-zeros per row, every "
                "row's columns within +-256 of the diagonal (gathers with locality; separates the SpMV engine from cfg5's random columns)")
    elif name == "cfg2":
        P, q, A, l, u = problems.config_cfg2(seed=1234)
        desc = f"cfg2 sparse QP n={P.shape[0]} m={A.shape[0]} nnz(P)={P.nnz} nnz(A)={A.nnz} (d=1e-3, feasible bounds)"
    else:
        raise SystemExit(f"unknown workload {name}")
    return (P, q, A, l, u), desc, time.time() - t0


def solver_kwargs():
    # reference defaults (SolveQuadraticProgram.jl:15-17, LinearSystemSolvers.jl:125) + iteration cap
    return dict(numIterations=ITERS, epsAbs=1e-6, epsRel=1e-6, rho=1.0, sigma=1e-6, alpha=1.6, adptRho=False,
                numItrConv=25, epsPcg=1e-6, numItrPcg=1000)


def bench_config(args, desc):
    """The `config` object: identical in both arms (b200 / reference) for the same command line."""
    world = max(1, args.gpus)
    return {"workload": desc, "iters_per_step": ITERS,
            "settings": "reference defaults (rho=1, sigma=1e-6, alpha=1.6, eps 1e-6, check every 25), Jacobi-PCG abstol 1e-6",
            "parallelism": "1 GPU" if world == 1 else f"one QP row-partitioned over {world} GPUs (rows of A / columns of P per rank)",
            "l2": "matrix streams exceed the 126 MB L2 at cfg5 size; stand-alone SpMV timings flush L2 between launches"}


def cpu_sample(prob, precond, seconds, **over):
    """Time-bounded run of the compiled oracle on the same QP, on ALL host cores whatever OMP_NUM_THREADS says
    (torch.distributed.run exports OMP_NUM_THREADS=1)."""
    from oracle import c_oracle
    c_oracle.use_all_cores()
    P, q, A, l, u = prob
    kw = solver_kwargs()
    kw.update(over)
    x, flag, info = c_oracle.solve_sparse(P, q, A, l, u, precond=precond, time_limit_s=seconds, **kw)
    its, sec = info["iterations"], info["solve_seconds"]
    return {"value": its / sec, "unit": "iter/s", "cores": c_oracle.num_threads(), "kind": "port",
            "sample": f"ADMM iterations 1..{its} of the same QP ({sec:.1f} s wall, oracle/qp_oracle.c, "
                      f"{'Jacobi-PCG' if precond else 'un-preconditioned CG (reference)'}, OpenMP)",
            "cg_iters_per_s": info["cg_iters_total"] / sec, "cg_iters_per_admm_iter": info["cg_iters_total"] / max(1, its),
            "iterations": its, "host_cpus": os.cpu_count(), "x": x, "flag": int(flag), "info": info}


def _public(cb):
    return {k: v for k, v in cb.items() if k not in ("x", "flag", "info")}


def quiet_gate(dist, key, open_it):
    """N > 1: the ranks that have nothing to do while rank 0 runs host-only work SLEEP on the process group's
    rendezvous store (a blocking socket read) instead of spinning in a NCCL barrier -- a spinning rank (host thread
    polling the stream, NCCL proxy) takes cores away from the OpenMP team of the CPU port: measured on the 2-GPU
    box, the 25-iteration parity leg took 77-86 s next to a rank waiting in NCCL against 15 s alone."""
    if dist is None:
        return
    try:
        from datetime import timedelta

        from torch.distributed.distributed_c10d import _get_default_store
        store = _get_default_store()
        if open_it:
            store.set(key, "1")
        else:
            store.wait([key], timedelta(seconds=3600))
    except Exception as e:                      # no store: fall through, the collective that follows still synchronises
        print(f"quiet_gate({key}): {e}", file=sys.stderr)


def parity_check(prob, solve_gpu, iters=25, eps_pcg=1e-10, after_cpu=None):
    """north_star's criterion on the BENCHMARKED problem: the CPU port and the GPU run the same `iters` ADMM
    iterations (tight inner solve so that the trajectory is well defined) and x, flag, iteration count are compared."""
    over = dict(numIterations=iters, epsPcg=eps_pcg)
    t0 = time.time()
    ref = cpu_sample(prob, 1, 0.0, **over)
    cpu_s = time.time() - t0
    if after_cpu is not None:
        after_cpu()
    x_gpu, flag_gpu, info_gpu = solve_gpu(over)
    xr = ref["x"]
    err = float(np.max(np.abs(x_gpu - xr)) / (1.0 + np.max(np.abs(xr))))
    return {"admm_iterations": iters, "eps_pcg": eps_pcg, "flag_gpu": int(flag_gpu), "flag_cpu_port": ref["flag"],
            "iterations_gpu": int(info_gpu["iterations"]), "iterations_cpu_port": int(ref["iterations"]),
            "cg_iters_gpu": int(info_gpu["pcg_iters_total"]), "cg_iters_cpu_port": int(ref["info"]["cg_iters_total"]),
            "x_rel_err_inf": err, "tolerance": 1e-6, "pass": bool(err <= 1e-6 and int(flag_gpu) == ref["flag"]
                                                                  and abs(int(info_gpu["iterations"]) - int(ref["iterations"])) <= 2),
            "res_prim_gpu": float(info_gpu["res_prim"]), "res_prim_cpu_port": float(ref["info"]["res_prim"]),
            "res_dual_gpu": float(info_gpu["res_dual"]), "res_dual_cpu_port": float(ref["info"]["res_dual"]),
            "cpu_seconds": round(cpu_s, 1), "criterion": "|x - x_ref|inf <= 1e-6 (1 + |x_ref|inf), same flag, iterations within 2"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    prob, desc, gen_s = make_workload(args.workload, args.scale)
    # warm-up steps only need to touch the matrices and spin the thread pool up: 1.5 s each; the timed steps share
    # what is left of ~170 s.  A sample stops at ADMM iteration ITERS at the latest (the b200 arm's range 1..ITERS).
    warm_s = 1.5
    per_step = max(2.0, min(CPU_SAMPLE_SECONDS, (170.0 - warm_s * args.warmup) / max(1, args.steps)))
    for _ in range(args.warmup):
        cpu_sample(prob, 1, warm_s)
    vals, samples = [], []
    t0 = time.time()
    for _ in range(args.steps):
        s = cpu_sample(prob, 1, per_step)
        vals.append(s["value"]); samples.append(s)
    wall = time.time() - t0
    its = sum(s["iterations"] for s in samples)
    value = its / sum(s["iterations"] / s["value"] for s in samples)
    cb = _public(samples[-1]); cb["value"] = value
    cb["cg_iters_per_admm_iter"] = sum(s["info"]["cg_iters_total"] for s in samples) / max(1, its)
    cb["sample"] += f"; {args.steps} such samples, {warm_s} s warm-up samples"
    line = {"impl": "reference", "metric": "admm_iters_per_s", "value": value, "unit": "iter/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(1, args.steps),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": bench_config(args, desc),
            "cpu_baseline": cb,
            "e2e": {"value": value, "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def run_b200(args):
    from quadraticprogramsolver_b200 import solver as S

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch multi-GPU runs with torch.distributed.run (one rank per GPU)")
    import torch
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist = dist_mod

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    prob, desc, gen_s = make_workload(args.workload, args.scale)
    P, q, A, l, u = prob
    n, m = P.shape[0], A.shape[0]
    kw = solver_kwargs()
    kw["device"] = local_rank
    peak, peak_src = load_peaks()

    # ---- device-resident arm ------------------------------------------------------------------
    # N = 1: the single-GPU persistent kernel.  N > 1: the SAME QP row-partitioned over the N ranks
    # (rows of A / columns of P per rank, n-vector partial sums reduced in-kernel over NVLink) -> strong scaling.
    if world == 1:
        s = S.QPB200Solver(P, q, A, l, u, **kw)
    else:
        s = S.QPB200DistSolver(P, q, A, l, u, **kw)          # qpb200_dist_create_full: the library partitions the QP
    x = np.zeros(n)
    for _ in range(args.warmup):
        x[:] = 0.0
        s.solve(x)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    dev_ms, iters, pcg, launches, bytes_total = 0.0, 0, 0, 0, 0
    for _ in range(args.steps):
        x[:] = 0.0
        s.solve(x)
        dev_ms += s.info["solve_ms"]; iters += s.info["iterations"]; pcg += s.info["pcg_iters_total"]
        launches += s.info["kernel_launches"]
        bytes_total += s.apply_bytes(100)
    barrier()
    wall_resident = time.perf_counter() - t0
    clocks = sampler.stop()
    flag = int(s.info["conv_flag"])
    if dist is not None:
        t = torch.tensor([dev_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms = float(t.item())
        t = torch.tensor([float(bytes_total)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        bytes_total = float(t.item())
    value = iters / (dev_ms * 1e-3)               # one QP: iterations of the whole job per second
    achieved = bytes_total / 1e9 / (dev_ms * 1e-3) / world   # per-GPU algorithmic GB/s

    # stand-alone SpMV (the operator the north_star quotes): L2 flushed between launches
    spmv = {}
    if rank == 0 and world == 1:
        for which, name in ((1, "A"), (4, "H=[P A']")):
            ms = s.time_apply(which, reps=20, flush_l2=True)
            gb = s.apply_bytes(which) / 1e9
            spmv[name] = {"ms": ms, "GBs": gb / (ms * 1e-3), "frac": gb / (ms * 1e-3) / peak}
    # ---- parity on the benchmarked problem (rank 0 runs the CPU port; every rank runs the GPU solve)
    parity = None
    if not args.no_parity:
        def solve_gpu(over):
            kk = dict(kw); kk.update(over)
            if world == 1:
                s.update_settings(**kk)
                xx = np.zeros(n)
                fl = s.solve(xx)
                return xx, fl, s.info
            with S.QPB200DistSolver(P, q, A, l, u, **kk) as ds:
                xx = np.zeros(n)
                fl = ds.solve(xx)
                return xx, fl, ds.info
        if rank == 0:
            parity = parity_check(prob, solve_gpu, after_cpu=lambda: quiet_gate(dist, "qpb200_parity_cpu_done", True))
        else:
            quiet_gate(dist, "qpb200_parity_cpu_done", False)      # sleep while rank 0 runs the CPU port
            solve_gpu(dict(numIterations=25, epsPcg=1e-10))
        barrier()
    s.close()

    # ---- end-to-end arm: host buffers -> create -> solve -> results on host, every step --------
    Pp, Pi, Pv = S._csc_arrays(P)
    Ap, Ai, Av = S._csc_arrays(A)
    if world == 1:
        m_loc = m
        h2d = Pp.nbytes + Pi.nbytes + Pv.nbytes + Ap.nbytes + Ai.nbytes + Av.nbytes + 8 * (2 * n + 2 * m_loc)
    else:
        # what this rank uploads: its slice (rows of A, columns of P) as the library cuts it; counted from the partition
        from quadraticprogramsolver_b200 import partition
        rb, cb = partition.plan(P, A, world)
        i0, i1, j0, j1 = int(rb[rank]), int(rb[rank + 1]), int(cb[rank]), int(cb[rank + 1])
        m_loc = i1 - i0
        nnz_p = int(Pp[j1] - Pp[j0])
        nnz_a = int(np.count_nonzero((Ai >= i0) & (Ai < i1)))
        h2d = 16 * (nnz_p + nnz_a) + 8 * 2 * (n + 1) + 8 * (2 * n + 2 * m_loc)
    d2h = 8 * (n + 2 * m_loc)

    # The host buffers are what the reference caller owns: SparseMatrixCSC arrays (Int64 indices), q, l, u.
    # (scipy stores int32 indices; the one-off widening to Julia's Int64 layout is not part of the path.)
    Parr, Aarr = (Pp, Pi, Pv), (Ap, Ai, Av)
    q64, l64, u64 = np.ascontiguousarray(q), np.ascontiguousarray(l), np.ascontiguousarray(u)

    def e2e_once():
        if world == 1:
            xx = np.zeros(n)
            return S.solve_csc_arrays(n, m, Parr, q64, Aarr, l64, u64, xx, want_zy=True, **kw)[1]
        # the caller holds the whole P, A on the host (SparseMatrixCSC arrays): cutting this rank's slice out of them is
        # part of every call (qpb200_dist_create_full)
        with S.QPB200DistSolver(None, q64, None, l64, u64, arrays=(Parr, Aarr), **kw) as ds:
            xx = np.zeros(n)
            ds.solve(xx, want_zy=True)
            return ds.info

    e2e_steps = max(1, min(args.steps, 3))
    e2e_once()                                            # warm-up
    barrier()
    t0 = time.perf_counter()
    e2e_iters = 0
    e2e_create_ms, e2e_solve_ms = 0.0, 0.0
    for _ in range(e2e_steps):
        info = e2e_once()
        e2e_iters += info["iterations"]
        e2e_create_ms += info["setup_ms"]; e2e_solve_ms += info["solve_ms"]
    barrier()
    e2e_wall = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_wall, float(h2d), float(d2h)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_wall = float(t[0].item())
    e2e_value = e2e_iters / e2e_wall

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    traffic, traffic_src = load_traffic(args.workload, world)
    line = {
        "metric": "admm_iters_per_s", "value": value, "unit": "iter/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms / max(1, args.steps), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(args, desc),
        "run": {"conv_flag": flag, "pcg_iters_per_step": pcg / max(1, args.steps), "cg_iters_per_admm_iter": pcg / max(1, iters),
                "gen_s": round(gen_s, 1), "collectives": None if world == 1 else
                "one persistent kernel per GPU; reduce-scatter / all-gather of the n-vectors in-kernel over NVLink peer memory"},
        "parity": parity,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "iter/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "steps": e2e_steps, "ms_per_step": 1e3 * e2e_wall / e2e_steps,
                "breakdown_ms_per_step": {"qpb200_create": e2e_create_ms / e2e_steps, "solve_device": e2e_solve_ms / e2e_steps,
                                          "copies_destroy_host": 1e3 * e2e_wall / e2e_steps - (e2e_create_ms + e2e_solve_ms) / e2e_steps}},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src,
                     "kernel": "admm_kernel (persistent; 1 launch per step)" if world == 1 else
                     "admm_peer_sliced_kernel (persistent; 1 launch per GPU per step; per-GPU GB/s)", "peak_source": peak_src,
                     "spmv": spmv, "note": "cfg5's uniformly random columns make every 8-byte gather of x move a 32-byte L2 "
                     "sector: 44 B per non-zero cross the SM's L2->L1 port (ncu), which caps the H pass at 0.46 of the HBM peak; "
                     "the kernel runs at about 80 % of the best rate measured through that port, as does a structurally different "
                     "CSR-vector kernel (DESIGN.md 4.2, profiles/r1d_csrvec_spmv_bench.jsonl)"},
        "cg_iters_per_s": pcg / (dev_ms * 1e-3),
    }
    if world == 1 and not args.no_cpu:
        line["cpu_baseline"] = _public(cpu_sample(prob, 1, CPU_SAMPLE_SECONDS))
        line["cpu_baseline_reference_cg"] = _public(cpu_sample(prob, 0, CPU_SAMPLE_SECONDS / 2))
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def run_cfg3(args):
    """configs[2]: 65 536 small dense QPs (n = 64, m = 96), contiguous batch slices per GPU, no collective.
    metric: QP solves/s (whole batch / slowest rank's device time)."""
    from quadraticprogramsolver_b200 import solver as S
    from workloads.problems import config_cfg3_batch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    batch = int(args.batch)
    if args.impl == "reference":
        if rank != 0:
            return
        from oracle import c_oracle
        ns = min(batch, 4096)
        P, q, A, l, u = config_cfg3_batch(ns, 64, 96, seed=1234)
        X, fl, it, sec, rc = c_oracle.solve_dense_batch(P, q, A, l, u)
        v = ns / sec
        print(json.dumps({"impl": "reference", "metric": "qp_solves_per_s", "value": v, "unit": "solves/s", "n_gpus": args.gpus,
                          "steps": 1, "warmup": 0, "ms_per_step": 1e3 * sec, "higher_is_better": True, "scaling": "strong",
                          "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                          "config": {"workload": f"cfg3 dense QPs n=64 m=96, first {ns} problems of the batch"},
                          "cpu_baseline": {"value": v, "unit": "solves/s", "cores": c_oracle.num_threads(), "kind": "port",
                                           "sample": f"{ns} problems, oracle/qp_oracle.c dense Cholesky mode, OpenMP"},
                          "e2e": {"value": v, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))
        return
    import torch
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist = dist_mod

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    lo, hi = batch * rank // world, batch * (rank + 1) // world
    shared = args.workload == "cfg3shared"
    if shared:          # MPC-style: ONE (P, A) for the whole batch, 65 536 different (q, l, u) -- SURVEY.md 8(f) row 3
        from workloads.problems import config_cfg3_shared
        P, q, A, l, u = config_cfg3_shared(batch, 64, 96, seed=1234)
        q, l, u = q[lo:hi].copy(), l[lo:hi].copy(), u[lo:hi].copy()
    else:
        P, q, A, l, u = config_cfg3_batch(batch, 64, 96, seed=1234)
        P, q, A, l, u = P[lo:hi].copy(), q[lo:hi].copy(), A[lo:hi].copy(), l[lo:hi].copy(), u[lo:hi].copy()
    kw = dict(device=local_rank)
    b = S.QPB200Batch(P, q, A, l, u, **kw)
    for _ in range(args.warmup):
        b.solve()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    dev_ms, iters = 0.0, 0
    for _ in range(args.steps):
        X, flags, its = b.solve()
        dev_ms += b.info["solve_ms"]; iters += int(its.sum())
    barrier()
    clocks = sampler.stop()
    b.close()
    e2e_steps = max(1, min(args.steps, 2))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        S.SolveQuadraticProgramBatch(P, q, A, l, u, **kw)
    barrier()
    e2e_wall = time.perf_counter() - t0
    tot = [dev_ms, e2e_wall, float(iters)]
    if dist is not None:
        t = torch.tensor(tot[:2], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t2 = torch.tensor([tot[2]], device="cuda", dtype=torch.float64)
        dist.all_reduce(t2, op=dist.ReduceOp.SUM)
        tot = [float(t[0]), float(t[1]), float(t2[0])]
    if rank == 0:
        fp64_peak, fp64_src = load_fp64_peak()
        value = batch * args.steps / (tot[0] * 1e-3)
        flops = 2.0 * 16448.0 * tot[2]     # FMAs per ADMM iteration: 2 * 96 * 64 (A, A') + 2 * 2080 (L^-1, L^-T)
        line = {"metric": "qp_solves_per_s", "value": value, "unit": "solves/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": tot[0] / args.steps, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": (f"cfg3 (shared matrices, MPC-style): {batch} dense QPs n=64 m=96 with ONE (P, A) and "
                                        "different q, l, u; reference defaults; one Cholesky factor for the batch" if shared else
                                        f"cfg3: {batch} dense QPs n=64 m=96 (randomQp recipe d=1), reference defaults, Cholesky of "
                                        "P+sigma I+rho A'A per QP"), "parallelism": f"batch slices over {world} GPU(s), no collective",
                           "admm_iters_per_solve": tot[2] / (batch * args.steps)},
                "clocks": clocks,
                "e2e": {"value": batch * e2e_steps / tot[1], "unit": "solves/s", "h2d_bytes_per_step": int(8 * (P.size + A.size + q.size + l.size + u.size + q.size)) * world,
                        "d2h_bytes_per_step": int(8 * q.size + 12 * len(q)) * world},
                "gpu_launches": args.steps,
                "roofline": {"bound": "fp64_fma", "achieved": flops / (tot[0] * 1e-3) / 1e12 / world, "peak": fp64_peak, "unit": "TFLOP/s",
                             "frac": flops / (tot[0] * 1e-3) / 1e12 / world / fp64_peak, "traffic": None,
                             "peak_source": fp64_src,
                             "kernel": ("dense_shared_kernel (16 problems per CTA as the columns of three DMMA GEMMs per iteration, "
                                        "iterates in accumulator-fragment registers)" if shared else
                                        "dense_batch_kernel<96,2> (A and K^-1 register-resident during the iterations, shuffle "
                                        "reduce-scatters; DMMA only in the factorisation)")},
                "admm_iters_per_s": tot[2] / (tot[0] * 1e-3)}
        if world == 1 and not args.no_cpu:
            from oracle import c_oracle
            ns = 2048
            Pn = np.ascontiguousarray(np.broadcast_to(P, (ns, 64, 64))) if shared else P[:ns]
            An = np.ascontiguousarray(np.broadcast_to(A, (ns, 64, 96))) if shared else A[:ns]
            Xr, fr, ir, sec, rc = c_oracle.solve_dense_batch(Pn, q[:ns], An, l[:ns], u[:ns])
            line["cpu_baseline"] = {"value": ns / sec, "unit": "solves/s", "cores": c_oracle.num_threads(), "kind": "port",
                                    "sample": f"first {ns} problems of the batch, oracle/qp_oracle.c, OpenMP",
                                    "parity": {"flags_equal": bool(np.array_equal(fr, flags[:ns])),
                                               "iters_max_diff": int(np.max(np.abs(ir - its[:ns]))),
                                               "x_max_rel_err": float(np.max(np.abs(Xr - X[:ns]) / (1 + np.max(np.abs(Xr), axis=1, keepdims=True))))}}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg5", choices=["cfg5", "cfg2", "cfg3", "cfg3shared", "cfg4", "banded"])
    ap.add_argument("--batch", type=int, default=65536, help="cfg3 batch size")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink cfg5 (tests only)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--no-parity", action="store_true", help="skip the in-bench parity check against the CPU port")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    # The contract is ONE JSON line on stdout.  Native libraries write there too (NCCL prints "NCCL version ..." from
    # C when NCCL_DEBUG is set on the box), so file descriptor 1 points at stderr while the benchmark runs and the
    # JSON line goes to the real stdout at the end.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    buf = io.StringIO()
    try:
        with contextlib.redirect_stdout(buf):
            if args.workload in ("cfg3", "cfg3shared"):
                run_cfg3(args)
            elif args.impl == "reference":
                run_reference(args)
            else:
                run_b200(args)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
        out = buf.getvalue()
        if out:
            sys.stdout.write(out)
            sys.stdout.flush()


if __name__ == "__main__":
    main()
