"""Run under torch.distributed.run (one rank per GPU): parity of the row-partitioned multi-GPU path against
the single-GPU path and the compiled oracle, then a timing at a larger scale.
usage: dist_check.py [parity_scale] [timing_scale] [iters]"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quadraticprogramsolver_b200 import solver as S  # noqa: E402
from workloads.problems import config_cfg5  # noqa: E402
from oracle import c_oracle  # noqa: E402  (the checker: this script is test infrastructure, tests/test_gpu_dist.py runs it)

pscale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.02
tscale = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 100
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
rank, world = dist.get_rank(), dist.get_world_size()

out = {}
P, q, A, l, u = config_cfg5(seed=1234, scale=pscale)
n, m = P.shape[0], A.shape[0]
modes = os.environ.get("QPB_DIST_MODES", "peer,nccl").split(",")
for mode, precond, adapt in [(mo, pc, ad) for mo in modes for pc, ad in (("jacobi", False), ("none", False), ("jacobi", True))]:
    kw = dict(numIterations=400, epsPcg=1e-10, precond=precond, adptRho=adapt, rho=0.1 if adapt else 1.0)
    with S.QPB200DistSolver(P, q, A, l, u, distMode=mode, **kw) as ds:
        x = np.zeros(n)
        flag = ds.solve(x, want_zy=True)
        info = dict(ds.info)
        rows = ds.rows
    with S.QPB200Solver(P, q, A, l, u, device=local_rank, **kw) as s1:
        x1 = np.zeros(n)
        flag1 = s1.solve(x1, want_zy=True)
        info1 = dict(s1.info)
    # the oracle (compiled restatement of the reference), same linear-solver mode and settings
    okw = {k: v for k, v in kw.items() if k != "precond"}
    xo, fo, io = c_oracle.solve_sparse(P, q, A, l, u, precond=1 if precond == "jacobi" else 0, **okw)
    err_o = float(np.max(np.abs(x - xo)) / (1 + np.max(np.abs(xo))))
    ok_o = int(flag) == int(fo) and abs(info["iterations"] - io["iterations"]) <= 2 and err_o <= 1e-6
    err = float(np.max(np.abs(x - x1)) / (1 + np.max(np.abs(x1))))
    zerr = float(np.max(np.abs(info["z_local"] - info1["z"][rows[0]:rows[1]]), initial=0.0) / (1 + np.max(np.abs(info1["z"]))))
    ok = (int(flag) == int(flag1) and abs(info["iterations"] - info1["iterations"]) <= 2 and err <= 1e-6 and zerr <= 1e-6
          and ok_o)
    res = dict(ok=bool(ok), err_vs_oracle=err_o, flag_oracle=int(fo), it_oracle=int(io["iterations"]), flag=int(flag), flag1=int(flag1), it=info["iterations"], it1=info1["iterations"], err=err, zerr=zerr,
               pcg=info["pcg_iters_total"], pcg1=info1["pcg_iters_total"], rho_updates=info["rho_updates"],
               ms=info["solve_ms"], ms1=info1["solve_ms"], launches=info["kernel_launches"])
    out[f"parity_{mode}_{precond}_{'adapt' if adapt else 'fixed'}"] = res
    t = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({f"parity_{mode}_{precond}_{adapt}": res, "all_ranks_ok": bool(t.item() == 1.0)}), flush=True)
    assert t.item() == 1.0, res

if tscale > 0:
    P, q, A, l, u = config_cfg5(seed=1234, scale=tscale)
    n = P.shape[0]
    for mode in modes:
        kw = dict(numIterations=iters)
        t0 = time.time()
        with S.QPB200DistSolver(P, q, A, l, u, distMode=mode, **kw) as ds:
            create_s = time.time() - t0
            for rep in range(3):
                x = np.zeros(n)
                dist.barrier(); torch.cuda.synchronize()
                t0 = time.time()
                ds.solve(x)
                torch.cuda.synchronize()
                wall = time.time() - t0
            info = dict(ds.info)
        t = torch.tensor([info["solve_ms"]], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            res = dict(mode=mode, world=world, n=n, create_s=create_s, wall_s=wall, solve_ms_max=float(t.item()),
                       iters=info["iterations"], pcg=info["pcg_iters_total"],
                       it_per_s=info["iterations"] / (float(t.item()) * 1e-3), launches=info["kernel_launches"])
            print(json.dumps({"timing_" + mode: res}), flush=True)
            out["timing_" + mode] = res
if rank == 0:
    json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", f"dist_check_{world}.json"), "w"), indent=1)
dist.destroy_process_group()
