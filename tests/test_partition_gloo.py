"""CPU tests of the multi-GPU host logic: the row/column partition of SURVEY.md 8(e), checked (a) in one
process and (b) across a real 2-rank torch.distributed group on the gloo backend, where every rank applies
its H_r = [P[:, J_r]  A_r'] slice and the n-vectors are combined by all_reduce -- the exact data flow of
dist_kernels.cuh with numpy standing in for the kernels."""
import os
import subprocess
import sys
import textwrap

import numpy as np

from quadraticprogramsolver_b200 import partition
from workloads.problems import config_sparse

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_balanced_blocks_cover_and_balance():
    w = np.random.default_rng(0).integers(0, 50, size=10000)
    for parts in (1, 2, 3, 8):
        b = partition.balanced_blocks(w, parts)
        assert b[0] == 0 and b[-1] == len(w) and np.all(np.diff(b) >= 0) and len(b) == parts + 1
        tot = np.array([w[b[i]:b[i + 1]].sum() for i in range(parts)])
        assert tot.max() <= tot.mean() * 1.1 + 100
    assert list(partition.balanced_blocks(np.zeros(5), 8)) == sorted(partition.balanced_blocks(np.zeros(5), 8))


def test_slices_reassemble_operator():
    P, q, A, l, u = config_sparse(400, 900, 0.02, seed=3)
    rng = np.random.default_rng(1)
    uvec = rng.standard_normal(400)
    yvec = rng.standard_normal(900)
    rho, sigma = 0.7, 1e-3
    for R in (1, 2, 4, 7):
        Ku = np.zeros(400); Px = np.zeros(400); Aty = np.zeros(400); dP = np.zeros(400); dAA = np.zeros(400)
        rows_seen = 0
        for r in range(R):
            P_r, A_r, l_r, u_r, (i0, i1), (j0, j1) = partition.slice_problem(P, A, l, u, r, R)
            assert P_r.shape == (400, 400) and A_r.shape == (i1 - i0, 400)
            assert np.array_equal(l_r, l[i0:i1]) and np.array_equal(u_r, u[i0:i1])
            rows_seen += i1 - i0
            Ku += P_r @ uvec + rho * (A_r.T @ (A_r @ uvec))
            Px += P_r @ uvec
            Aty += A_r.T @ yvec[i0:i1]
            dP += P_r.diagonal()
            dAA += np.asarray(A_r.multiply(A_r).sum(axis=0)).ravel()
        assert rows_seen == 900
        assert np.allclose(Ku + sigma * uvec, P @ uvec + rho * (A.T @ (A @ uvec)) + sigma * uvec, rtol=1e-12, atol=1e-12)
        assert np.allclose(Px, P @ uvec) and np.allclose(Aty, A.T @ yvec)
        assert np.allclose(dP, P.diagonal()) and np.allclose(dAA, np.asarray(A.multiply(A).sum(axis=0)).ravel())


_WORKER = textwrap.dedent('''
    import os, sys
    import numpy as np
    import torch, torch.distributed as dist
    sys.path.insert(0, os.environ["QPB_ROOT"])
    from quadraticprogramsolver_b200 import partition
    from workloads.problems import config_sparse
    from oracle import qp_oracle

    dist.init_process_group("gloo")
    rank, R = dist.get_rank(), dist.get_world_size()
    P, q, A, l, u = config_sparse(120, 260, 0.05, seed=11)
    n, m = 120, 260
    P_r, A_r, l_r, u_r, (i0, i1), (j0, j1) = partition.slice_problem(P, A, l, u, rank, R)
    A_rt = A_r.T.tocsr()

    def allreduce(v, op=dist.ReduceOp.SUM):
        t = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float64).copy())
        dist.all_reduce(t, op=op)
        return t.numpy()

    # distributed ADMM with the exact data flow of dist_kernels.cuh (numpy in place of the kernels)
    rho, sigma, alpha = 1.0, 1e-6, 1.6
    dinv = 1.0 / (allreduce(P_r.diagonal()) + sigma + rho * allreduce(np.asarray(A_r.multiply(A_r).sum(axis=0)).ravel()))
    x = np.zeros(n); xt = np.zeros(n); z = np.zeros(i1 - i0); y = np.zeros(i1 - i0); g = np.zeros(i1 - i0)
    flag, iters = 1, 0
    for ii in range(1, 2001):
        w = allreduce(P_r @ xt + A_rt @ g)                      # kSegBegin + all-reduce
        r = sigma * (x - xt) - q - w                            # kSegPcgInit
        zp = dinv * r; uvec = zp.copy()
        res = np.sqrt(r @ r); rz = r @ zp
        tol = max(np.sqrt(np.finfo(float).eps) * res, 1e-10)
        k = 0
        while k < 1000 and not res <= tol:
            w = allreduce(P_r @ uvec + A_rt @ (rho * (A_r @ uvec)))   # S2, S3 partial + all-reduce
            c = w + sigma * uvec                                # kSegPcgStep
            a = rz / (uvec @ c)
            xt += a * uvec; r -= a * c; zp = dinv * r
            res = np.sqrt(r @ r); rz_new = r @ zp; k += 1
            if k < 1000 and not res <= tol:
                uvec = zp + (rz_new / rz) * uvec
            rz = rz_new
        zt = A_r @ xt                                           # kSegUpdate (rows I_r only)
        x_old = x.copy(); x = alpha * xt + (1 - alpha) * x
        z_old = z.copy(); zr = alpha * zt + (1 - alpha) * z
        z = np.where(zr + y / rho > u_r, u_r, np.where(zr + y / rho < l_r, l_r, zr + y / rho))
        y = y + rho * (zr - z)
        g = rho * (zt - z) + y
        if ii % 25 == 0:
            ax = A_r @ x
            lmax = allreduce(np.array([np.max(np.abs(x - x_old)), np.max(np.abs(z - z_old), initial=0.0),
                                       np.max(np.abs(ax - z), initial=0.0),
                                       max(np.max(np.abs(ax), initial=0.0), np.max(np.abs(z), initial=0.0))]), dist.ReduceOp.MAX)
            w2 = allreduce(np.concatenate([P_r @ x, A_rt @ y]))
            px, aty = w2[:n], w2[n:]
            rd = np.max(np.abs(px + q + aty)); md = max(np.max(np.abs(px)), np.max(np.abs(aty)), np.max(np.abs(q)))
            if lmax[2] < 1e-6 + 1e-6 * lmax[3] and rd < 1e-6 + 1e-6 * md: flag = 3
            if lmax[0] <= 1e-8 and lmax[1] <= 1e-8: flag = 2
            if flag != 1:
                iters = ii
                break
    x_ref, flag_ref, info = qp_oracle.solve(P, q, A, l, u, mode="J", epsPcg=1e-10, numIterations=2000)
    assert flag == int(flag_ref), (flag, int(flag_ref))
    assert iters == info["iterations"], (iters, info["iterations"])
    assert np.max(np.abs(x - x_ref)) <= 1e-6 * (1 + np.max(np.abs(x_ref)))
    assert np.max(np.abs(z - info["z"][i0:i1])) <= 1e-6 * (1 + np.max(np.abs(info["z"])))
    print("rank", rank, "ok", iters, flush=True)
    dist.destroy_process_group()
''')


def test_two_rank_gloo_data_flow(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, QPB_ROOT=ROOT, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29613", str(script)],
                         env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("ok") == 2


_GATE_WORKER = textwrap.dedent('''
    import os, sys, time
    import torch.distributed as dist
    sys.path.insert(0, os.environ["QPB_ROOT"])
    import bench
    dist.init_process_group("gloo")
    rank = dist.get_rank()
    t0 = time.time()
    if rank == 0:
        time.sleep(1.5)                                   # "the CPU port runs"
        bench.quiet_gate(dist, "gate_test", True)
    else:
        c0 = time.process_time()
        bench.quiet_gate(dist, "gate_test", False)
        waited, cpu = time.time() - t0, time.process_time() - c0
        assert waited >= 1.2, waited                      # it did wait for rank 0 ...
        assert cpu <= 0.5, cpu                            # ... asleep, not spinning
    bench.quiet_gate(None, "gate_test", False)            # single process: no-op
    dist.barrier()
    print("rank", rank, "ok", flush=True)
    dist.destroy_process_group()
''')


def test_bench_quiet_gate_sleeps_until_rank0_opens_it(tmp_path):
    """bench.py's N > 1 parity leg: the other ranks sleep on the rendezvous store while rank 0 runs the CPU port."""
    script = tmp_path / "gate_worker.py"
    script.write_text(_GATE_WORKER)
    env = dict(os.environ, QPB_ROOT=ROOT, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29617", str(script)],
                         env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("ok") == 2


_SLICED_WORKER = textwrap.dedent('''
    import os, sys
    import numpy as np
    import torch, torch.distributed as dist
    sys.path.insert(0, os.environ["QPB_ROOT"])
    from quadraticprogramsolver_b200 import partition
    from workloads.problems import config_sparse
    from oracle import qp_oracle

    dist.init_process_group("gloo")
    rank, R = dist.get_rank(), dist.get_world_size()
    P, q, A, l, u = config_sparse(120, 260, 0.05, seed=11)
    n, m = 120, 260
    adaptive = os.environ.get("QPB_ADAPTIVE") == "1"
    cap = 3000 if adaptive else 200                         # fixed rho: the first 200 iterations (1200 to converge)
    P_r, A_r, l_r, u_r, (i0, i1), (j0, j1) = partition.slice_problem(P, A, l, u, rank, R)
    A_rt = A_r.T.tocsr()
    sb = [n * k // R for k in range(R + 1)]                 # vector slices S_r (dist_solver.cu: pd.sb)
    s0, s1 = sb[rank], sb[rank + 1]

    def allsum(v):
        t = torch.from_numpy(np.atleast_1d(np.asarray(v, dtype=np.float64)).copy())
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.numpy()

    def allmax(v):
        t = torch.from_numpy(np.atleast_1d(np.asarray(v, dtype=np.float64)).copy())
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.numpy()

    def reduce_scatter(partial):                            # push reduce-scatter: the owner of S_q sums R partials
        return allsum(partial)[s0:s1]

    def all_gather(slice_vals):                             # the owner pushes its slice into every rank's copy
        full = np.zeros(n); full[s0:s1] = slice_vals
        return allsum(full)

    # the data flow of admm_peer_sliced_kernel (peer_kernels.cuh) with numpy in place of the kernels
    rho, sigma, alpha, rho_factor = (0.1 if adaptive else 1.0), 1e-6, 1.6, 5.0
    eps_abs = eps_rel = 1e-6
    dPf = allsum(P_r.diagonal()); dAAf = allsum(np.asarray(A_r.multiply(A_r).sum(axis=0)).ravel())
    x = np.zeros(n); xt = np.zeros(n); z = np.zeros(i1 - i0); y = np.zeros(i1 - i0); zt = np.zeros(i1 - i0); g = np.zeros(i1 - i0)
    r = np.zeros(n); pvec = np.zeros(n); svec = np.zeros(n)  # only [s0:s1) of these is ever touched by this rank
    rhorho, flag, iters, rho_updates, cg_total = rho, 1, 0, 0, 0
    dinv = None
    for ii in range(1, cap + 1):
        iters = ii
        changed = False
        if adaptive and (rhorho * rho_factor < rho or rhorho > rho_factor * rho):
            rho = rhorho; changed = True; rho_updates += 1
        if changed or dinv is None:
            dinv = 1.0 / (dPf + sigma + rho * dAAf)
            if changed:
                g = rho * (zt - z) + y
        w = reduce_scatter(P_r @ xt + A_rt @ g)              # spmv_H_push(XG)
        r[s0:s1] = sigma * (x[s0:s1] - xt[s0:s1]) - q[s0:s1] - w
        zs = dinv[s0:s1] * r[s0:s1]
        zfull = all_gather(zs)                               # push_all(off_u, ...)
        d3 = allsum([r[s0:s1] @ zs, r[s0:s1] @ r[s0:s1], zs @ zs])
        residual = np.sqrt(d3[1]); tol = max(np.sqrt(np.finfo(float).eps) * residual, 1e-10)
        k, first, gam_prev, a_cg = 0, True, 0.0, 0.0
        while k < 1000 and not residual <= tol:
            partial = P_r @ zfull + A_rt @ (rho * (A_r @ zfull))   # t = rho A_r z ; H_r [z ; t]
            zw = allsum(zfull @ partial)[0]                  # z . sum_r w_r rides on the system barrier
            w = reduce_scatter(partial)
            gam, delta = d3[0], zw + sigma * d3[2]
            if first:
                assert delta > 0
                beta = 0.0; a_cg = gam / delta
            else:
                beta = gam / gam_prev
                den = delta - beta * gam / a_cg
                assert den > 0
                a_cg = gam / den
            gam_prev = gam
            zsl = zfull[s0:s1]
            pj = zsl + (0.0 if first else beta * pvec[s0:s1])
            sj = w + sigma * zsl + (0.0 if first else beta * svec[s0:s1])
            pvec[s0:s1] = pj; svec[s0:s1] = sj
            xt[s0:s1] += a_cg * pj
            r[s0:s1] -= a_cg * sj
            zs = dinv[s0:s1] * r[s0:s1]
            zfull = all_gather(zs)
            d3 = allsum([r[s0:s1] @ zs, r[s0:s1] @ r[s0:s1], zs @ zs])
            first = False
            residual = np.sqrt(d3[1]); k += 1
        cg_total += k
        xt = all_gather(xt[s0:s1])                           # all-gather x~
        zt = A_r @ xt                                        # row-local update (rows I_r only)
        z_old = z.copy(); zr = alpha * zt + (1 - alpha) * z
        z = np.minimum(np.maximum(zr + y / rho, l_r), u_r)
        y = y + rho * (zr - z)
        g = rho * (zt - z) + y
        x_old = x.copy(); x = alpha * xt + (1 - alpha) * x
        if ii % 25 == 0:
            ax = A_r @ x
            nrm = allmax([np.max(np.abs(x - x_old)), np.max(np.abs(z - z_old), initial=0.0), np.max(np.abs(ax - z), initial=0.0),
                          max(np.max(np.abs(ax), initial=0.0), np.max(np.abs(z), initial=0.0))])
            w2 = allsum(np.concatenate([P_r @ x, A_rt @ y]))
            px, aty = w2[:n], w2[n:]
            res_prim, res_dual = nrm[2], np.max(np.abs(px + q + aty))
            max_prim, max_dual = nrm[3], max(np.max(np.abs(px)), np.max(np.abs(aty)), np.max(np.abs(q)))
            if adaptive:
                rhorho = min(max(rho * np.sqrt((res_prim * max_dual) / (res_dual * max_prim)), 1e-3), 1e6)
            if res_prim < eps_abs + eps_rel * max_prim and res_dual < eps_abs + eps_rel * max_dual: flag = 3
            if nrm[0] <= 1e-8 and nrm[1] <= 1e-8: flag = 2
            if flag != 1:
                break
    kw = dict(epsPcg=1e-10, numIterations=cap, rho=0.1 if adaptive else 1.0, adptRho=adaptive)
    x_ref, flag_ref, info = qp_oracle.solve(P, q, A, l, u, mode="J", **kw)
    assert flag == int(flag_ref), (flag, int(flag_ref))
    assert iters == info["iterations"], (iters, info["iterations"])
    assert rho_updates == info["rho_updates"], (rho_updates, info["rho_updates"])
    assert np.max(np.abs(x - x_ref)) <= 1e-6 * (1 + np.max(np.abs(x_ref)))
    assert np.max(np.abs(z - info["z"][i0:i1])) <= 1e-6 * (1 + np.max(np.abs(info["z"])))
    print("rank", rank, "ok", iters, rho_updates, cg_total, flush=True)
    dist.destroy_process_group()
''')


import pytest  # noqa: E402


@pytest.mark.parametrize("adaptive", ["0", "1"])
def test_two_rank_gloo_sliced_one_reduction_cg_data_flow(tmp_path, adaptive):
    """The DEFAULT multi-GPU arrangement (admm_peer_sliced_kernel): vector slices per rank, reduce-scatter of the H
    partials, all-gather of z = Pl \\ r, gamma / delta / |r|^2 from fused reductions (Chronopoulos-Gear), the row-local
    update and the distributed CheckConvergence -- across a real 2-rank gloo group with numpy in place of the kernels,
    against the oracle (same flag, iteration count, rho updates, x, z)."""
    script = tmp_path / "sliced_worker.py"
    script.write_text(_SLICED_WORKER)
    env = dict(os.environ, QPB_ROOT=ROOT, OMP_NUM_THREADS="1", QPB_ADAPTIVE=adaptive)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29621" if adaptive == "0" else "29623", str(script)],
                         env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("ok") == 2
