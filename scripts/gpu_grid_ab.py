"""cfg2 / cfg4 solve time vs persistent grid size (QPB200_GRID), plus cfg4 parity vs the compiled oracle."""
import json, os, subprocess, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    from quadraticprogramsolver_b200 import solver as S
    from workloads.problems import config_cfg2, config_cfg4
    out = {"grid": os.environ.get("QPB200_GRID", "auto")}
    for name, cfg, kw in (("cfg2", config_cfg2, dict(numIterations=1000)), ("cfg4", config_cfg4, dict(numIterations=300))):
        P, q, A, l, u = cfg()
        with S.QPB200Solver(P, q, A, l, u, **kw) as s:
            x = np.zeros(P.shape[0]); s.solve(x)
            x = np.zeros(P.shape[0]); s.solve(x)
            out[name] = dict(ms=round(s.info["solve_ms"], 2), it=s.info["iterations"], pcg=s.info["pcg_iters_total"],
                             us_per_pcg=round(1e3 * s.info["solve_ms"] / max(1, s.info["pcg_iters_total"]), 2))
    print(json.dumps(out), flush=True)
else:
    for g in ("auto", "74", "148", "296", "592"):
        env = dict(os.environ)
        if g != "auto":
            env["QPB200_GRID"] = g
        subprocess.run([sys.executable, __file__, "child"], env=env)
