"""Generates the golden fixtures of tests/golden/*.npz.

The reference is Julia (not runnable here) and ships no golden vectors, so these fixtures are outputs of
the ORACLE (oracle/qp_oracle.py), not of the reference: they pin the oracle against regressions and give
the GPU parity tests a committed input/output set that does not depend on scipy/numpy RNG stability.
Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import qp_oracle  # noqa: E402
from workloads.problems import GenerateRandomQP, ProblemClass  # noqa: E402

CASES = [
    # name, class, n, m, seed, mode, kwargs
    ("randomqp_n100_default_M", ProblemClass.randomQp, 100, 0, 1234, "M", dict(epsPcg=1e-10)),
    ("randomqp_n100_default_J", ProblemClass.randomQp, 100, 0, 1234, "J", dict(epsPcg=1e-10)),
    ("randomqp_n100_runtests_J", ProblemClass.randomQp, 100, 0, 1235, "J",
     dict(numIterations=50000, epsAbs=1e-7, epsRel=1e-7, rho=0.1, adptRho=True, epsPcg=1e-11)),
    ("ineq_n10_runtests_J", ProblemClass.inequalityConstrainedQp, 10, 0, 1234, "J",
     dict(numIterations=50000, epsAbs=1e-7, epsRel=1e-7, rho=0.1, adptRho=True, epsPcg=1e-11)),
    ("eq_n100_default_M", ProblemClass.equalityConstrainedQp, 100, 50, 1234, "M", dict(epsPcg=1e-10)),
    ("portfolio_n100_default_J", ProblemClass.portfolioOptimization, 100, 0, 1234, "J", dict(epsPcg=1e-10)),
    ("lasso_n10_default_J", ProblemClass.lassoOptimization, 10, 0, 1234, "J", dict(epsPcg=1e-10)),
    ("svm_n10_default_J", ProblemClass.supportVectorMachine, 10, 0, 1234, "J", dict(epsPcg=1e-10)),
    ("isotonic_n100_default_J", ProblemClass.isotonicRegression, 100, 0, 1234, "J", dict(epsPcg=1e-10)),
    # SURVEY 8(f) row 1 (not in the reference): Ruiz equilibration, per-constraint rho (1e3 on the equality rows)
    ("randomqp_n100_scaled_J", ProblemClass.randomQp, 100, 0, 1236, "J",
     dict(numIterations=3000, rho=0.1, adptRho=True, epsPcg=1e-11, numItrScaling=10)),
    ("eq_n100_rhoeq_J", ProblemClass.equalityConstrainedQp, 100, 50, 3, "J",
     dict(numIterations=3000, rho=0.1, epsPcg=1e-10, rhoEqScale=1e3)),
]


def oracle_kwargs(kw, l, u):
    """Fixture kwargs -> qp_oracle.solve kwargs (``rhoEqScale`` is the mirrors' shorthand for a ``rhoScale`` vector)."""
    kw = dict(kw)
    if "rhoEqScale" in kw:
        kw["rhoScale"] = np.where(np.asarray(l) == np.asarray(u), float(kw.pop("rhoEqScale")), 1.0)
    return kw


if __name__ == "__main__":
    only = set(sys.argv[1:])                     # optional: names of the fixtures to (re)generate
    for name, pc, n, m, seed, mode, kw in CASES:
        if only and name not in only:
            continue
        P, q, A, l, u = GenerateRandomQP(pc, n, numConstraints=m, seed=seed)
        x, flag, info = qp_oracle.solve(P, q, A, l, u, mode=mode, **oracle_kwargs(kw, l, u))
        out = dict(P_data=P.data, P_indices=P.indices, P_indptr=P.indptr, P_shape=np.array(P.shape),
                   A_data=A.data, A_indices=A.indices, A_indptr=A.indptr, A_shape=np.array(A.shape),
                   q=q, l=l, u=u, x=x, z=info["z"], y=info["y"], flag=int(flag), iterations=info["iterations"],
                   cg_iters=info["cg_iters"], rho=info["rho"], mode=mode)
        out.update({"kw_" + k: v for k, v in kw.items()})
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, P.shape, A.shape, int(flag), info["iterations"], info["cg_iters"])
