"""quadraticprogramsolver_b200 -- B200-native (sm_100a) drop-in for the OSQP-style ADMM hot path of
RoyiAvital/QuadraticProgramSolver.  See DESIGN.md / INTEGRATION.md."""

def __getattr__(name):
    # solver entry points load libqpb200.so lazily
    if name in ("SolveQuadraticProgram", "SolveQuadraticProgram_", "QPB200Solver", "ConvergenceFlag", "B200Init",
                "B200Sol", "QPB200Error", "make_settings", "SolveQuadraticProgramBatch", "QPB200Batch"):
        from . import solver
        return getattr(solver, name)
    raise AttributeError(name)
