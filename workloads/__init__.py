"""Input generators for tests, bench.py and the study scripts (NOT part of the product package):
a cited restatement of the reference's GenerateRandomQP (GenerateQuadraticProgram.jl:8-115) and the
five BASELINE.json configurations built on it."""
from .problems import GenerateRandomQP, ProblemClass  # noqa: F401
