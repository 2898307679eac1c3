#!/bin/bash
# Build an A/B variant of libqpb200.so with extra nvcc flags (e.g. tile constants):
#   scripts/build_variant.sh t4096s2 "-DQPB_TILE_NNZ=4096 -DQPB_STAGES=2"
# -> quadraticprogramsolver_b200/variants/libqpb200_<name>.so ; select it with QPB200_LIB=<path>.
set -e
name=$1; flags=$2
root=$(cd "$(dirname "$0")/.." && pwd)
tmp=/tmp/qpb_build_$name
rm -rf "$tmp"; mkdir -p "$tmp/q/csrc" "$tmp/include" "$root/quadraticprogramsolver_b200/variants"
cp "$root"/quadraticprogramsolver_b200/csrc/*.cu "$root"/quadraticprogramsolver_b200/csrc/*.cuh "$root"/quadraticprogramsolver_b200/csrc/*.h \
   "$root"/quadraticprogramsolver_b200/csrc/*.cpp "$root"/quadraticprogramsolver_b200/csrc/Makefile "$tmp/q/csrc/"
cp "$root"/include/qpb200.h "$tmp/include/"
make -s -C "$tmp/q/csrc" EXTRA="$flags" OUT="$root/quadraticprogramsolver_b200/variants/libqpb200_$name.so" all
grep -E "spill|registers" "$tmp/q/csrc/sparse_solver.ptxas.log" | sort | uniq -c | sort -rn | head -6
