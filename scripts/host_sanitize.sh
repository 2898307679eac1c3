#!/bin/bash
# Host side of libqpb200.so under AddressSanitizer + UBSan (no GPU needed): builds a variant with the host compiler's
# sanitizers on and runs the CPU tests that call the library's host-only entry points (CSC -> tiled CSR conversion,
# equilibration, tile plan, multi-GPU partition and slicing, argument validation) against it.
#   bash scripts/host_sanitize.sh > profiles/r2_host_sanitize.log 2>&1
set -e
root=$(cd "$(dirname "$0")/.." && pwd)
bash "$root/scripts/build_variant.sh" asan "-Xcompiler -fsanitize=address -Xcompiler -fsanitize=undefined -Xcompiler -fno-omit-frame-pointer -g" > /dev/null
export QPB200_LIB="$root/quadraticprogramsolver_b200/variants/libqpb200_asan.so"
export LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)"
export ASAN_OPTIONS=detect_leaks=0 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1
cd "$root"
python -m pytest tests/test_host.py -q -p no:cacheprovider
rm -rf "$root/quadraticprogramsolver_b200/variants/libqpb200_asan.so"
