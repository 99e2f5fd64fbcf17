#!/bin/bash
# Host analysis (ordering, elimination tree, supernodes, plan), the static row permutation and the CPU plan interpreter
# under AddressSanitizer + UndefinedBehaviorSanitizer (compute-sanitizer is closed on this pool; this covers the host side).
set -e
cd "$(dirname "$0")/.."
OUT=${TMPDIR:-/tmp}/libnkp_sim_asan.so
g++ -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -fopenmp -shared -fPIC -std=c++17 -o "$OUT" \
    oracle/plan_sim.cpp nk_ocn_tracer_jacobian_precond_b200/csrc/analysis.cpp nk_ocn_tracer_jacobian_precond_b200/csrc/rowperm.cpp
NKP_ASAN_LIB="$OUT" ASAN_OPTIONS=detect_leaks=0 \
    LD_PRELOAD="$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so)" python scripts/asan_host_check.py
