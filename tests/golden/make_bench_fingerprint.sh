#!/bin/sh
# Runs the compiled, unmodified reference (oracle/_ref/ref_harness worker = UpperReachEvolveWorker::get_result(),
# include/workers.h:51-153) over the FULL bench frontier (van der Pol, accuracy 7, 101,520 cells, one lap) and writes
# the fingerprint bench.py asserts on every run: tests/golden/vanderpol_bench_fingerprint.json.
# About 30 min on 16 threads. Usage: tests/golden/make_bench_fingerprint.sh [threads]
set -e
cd "$(dirname "$0")/../.."
T=${1:-$(nproc)}
OUT=${TMPDIR:-/tmp}/vanderpol_bench_dump.txt
ORACLE_THREADS=$T oracle/_ref/ref_harness vanderpol worker acc=7 ncells=1000000 stride=1 init=-2.2,2.2,-2.8,2.8 dump=1 out=$OUT
python tests/golden/bench_fingerprint.py $OUT tests/golden/vanderpol_bench_fingerprint.json
