#!/bin/sh
# The fingerprint bench.py asserts on every full-frontier run: tests/golden/vanderpol_bench_fingerprint.json.
# Runs the compiled, unmodified reference (oracle/_ref/ref_harness worker = UpperReachEvolveWorker::get_result(),
# include/workers.h:51-153) over the FULL bench frontier (van der Pol, accuracy 7, 101,520 cells, one lap) in P chunks
# (chunk j = frontier cells j, j+P, j+2P, ...; one process each, so the mutex-protected discretisation of the worker does
# not serialise the run) and fingerprints the union of the chunks' reach / evolve sets by their leaf cells.
# About 8 CPU-hours. Usage: tests/golden/make_bench_fingerprint.sh [processes]
set -e
cd "$(dirname "$0")/../.."
P=${1:-$(nproc)}
D=${TMPDIR:-/tmp}
j=0
while [ $j -lt $P ]; do
  ORACLE_THREADS=1 oracle/_ref/ref_harness vanderpol worker acc=7 ncells=1000000 stride=$P offset=$j init=-2.2,2.2,-2.8,2.8 dump=1 out=$D/vanderpol_bench_chunk_$j.txt &
  j=$((j+1))
done
wait
python tests/golden/bench_fingerprint.py tests/golden/vanderpol_bench_fingerprint.json $D/vanderpol_bench_chunk_*.txt
