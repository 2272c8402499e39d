#!/bin/bash
# smoke, full GPU test suite, default bench (run from the repo root)
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -6 | cut -c1-200
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench exit $?"; cut -c1-2200 gpurun_out/bench_final.json; tail -3 gpurun_out/bench_final.err
