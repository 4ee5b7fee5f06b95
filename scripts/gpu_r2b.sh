#!/bin/bash
# Round-2 GPU session: full parity suite (incl. config-size and kernel-vs-reference tests), then the default bench line.
TAG=${1:-r2b}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
which g++ /usr/bin/g++; nproc
python -m pytest tests -m gpu -q -x --timeout 1500 -s 2>&1 | tail -150 | tee gpurun_out/pytest_gpu_$TAG.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 6000 gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
