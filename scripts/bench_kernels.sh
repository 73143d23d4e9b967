#!/bin/bash
# Short bench + per-kernel table (under gpurun).  usage: scripts/bench_kernels.sh <tag> [bench args]
TAG=$1; shift
mkdir -p gpurun_out
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline "$@" > gpurun_out/bench_$TAG.log 2> gpurun_out/bench_$TAG.err || { tail -5 gpurun_out/bench_$TAG.err; exit 1; }
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_$TAG.log").read().strip().splitlines()[-1])
print("frames/s %.3f  ms/frame %.2f  e2e %.3f  iter_ms %.4f  frac %.4f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["iteration_avg_ms"], d["roofline"]["frac"]))
for k, v in d["kernels"].items():
    print("  %-14s %8.4f ms x %5.1f  share %.3f" % (k, v["ms_per_step"] / max(1, v["launches_per_step"]), v["launches_per_step"], v["share"]))
PY
