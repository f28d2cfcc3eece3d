#!/bin/bash
# Run on the GPU box (via gpurun): GPU test suite, then short benches with the per-kernel table.
# usage: tools/run_gpu_check.sh [pytest-k-expr|all|none] [seqs...]
K=${1:-all}; shift
SEQS=${@:-"16 32"}
mkdir -p gpurun_out
if [ "$K" = "all" ]; then python -m pytest tests -m gpu -x -q 2>&1 | tail -15
elif [ "$K" != "none" ]; then python -m pytest tests -m gpu -x -q -k "$K" 2>&1 | tail -15; fi
for s in $SEQS; do
  python bench.py --steps 100 --warmup 5 --seqs $s --no-cpu --no-sweep > gpurun_out/c$s.json 2> gpurun_out/c$s.err || { echo "bench S=$s failed"; tail -5 gpurun_out/c$s.err; continue; }
  python - <<PY
import json
d=json.load(open("gpurun_out/c$s.json"))
print("S=$s value %.0f e2e %.0f ms/step %.3f launches %d"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["gpu_launches"]))
for k in d.get("kernels_ms",[]): print("   %-40s %8.2f ms %5d  %.1f us"%(k[0],k[1],k[2],1e3*k[1]/k[2]))
PY
done
