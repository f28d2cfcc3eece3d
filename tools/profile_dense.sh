#!/bin/bash
# Run on the GPU box (via gpurun): the configs[2] workload (one sequence, ~1e6-point live maps) plain, then the ncu launch list of
# steady-state frames and a --set full capture of the kernels of one frame.
# usage: tools/profile_dense.sh <tag> [preroll]
set -u
TAG=$1; PRE=${2:-200}
ARGS="--workload hdl64_dense --seqs 1 --groups 1 --preroll $PRE --steps 12 --warmup 3 --young-steps 0 --no-cpu --no-sweep --no-roofline --no-latency"
mkdir -p gpurun_out
python bench.py $ARGS > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.err; exit 1; }
LPF=$(python -c "import json;d=json.load(open('gpurun_out/${TAG}_plain.json'));print(int(round(d['gpu_launches']/d['steps'])))")
echo "launches per frame: $LPF"
SKIP=$((LPF * (PRE + 6) + 40))
ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c $((LPF * 3)) --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py $ARGS > gpurun_out/${TAG}_ncu1.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -s $SKIP -c $LPF -f -o gpurun_out/${TAG}_full python bench.py $ARGS > gpurun_out/${TAG}_ncu2.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out | tail -5
