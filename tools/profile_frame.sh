#!/bin/bash
# Run on the GPU box (via gpurun): plain run first, then the ncu launch list, then a --set full capture of ONE whole
# steady-state frame (every kernel of the launch sequence once).  One stream group, so launches are not interleaved.
# usage: tools/profile_frame.sh <tag> <seqs> <launches_per_frame> [frame_to_capture]
set -u
TAG=$1; S=$2; LPF=$3; FR=${4:-20}
ARGS="--steps 25 --warmup 3 --seqs $S --groups 1 --no-cpu --no-sweep --no-roofline"
mkdir -p gpurun_out
python bench.py $ARGS > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.err; exit 1; }
SKIP=$((LPF * FR))
ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c $((LPF * 4)) --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py $ARGS > gpurun_out/${TAG}_ncu1.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -s $SKIP -c $LPF -f -o gpurun_out/${TAG}_full python bench.py $ARGS > gpurun_out/${TAG}_ncu2.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out | tail -5
