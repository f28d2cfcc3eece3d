#!/bin/bash
# Run on the GPU box (via gpurun): plain run first, then the ncu launch list and one full capture of a kernel.
# usage: tools/profile.sh <tag> <kernel-regex> [bench args...]
set -u
TAG=$1; KRE=$2; shift 2
ARGS=${@:-"--steps 12 --warmup 3 --seqs 4 --no-cpu --no-sweep --no-roofline"}
mkdir -p gpurun_out
python bench.py $ARGS > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py $ARGS > gpurun_out/${TAG}_ncu1.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:${KRE} -s 6 -c 3 -f -o gpurun_out/${TAG}_full python bench.py $ARGS > gpurun_out/${TAG}_ncu2.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out | tail -8
