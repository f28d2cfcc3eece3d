# usage: tools/dev_groups.sh "S G D" ...
for cfg in "$@"; do set -- $cfg
S=$1; G=$2; D=${3:-3}
python bench.py --steps 60 --warmup 5 --seqs $S --groups $G --depth $D --no-cpu --no-roofline --no-sweep > gpurun_out/g_${S}_${G}_${D}.json 2> gpurun_out/g_${S}_${G}_${D}.err || { echo "failed $cfg"; tail -3 gpurun_out/g_${S}_${G}_${D}.err; continue; }
python -c "
import json; d=json.load(open('gpurun_out/g_${S}_${G}_${D}.json')); print('S=$S G=$G D=$D value %.0f e2e %.0f ms/step %.3f / %.3f same %s'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['e2e']['ms_per_step'],d['e2e']['poses_equal_to_device_resident_run']))"
done
