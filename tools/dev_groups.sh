for cfg in "32 1" "32 2" "32 4" "64 4" "64 8" "48 6"; do set -- $cfg
python bench.py --steps 60 --warmup 5 --seqs $1 --groups $2 --no-cpu --no-roofline > gpurun_out/g_$1_$2.json 2> gpurun_out/g_$1_$2.err || { echo "failed $cfg"; tail -3 gpurun_out/g_$1_$2.err; continue; }
python -c "
import json; d=json.load(open('gpurun_out/g_$1_$2.json')); print('S=$1 G=$2 value %.0f e2e %.0f ms/step %.3f same %s'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['e2e']['poses_equal_to_device_resident_run']))"
done
