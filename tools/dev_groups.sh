for cfg in "$@"; do set -- $cfg
S=${cfg% *}; G=${cfg#* }
python bench.py --steps 60 --warmup 5 --seqs $S --groups $G --no-cpu --no-roofline > gpurun_out/g_${S}_${G}.json 2> gpurun_out/g_${S}_${G}.err || { echo "failed $cfg"; tail -3 gpurun_out/g_${S}_${G}.err; continue; }
python -c "
import json; d=json.load(open('gpurun_out/g_${S}_${G}.json')); print('S=$S G=$G value %.0f e2e %.0f ms/step %.3f same %s'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['e2e']['poses_equal_to_device_resident_run']))"
done
