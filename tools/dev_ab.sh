# usage: tools/dev_ab.sh lib1.so lib2.so ...   (relative to vil_fusion_b200/): default bench config on each library
for lib in "$@"; do
  export VILF_LIB_PATH=$PWD/vil_fusion_b200/$lib
  python -m pytest tests -m gpu -x -q -k "golden or solve" 2>&1 | tail -1
  python bench.py --steps 60 --warmup 5 --no-cpu --no-sweep > gpurun_out/ab.json 2> gpurun_out/ab.err || { echo "$lib failed"; tail -3 gpurun_out/ab.err; continue; }
  python -c "
import json; d=json.load(open('gpurun_out/ab.json')); ks={k[0].split('/')[-1]:1e3*k[1]/k[2] for k in d['kernels_ms']}
print('$lib value %.0f e2e %.0f  k_solve %.1f us'%(d['value'],d['e2e']['value'],ks.get('k_solve',0)))"
done
