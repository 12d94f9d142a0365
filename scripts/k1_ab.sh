# A/B of the channels-last cost-volume kernel variants (env MVS_K1_S3 / MVS_K1_S4): parity tests, then per-kernel-class times
mkdir -p gpurun_out/k1ab
for V in "0 0" "1 0" "0 1" "0 2"; do
  set -- $V
  export MVS_K1_S3=$1 MVS_K1_S4=$2
  timeout 300 python -m pytest tests/test_gpu_variants.py tests/test_gpu_parity.py -q -x -k "channels_last or cost_volume or stagenet or cascade" 2>&1 | tail -1
  timeout 200 python bench.py --steps 10 --warmup 3 --no-train-step --no-eager --no-cpu-baseline --no-parity > gpurun_out/k1ab/bench_$1_$2.json 2> gpurun_out/k1ab/bench_$1_$2.err
  python -c "
import json; d=json.load(open('gpurun_out/k1ab/bench_$1_$2.json')); print('S3=$1 S4=$2', round(d['ms_per_step'],3), {k: round(v['ms_per_step'],3) for k,v in d['kernels'].items() if k.startswith('cv_')})"
done
