# A/B of the number of reference views in flight per GPU (bench.py --lanes): parity tests of the pipeline, then value / e2e per setting
mkdir -p gpurun_out/lanes
timeout 300 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_parity.py -q -x 2>&1 | tail -3 > gpurun_out/lanes/pytest.log
for L in ${LANES:-1 2 3 4}; do
  timeout 200 python bench.py --lanes $L --steps 20 --warmup 4 --no-train-step --no-eager --no-cpu-baseline --no-parity > gpurun_out/lanes/bench_$L.json 2> gpurun_out/lanes/bench_$L.err
  python -c "
import json; d=json.load(open('gpurun_out/lanes/bench_$L.json')); print('lanes=$L value', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],3), 'dense', round(d['e2e_dense']['value'],1), 'wait', round(d['e2e']['host_blocked_on_results_ms_per_step'],2), 'head', round(d['kernels']['head+schedule']['ms_per_step'],3))"
done
cat gpurun_out/lanes/pytest.log
