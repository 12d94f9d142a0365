# A/B of the stage-4 tiling of the channels-last cost-volume kernels (MVS_K1_STAGE4=b: 2 hypotheses per thread, 3 CTAs per SM):
# cost-volume parity tests with the candidate, then bench.py with each tiling, same box, back to back.
mkdir -p gpurun_out/k1s4
OUT=gpurun_out/k1s4
MVS_K1_STAGE4=b timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pipeline.py -m gpu -q 2>&1 | tail -5 > $OUT/pytest_b.log
BENCH="--steps 20 --warmup 4 --no-train-step --no-eager --no-cpu-baseline --no-parity"
for v in a b a b; do
  MVS_K1_STAGE4=$v timeout 200 python bench.py $BENCH > $OUT/bench_$v.json 2> $OUT/bench_$v.err
  python - $OUT/bench_$v.json $v <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1])); k = d["kernels"]
    print(sys.argv[2], "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1),
          "s4A", round(k["cv_cl_passA(stage4)"]["ms_per_step"], 3), "s4B", round(k["cv_cl_passB(stage4)"]["ms_per_step"], 3))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
done
cat $OUT/pytest_b.log
