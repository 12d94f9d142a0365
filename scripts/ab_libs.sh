# A/B of two builds of libmvs_b200.so on one box, back to back (MVS_LIB_PATH selects the build; see mvsformer_b200/_lib.py):
#   build/ab/libmvs_b200_base.so   the previous kernel generation (copied there before the change)
#   mvsformer_b200/lib/...         the current build
# Order: the whole GPU suite on the current build (parity gate), then bench.py per build, then an ncu launch list of one
# steady-state step and a --set full capture of the visibility and cost-volume kernels of the current build.
#   bash scripts/ab_libs.sh            (from the repo root, on a B200)
mkdir -p gpurun_out/ab
OUT=gpurun_out/ab
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -15 > $OUT/pytest.log
BENCH="--steps 20 --warmup 4 --no-train-step --no-eager --no-cpu-baseline --no-parity"
run() {  # name, extra bench flags, env...
  name=$1; flags=$2; shift; shift
  env "$@" timeout 200 python bench.py $BENCH $flags > $OUT/bench_$name.json 2> $OUT/bench_$name.err
  python - "$OUT/bench_$name.json" "$name" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    k = d.get("kernels", {})
    print(sys.argv[2], "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1),
          {n: round(v["ms_per_step"], 3) for n, v in k.items()})
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
if ! grep -q " passed" $OUT/pytest.log || grep -q "failed" $OUT/pytest.log; then     # is the pipelined conv epilogue the cause?
  MVS_TMA_EPI_PIPE=0 timeout 300 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -8 > $OUT/pytest_nopipe.log
fi
for lib in base mid; do
  if [ -f build/ab/libmvs_b200_$lib.so ]; then
    run $lib "" MVS_LIB_PATH=build/ab/libmvs_b200_$lib.so MVS_PROB_FUSED=$([ $lib = base ] && echo 0 || echo 1)
  fi
done
run new_nopipe "" MVS_TMA_EPI_PIPE=0
run new "" MVS_TMA_EPI_PIPE=1
run new_lanes1 "--lanes 1" MVS_TMA_EPI_PIPE=1
KERN='mvs|tc::|k1cl|conv3d|vis_|corr_|cost_|tma3|prob_|regression|schedule|init_|confidence|relproj|argmax|nchw'
timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__inst_executed_pipe_fma.sum \
  --clock-control none -k regex:"$KERN" -c 900 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 1 --warmup 3 --lanes 1 --no-cpu-baseline --no-train-step --no-eager --no-parity > $OUT/ncu_bench.log 2>&1
[ -n "$AB_FULL" ] && timeout 400 ncu --set full --clock-control none --import-source on -k regex:"cost_volume_cl_kernel|vis_fused_kernel" -s 27 -c 9 -o $OUT/k1vis \
  python bench.py --steps 1 --warmup 3 --lanes 1 --no-cpu-baseline --no-train-step --no-eager --no-parity > $OUT/ncu_full.log 2>&1
[ -n "$AB_FULL" ] && python scripts/summarise_ncu.py $OUT/k1vis.ncu-rep $OUT/k1vis_full.csv > $OUT/summ.log 2>&1
rm -f $OUT/k1vis.ncu-rep
cat $OUT/pytest.log $OUT/pytest_nopipe.log 2>/dev/null
