# Round-end GPU check used for profiles/: full GPU suite, smoke, the default bench line, a same-box A/B against the previous
# build of the library (build/ab/libmvs_b200_prev.so, when present), an ncu launch list with DRAM bytes / tensor-pipe activity /
# issued instructions of every launch, and a --set full capture of the cost-volume and visibility kernels of
# one steady-state step (summarised on the box; the .ncu-rep is dropped to stay under the gpurun_out size limit).
#   bash scripts/gpu_round_check.sh          (from the repo root, on a B200; REF_ARM=1 also runs the CPU reference arm)
mkdir -p gpurun_out/check
OUT=gpurun_out/check
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3 > $OUT/pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1
timeout 500 python bench.py > $OUT/bench.json 2> $OUT/bench.err
BENCH="--steps 20 --warmup 4 --no-train-step --no-eager --no-cpu-baseline --no-parity"
if [ -f build/ab/libmvs_b200_prev.so ]; then
  MVS_LIB_PATH=build/ab/libmvs_b200_prev.so timeout 200 python bench.py $BENCH > $OUT/bench_prev_build.json 2> $OUT/bench_prev_build.err
fi
timeout 200 python bench.py $BENCH > $OUT/bench_20steps.json 2> $OUT/bench_20steps.err
[ -n "$REF_ARM" ] && timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
KERN='mvs|tc::|k1cl|conv3d|vis_|corr_|cost_|tma3|prob_|regression|schedule|init_|confidence|relproj|argmax|nchw'
NOLEGS="--lanes 1 --no-cpu-baseline --no-train-step --no-eager --no-parity"
timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum \
  --clock-control none -k regex:"$KERN" -c 900 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 1 --warmup 3 $NOLEGS > $OUT/ncu_bench.log 2>&1
# per step: 5 cost-volume + 2 re-layout + 3 aggregation + 4 visibility launches = 14; skip three steps
timeout 400 ncu --set full --clock-control none --import-source on \
  -k regex:"cost_volume_cl_kernel|nchw_to_cl|corr_aggregate|vis_fused_kernel" -s 42 -c 14 -o $OUT/k1 \
  python bench.py --steps 1 --warmup 3 $NOLEGS > $OUT/ncu_full.log 2>&1
python scripts/summarise_ncu.py $OUT/k1.ncu-rep $OUT/k1_full.csv > $OUT/summ.log 2>&1
python scripts/traffic_from_launches.py $OUT/launches.csv $OUT/traffic.json >> $OUT/summ.log 2>&1
rm -f $OUT/k1.ncu-rep
cat $OUT/pytest.log $OUT/smoke.log; tail -c 600 $OUT/bench.json
for f in bench_prev_build bench_20steps; do python - $OUT/$f.json $f <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1])); k = d["kernels"]
    print(sys.argv[2], "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1),
          {n: round(v["ms_per_step"], 3) for n, v in k.items() if n.startswith("cv_")})
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
done
