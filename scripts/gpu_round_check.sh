# Round-end GPU check used for profiles/: full GPU suite, smoke, bench line (+ reference arm), ncu launch list with DRAM bytes
# and tensor-pipe activity of one steady-state step, and a --set full capture of the cost-volume kernels (summarised on the
# box; the .ncu-rep is dropped to stay under the gpurun_out size limit).  Run from the repo root on a B200:
#   bash scripts/gpu_round_check.sh
mkdir -p gpurun_out/check
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/check/pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/check/smoke.log 2>&1
timeout 400 python bench.py > gpurun_out/check/bench.json 2> gpurun_out/check/bench.err
timeout 200 python bench.py --steps 20 --warmup 3 --no-train-step --no-eager --no-cpu-baseline --no-parity > gpurun_out/check/bench_20steps.json 2> gpurun_out/check/bench_20steps.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/check/bench_reference.json 2> gpurun_out/check/bench_reference.err
KERN='mvs|tc::|k1cl|conv3d|vis_|corr_|cost_|tma3|prob_|regression|schedule|init_|confidence|relproj|argmax|nchw'
timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:"$KERN" -s 216 -c 170 --csv --log-file gpurun_out/check/launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train-step --no-eager --no-parity > gpurun_out/check/ncu_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"cost_volume_cl_kernel|nchw_to_cl|corr_aggregate" -s 27 -c 9 -o gpurun_out/check/k1 \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train-step --no-eager --no-parity > gpurun_out/check/ncu_full.log 2>&1
python scripts/summarise_ncu.py gpurun_out/check/k1.ncu-rep gpurun_out/check/k1_full.csv > gpurun_out/check/summ.log 2>&1
python scripts/traffic_from_launches.py gpurun_out/check/launches.csv gpurun_out/check/traffic.json >> gpurun_out/check/summ.log 2>&1
rm -f gpurun_out/check/k1.ncu-rep
cat gpurun_out/check/pytest.log gpurun_out/check/smoke.log; tail -c 400 gpurun_out/check/bench.json
