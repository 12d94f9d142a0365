# Round-end GPU check used for profiles/: full GPU suite, smoke, bench line, ncu launch list and a
# --set full capture of the cost-volume kernels (summarised on the box; the .ncu-rep is dropped to stay
# under the gpurun_out size limit).  Run from the repo root on a B200: bash scripts/gpu_round_check.sh
mkdir -p gpurun_out/check
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 > gpurun_out/check/pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/check/smoke.log 2>&1
timeout 300 python bench.py > gpurun_out/check/bench.json 2> gpurun_out/check/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"cost_volume|vis_|conv3d|deconv3d|prob_conv|regression|schedule|init_|confidence|relproj|relative|to_cl|ncdhw|argmax" --csv --log-file gpurun_out/check/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train-step > gpurun_out/check/ncu_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:cost_volume_kernel -s 24 -c 8 -o gpurun_out/check/k1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train-step > gpurun_out/check/ncu_full.log 2>&1
python scripts/summarise_ncu.py gpurun_out/check/k1.ncu-rep gpurun_out/check/k1_full.csv cost_volume_kernel > gpurun_out/check/summ.log 2>&1
rm -f gpurun_out/check/k1.ncu-rep
cat gpurun_out/check/pytest.log gpurun_out/check/smoke.log; tail -c 600 gpurun_out/check/bench.json
