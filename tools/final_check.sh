set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tee gpurun_out/r2_gpu_suite_1gpu.log | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_n1_final.json 2> gpurun_out/r2_bench_n1_final.err; tail -c 600 gpurun_out/r2_bench_n1_final.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r2_bench_ref_n1_final.json 2> gpurun_out/r2_bench_ref_n1_final.err; tail -c 400 gpurun_out/r2_bench_ref_n1_final.json
timeout 900 python tools/bench_configs.py > gpurun_out/r2_configs_final.jsonl 2> gpurun_out/r2_configs_final.err; cut -c1-300 gpurun_out/r2_configs_final.jsonl
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_launches_bench.log 2>&1; tail -2 gpurun_out/r2_launches_bench.log | cut -c1-300
