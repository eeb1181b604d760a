# round 2, run 25 (1 GPU): the driver's own N = 1 command at the final head
timeout 90 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_run25_bench_n1.json 2> gpurun_out/r2_run25_bench_n1.err
echo rc=$?; cut -c1-600 gpurun_out/r2_run25_bench_n1.json; tail -2 gpurun_out/r2_run25_bench_n1.err
