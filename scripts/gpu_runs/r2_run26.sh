# round 2, run 26 (1 GPU, the round's last GPU seconds): ncu launch list of the contract command with the final defaults
# (step = handshake kernel + flat-grid pull + stencil); bench.py exited 0 without ncu in run 25
timeout 62 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 150 --csv --log-file gpurun_out/r02_bench_n1_ncu_launches.csv python bench.py --steps 4 --warmup 3 --skip-cpu --skip-e2e > gpurun_out/r2_run26_ncu.log 2>&1
echo rc=$?; wc -l gpurun_out/r02_bench_n1_ncu_launches.csv; tail -2 gpurun_out/r2_run26_ncu.log | cut -c1-300
