mkdir -p gpurun_out
CMD2="python bench.py --steps 3 --warmup 3 --skip-cpu --skip-e2e"
$CMD2 > gpurun_out/plain2.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 30 --csv --log-file gpurun_out/launches_bench_n1.csv $CMD2 > gpurun_out/ncu2.log 2>&1
grep -E "k_halo|k_fv" gpurun_out/launches_bench_n1.csv | awk -F'","' '{print $5, $(NF-2), $NF}' | cut -c1-200 | tail -12
