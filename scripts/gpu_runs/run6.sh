mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
cd geosongpu-ci_b200
timeout 120 python -m b200stencil.bench.sweep --stencils fv_tp2d --iters 10 2>&1 | tail -2 | cut -c1-420
CMD="python -m b200stencil.bench.sweep --stencils fv_tp2d --iters 3 --warmup 2"
$CMD > ../gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_fv_tma -s 4 -c 2 -o ../gpurun_out/fv_tma_r1 -f $CMD > ../gpurun_out/ncu.log 2>&1
tail -2 ../gpurun_out/ncu.log
cd ..
CMD2="python bench.py --steps 3 --warmup 3 --skip-cpu --skip-e2e"
$CMD2 > gpurun_out/plain2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_bench_n1.csv $CMD2 > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log
