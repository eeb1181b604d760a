mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "fv_tp2d" 2>&1 | tail -3
cd geosongpu-ci_b200
for t in 0 2; do
  timeout 120 python -m b200stencil.bench.sweep --stencils fv_tp2d --iters 10 --option fv_tile=$t 2>&1 | tail -2 | cut -c1-420
done
CMD="python -m b200stencil.bench.sweep --stencils fv_tp2d --iters 3 --warmup 2"
$CMD > ../gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_fv_tma -s 4 -c 2 -o ../gpurun_out/fv_tma_r1 $CMD > ../gpurun_out/ncu.log 2>&1
tail -3 ../gpurun_out/ncu.log
