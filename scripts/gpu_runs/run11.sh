mkdir -p gpurun_out
cd geosongpu-ci_b200
CMD="python -m b200stencil.bench.sweep --stencils remap --config C384x72 --dtypes f64 --iters 3 --warmup 2 --option remap_variant=12"
$CMD > ../gpurun_out/plain_remap.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_remap -s 3 -c 1 -f -o ../gpurun_out/remap_r1 $CMD > ../gpurun_out/ncu_remap.log 2>&1
tail -2 ../gpurun_out/ncu_remap.log
