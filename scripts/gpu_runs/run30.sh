mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_halo.py tests/test_gpu_golden.py tests/test_gpu_fullsize.py -m gpu -x -q -k "split or ppm or frozen" 2>&1 | tail -5
cd geosongpu-ci_b200
timeout 200 python -m b200stencil.bench.sweep --stencils fv_tp2d_split --iters 10 2>&1 | tail -2 | cut -c1-330
timeout 200 python -m b200stencil.bench.sweep --stencils remap_ppm --iters 5 --config C384x72 2>&1 | tail -2 | cut -c1-330
timeout 200 python -m b200stencil.bench.sweep --stencils remap_ppm --iters 5 2>&1 | tail -2 | cut -c1-330
CMD="python -m b200stencil.bench.sweep --stencils remap_ppm --iters 5 --dtypes f64 --config C384x72"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_remap_ppm -s 2 -c 1 -f -o ../gpurun_out/remap_ppm_r1b $CMD > ../gpurun_out/ncu_ppm.log 2>&1; tail -2 ../gpurun_out/ncu_ppm.log
