mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_golden.py tests/test_gpu_halo.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "split or ppm" 2>&1 | tail -3
timeout 200 python scripts/halo_bench.py 384 72 2>&1 | tee gpurun_out/r01_halo_bench.jsonl | cut -c1-200
cd geosongpu-ci_b200
CMD="python -m b200stencil.bench.sweep --stencils fv_tp2d_split --iters 10 --dtypes f64"
timeout 200 $CMD 2>&1 | tail -1 | cut -c1-330
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fv_split_stream -s 2 -c 1 -f -o ../gpurun_out/fv_split_stream_r1 $CMD > ../gpurun_out/ncu_split.log 2>&1; tail -2 ../gpurun_out/ncu_split.log
CMD="python -m b200stencil.bench.sweep --stencils remap_ppm --iters 5 --dtypes f64 --config C384x72"
timeout 200 $CMD 2>&1 | tail -1 | cut -c1-330
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_remap_ppm -s 2 -c 1 -f -o ../gpurun_out/remap_ppm_r1 $CMD > ../gpurun_out/ncu_ppm.log 2>&1; tail -2 ../gpurun_out/ncu_ppm.log
