mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "remap" 2>&1 | tail -8
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "(remap_ppm and shape0 and float64) or (remap_ppm and shape4 and float32 and 4-1) or (remap_variants and shape0 and float64) or remap_ppm_unaligned" 2>&1 | tail -6
cd geosongpu-ci_b200
for c in 16 32; do
timeout 200 python -m b200stencil.bench.sweep --stencils remap_ppm --iters 10 --option remap_ppm_cols=$c 2>&1 | tail -2 | cut -c1-330
timeout 200 python -m b200stencil.bench.sweep --stencils remap_ppm --config C384x72 --iters 10 --option remap_ppm_cols=$c 2>&1 | tail -2 | cut -c1-330
done
timeout 200 python -m b200stencil.bench.sweep --stencils remap --iters 10 --dtypes f32 2>&1 | tail -1 | cut -c1-330
