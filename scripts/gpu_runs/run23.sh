mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "split or remap_ppm" 2>&1 | tail -12
cd geosongpu-ci_b200
for ti in 120 56; do
timeout 200 python -m b200stencil.bench.sweep --stencils fv_tp2d_split --iters 10 --option fv_split_ti=$ti 2>&1 | tail -2 | cut -c1-330
done
for c in 16 32; do
timeout 200 python -m b200stencil.bench.sweep --stencils remap_ppm --config C384x72 --iters 10 --option remap_ppm_cols=$c 2>&1 | tail -2 | cut -c1-330
done
