mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_halo.py tests/test_gpu_golden.py -m gpu -x -q -k "split or frozen_horizontal" 2>&1 | tail -5
cd geosongpu-ci_b200
for rp in 2 1; do
timeout 200 python -m b200stencil.bench.sweep --stencils fv_tp2d_split --iters 10 --option fv_split_rp=$rp 2>&1 | tail -2 | cut -c1-330
done
timeout 200 python -m b200stencil.bench.sweep --stencils fv_tp2d_split --iters 10 --dtypes f64 --option fv_split_ti=120 2>&1 | tail -1 | cut -c1-330
