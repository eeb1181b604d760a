mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_halo.py -m gpu -x -q -k "split" 2>&1 | tail -12
cd geosongpu-ci_b200
for v in 1 2; do
timeout 200 python -m b200stencil.bench.sweep --stencils fv_tp2d_split --iters 10 --option fv_split_variant=$v 2>&1 | tail -2 | cut -c1-330
done
timeout 200 python -m b200stencil.bench.sweep --stencils fv_tp2d_split --iters 10 --dtypes f64 --option fv_split_variant=2 --option fv_split_ti=120 2>&1 | tail -1 | cut -c1-330
timeout 200 python -m b200stencil.bench.sweep --stencils fv_tp2d_split --iters 10 --dtypes f64 --option fv_split_variant=2 --option fv_split_jb=128 2>&1 | tail -1 | cut -c1-330
timeout 200 python -m b200stencil.bench.sweep --stencils fv_tp2d_split --iters 10 --dtypes f64 --option fv_split_variant=2 --option fv_split_jb=32 2>&1 | tail -1 | cut -c1-330
