mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "fv_tp2d" 2>&1 | tail -15
cd geosongpu-ci_b200
for t in 0 1 2 3 4 5 6; do
  timeout 120 python -m b200stencil.bench.sweep --stencils fv_tp2d --iters 10 --option fv_tile=$t 2>&1 | tail -2 | cut -c1-420
done
timeout 120 python -m b200stencil.bench.sweep --stencils fv_tp2d --iters 10 --option fv_variant=1 2>&1 | tail -2 | cut -c1-420
timeout 120 python -m b200stencil.bench.sweep --stencils fv_tp2d --iters 10 --config C720x137 --dtypes f64 2>&1 | tail -1 | cut -c1-420
