# round 2, run 10 (1 GPU): has the tile / streaming kernel slowed down at the 8-GPU sub-domain shape? (r1: 61.0 / 63.1 us)
export PYTHONPATH=geosongpu-ci_b200
for v in 2 3; do
timeout 300 python -m b200stencil.bench.sweep --stencils fv_tp2d --dtypes f64 --sub 192,192,3,72 --graph --option fv_variant=$v 2>&1 | tail -1
timeout 300 python -m b200stencil.bench.sweep --stencils fv_tp2d --dtypes f64 --sub 192,192,3,72 --option fv_variant=$v 2>&1 | tail -1
done
timeout 300 python -m b200stencil.bench.sweep --stencils fv_tp2d --dtypes f64,f32 --config C384x72 --graph 2>&1 | tail -2
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit,clocks_throttle_reasons.active --format=csv
