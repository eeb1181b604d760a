mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_patterns.py -m gpu -x -q -k "while or patterns or golden" 2>&1 | tail -4
cd geosongpu-ci_b200
for v in 1 2; do
timeout 120 python -m b200stencil.bench.sweep --graph --iters 10 --stencils while_in_function --option while_variant=$v 2>&1 | tail -2 | cut -c1-330
done
timeout 120 python -m b200stencil.bench.sweep --graph --iters 10 --stencils while_in_function --config C180x72 --option while_variant=1 2>&1 | tail -2 | cut -c1-330
timeout 120 python -m b200stencil.bench.sweep --graph --iters 10 --stencils while_in_function --config C180x72 --option while_variant=2 2>&1 | tail -2 | cut -c1-330
