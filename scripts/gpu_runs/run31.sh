mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_fullsize.py -m gpu -x -q -k "ppm or frozen_vertical" 2>&1 | tail -3
cd geosongpu-ci_b200
timeout 200 python -m b200stencil.bench.sweep --stencils remap_ppm --iters 5 --config C384x72 2>&1 | tail -2 | cut -c1-330
timeout 200 python -m b200stencil.bench.sweep --stencils remap_ppm --iters 5 2>&1 | tail -2 | cut -c1-330
