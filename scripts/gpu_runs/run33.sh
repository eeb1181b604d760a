mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_c_abi_driver.py -m gpu -x -q -k "fv_tp2d and not split or compiled" 2>&1 | tail -5
cd geosongpu-ci_b200
for opt in "fv_variant=2" "fv_variant=3" "fv_variant=3 --option fv_stages=2" "fv_variant=3 --option fv_stages=4" "fv_variant=3 --option fv_jb=192" "fv_variant=3 --option fv_jb=384" "fv_variant=3 --option fv_jb=64"; do
timeout 200 python -m b200stencil.bench.sweep --stencils fv_tp2d --iters 20 --dtypes f64 --option $opt 2>&1 | tail -1 | cut -c100-420
done
timeout 200 python -m b200stencil.bench.sweep --stencils fv_tp2d --iters 20 --dtypes f32 --option fv_variant=3 2>&1 | tail -1 | cut -c100-420
timeout 200 python -m b200stencil.bench.sweep --stencils fv_tp2d --iters 20 --dtypes f32 --option fv_variant=2 2>&1 | tail -1 | cut -c100-420
