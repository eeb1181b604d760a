mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r01_bench_n1_f64_b.json 2> gpurun_out/bench_err.log; tail -c 600 gpurun_out/bench_err.log; cut -c1-700 gpurun_out/r01_bench_n1_f64_b.json
cd geosongpu-ci_b200
timeout 200 python -m b200stencil.bench.sweep --stencils remap_ppm,fv_tp2d_split --iters 10 --out ../gpurun_out/r01_sweep_next_rows.json 2>&1 | cut -c1-300
timeout 200 python -m b200stencil.bench.sweep --stencils remap_ppm --iters 10 --config C384x72 --out ../gpurun_out/r01_sweep_ppm_c384.json 2>&1 | cut -c1-300
