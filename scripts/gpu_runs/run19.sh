mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
cd geosongpu-ci_b200
for kc in 4 8 12; do for un in 1 2; do
timeout 120 python -m b200stencil.bench.sweep --stencils saturation_adjust --dtypes f64 --iters 10 --option sat_kchunk=$kc --option sat_unroll=$un 2>&1 | tail -1 | cut -c1-400
done; done
timeout 120 python -m b200stencil.bench.sweep --stencils saturation_adjust --config C384x72 --iters 10 2>&1 | tail -2 | cut -c1-400
cd ..
timeout 200 python bench.py --workload chain --steps 20 --warmup 3 --fused-remap > gpurun_out/r01_chain_n1_f64_fused_slab.json; grep -o '"ms_per_step[^,]*' gpurun_out/r01_chain_n1_f64_fused_slab.json; grep -o '"roofline".*' gpurun_out/r01_chain_n1_f64_fused_slab.json | cut -c1-200
