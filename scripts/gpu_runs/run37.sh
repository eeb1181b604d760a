mkdir -p gpurun_out
timeout 200 python bench.py --workload chain --fused-remap > gpurun_out/r01_chain_n1_f64_fused_stream.json 2> gpurun_out/chain.err; tail -c 300 gpurun_out/chain.err; cut -c1-900 gpurun_out/r01_chain_n1_f64_fused_stream.json
timeout 200 python bench.py --workload chain --fused-remap --dtype f32 > gpurun_out/r01_chain_n1_f32_fused_stream.json 2>> gpurun_out/chain.err; cut -c1-400 gpurun_out/r01_chain_n1_f32_fused_stream.json
cd geosongpu-ci_b200
timeout 300 python -m b200stencil.bench.sweep --stencils fv_tp2d,fv_tp2d_split,remap_ppm --iters 20 --out ../gpurun_out/r01_sweep_events_late.json 2>&1 | cut -c1-60,150-330
