mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 200 python bench.py --workload chain --steps 20 --warmup 3 > gpurun_out/r01_chain_n1_f64.json; cut -c1-250 gpurun_out/r01_chain_n1_f64.json; grep -o '"roofline".*' gpurun_out/r01_chain_n1_f64.json | cut -c1-200
timeout 200 python bench.py --workload chain --steps 20 --warmup 3 --fused-remap > gpurun_out/r01_chain_n1_f64_fused.json; grep -o '"ms_per_step[^,]*' gpurun_out/r01_chain_n1_f64_fused.json; grep -o '"roofline".*' gpurun_out/r01_chain_n1_f64_fused.json | cut -c1-200
timeout 200 python bench.py --workload chain --steps 20 --warmup 3 --fused-remap --dtype f32 > gpurun_out/r01_chain_n1_f32_fused.json; grep -o '"ms_per_step[^,]*' gpurun_out/r01_chain_n1_f32_fused.json
timeout 300 python bench.py --hws-dump gpurun_out/r01_hws_bench_n1 > gpurun_out/r01_bench_n1_f64.json 2>gpurun_out/bench_f64.err; cut -c1-200 gpurun_out/r01_bench_n1_f64.json; grep -o '"hws".*"gpu_launches"' gpurun_out/r01_bench_n1_f64.json; ls -la gpurun_out/r01_hws_bench_n1.npz
cd geosongpu-ci_b200 && timeout 100 python -m b200stencil.bench.sweep --stencils remap --iters 5 2>&1 | tail -2 | cut -c1-330
