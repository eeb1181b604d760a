mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -6
cd geosongpu-ci_b200
timeout 400 python -m b200stencil.bench.sweep --iters 10 --out ../gpurun_out/r01_sweep_events.json > ../gpurun_out/sweep_events.log 2>&1; tail -1 ../gpurun_out/sweep_events.log | cut -c1-200
timeout 300 python -m b200stencil.bench.sweep --graph --iters 10 --stencils top_of_column,while_in_function,hybrid_index_2dout,find_klcl,cloud_top,saturation_adjust --out ../gpurun_out/r01_sweep_graph.json > ../gpurun_out/sweep_graph.log 2>&1; tail -1 ../gpurun_out/sweep_graph.log | cut -c1-200
timeout 300 python -m b200stencil.bench.sweep --iters 10 --config C384x72 --stencils top_of_column,while_in_function,hybrid_index_2dout,find_klcl,cloud_top,saturation_adjust,pe_prefix,remap,remap_delp,tridiag --out ../gpurun_out/r01_sweep_c384.json > ../gpurun_out/sweep_c384.log 2>&1; tail -1 ../gpurun_out/sweep_c384.log | cut -c1-200
CMD="python -m b200stencil.bench.sweep --stencils remap --dtypes f64 --iters 3 --warmup 1"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_remap_slab -s 2 -c 1 -f -o ../gpurun_out/remap_slab_r1 $CMD > ../gpurun_out/ncu_remap_slab.log 2>&1; tail -2 ../gpurun_out/ncu_remap_slab.log
CMD="python -m b200stencil.bench.sweep --stencils saturation_adjust --dtypes f64 --iters 3 --warmup 1"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_saturation -s 2 -c 1 -f -o ../gpurun_out/sat_r1 $CMD > ../gpurun_out/ncu_sat.log 2>&1; tail -2 ../gpurun_out/ncu_sat.log
cd ..
timeout 300 python bench.py > gpurun_out/r01_bench_n1_f64.json 2> gpurun_out/bench_f64.err; cut -c1-300 gpurun_out/r01_bench_n1_f64.json
timeout 200 python bench.py --workload chain --steps 20 --warmup 3 --fused-remap > gpurun_out/r01_chain_n1_f64_fused.json; grep -o '"ms_per_step[^,]*' gpurun_out/r01_chain_n1_f64_fused.json
timeout 200 python bench.py --workload chain --steps 20 --warmup 3 > gpurun_out/r01_chain_n1_f64.json; grep -o '"ms_per_step[^,]*' gpurun_out/r01_chain_n1_f64.json
timeout 200 python bench.py --workload chain --steps 20 --warmup 3 --fused-remap --dtype f32 > gpurun_out/r01_chain_n1_f32_fused.json; grep -o '"ms_per_step[^,]*' gpurun_out/r01_chain_n1_f32_fused.json
