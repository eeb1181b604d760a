mkdir -p gpurun_out
cd geosongpu-ci_b200
timeout 400 python -m b200stencil.bench.sweep --iters 10 --out ../gpurun_out/r01_sweep_events.json > ../gpurun_out/sweep_events.log 2>&1; tail -1 ../gpurun_out/sweep_events.log | cut -c1-200
timeout 300 python -m b200stencil.bench.sweep --graph --iters 10 --stencils top_of_column,while_in_function,hybrid_index_2dout,find_klcl,cloud_top,saturation_adjust --out ../gpurun_out/r01_sweep_graph.json > ../gpurun_out/sweep_graph.log 2>&1; tail -1 ../gpurun_out/sweep_graph.log | cut -c1-200
timeout 300 python -m b200stencil.bench.sweep --iters 10 --config C384x72 --stencils top_of_column,while_in_function,hybrid_index_2dout,find_klcl,cloud_top,saturation_adjust,pe_prefix,remap,tridiag --out ../gpurun_out/r01_sweep_c384.json > ../gpurun_out/sweep_c384.log 2>&1; tail -1 ../gpurun_out/sweep_c384.log | cut -c1-200
cd ..
timeout 300 python bench.py > gpurun_out/r01_bench_n1_f64.json 2> gpurun_out/bench_f64.err; cut -c1-600 gpurun_out/r01_bench_n1_f64.json
timeout 300 python bench.py --dtype f32 > gpurun_out/r01_bench_n1_f32.json 2> gpurun_out/bench_f32.err; cut -c1-300 gpurun_out/r01_bench_n1_f32.json
timeout 300 python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/r01_bench_reference.json; cut -c1-300 gpurun_out/r01_bench_reference.json
lscpu | grep -E "Model name|^CPU\(s\)|Socket|Thread" > gpurun_out/host_cpu.txt; nvidia-smi --query-gpu=name,power.limit,clocks.max.sm,clocks.max.mem --format=csv >> gpurun_out/host_cpu.txt; cat gpurun_out/host_cpu.txt
