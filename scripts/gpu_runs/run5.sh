mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --steps 200 --warmup 5 2>gpurun_out/bench_err.log | tee gpurun_out/bench_n1.json | cut -c1-3000
tail -5 gpurun_out/bench_err.log
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 | cut -c1-1500
timeout 300 python bench.py --steps 200 --warmup 5 --dtype f32 --skip-cpu | cut -c1-1500
