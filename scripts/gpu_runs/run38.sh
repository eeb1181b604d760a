mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
cd geosongpu-ci_b200
timeout 200 python -m b200stencil.bench.sweep --stencils fv_tp2d --iters 20 --out ../gpurun_out/r01_sweep_fv_final.json 2>&1 | cut -c1-60,150-330
timeout 200 python -m b200stencil.bench.sweep --stencils fv_tp2d --iters 20 --config C720x137 2>&1 | cut -c1-60,150-330
for sub in 384,192,3,72 384,384,3,72; do timeout 100 python -m b200stencil.bench.sweep --stencils fv_tp2d --iters 20 --dtypes f64 --graph --sub $sub 2>&1 | tail -1 | cut -c1-60,150-330; done
cd ..
timeout 600 python bench.py > gpurun_out/r01_bench_n1_f64_final.json 2> gpurun_out/bench_err.log; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r01_bench_n1_f64_final.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','roofline','clocks','gpu_launches')})
PY
