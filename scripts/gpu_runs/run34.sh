mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r01_bench_n1_f64_stream.json 2> gpurun_out/bench_err.log; tail -c 300 gpurun_out/bench_err.log; python - <<'PY'
import json
d=json.load(open('gpurun_out/r01_bench_n1_f64_stream.json'))
print({k:d[k] for k in ('value','ms_per_step','roofline','clocks')})
PY
timeout 300 python bench.py --dtype f32 --skip-cpu > gpurun_out/r01_bench_n1_f32_stream.json 2>> gpurun_out/bench_err.log; cut -c1-330 gpurun_out/r01_bench_n1_f32_stream.json
CMD2="python bench.py --steps 6 --warmup 3 --skip-cpu --skip-e2e"
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 30 --csv --log-file gpurun_out/r01_bench_n1_ncu_launches_stream.csv $CMD2 > gpurun_out/ncu2.log 2>&1; tail -1 gpurun_out/ncu2.log | cut -c1-200
cd geosongpu-ci_b200
for d in f64 f32; do
CMD="python -m b200stencil.bench.sweep --stencils fv_tp2d --iters 5 --dtypes $d"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fv_stream -s 3 -c 1 -f -o ../gpurun_out/fv_stream_r1_$d $CMD > ../gpurun_out/ncu_fvs_$d.log 2>&1; tail -1 ../gpurun_out/ncu_fvs_$d.log
done
timeout 300 python -m b200stencil.bench.sweep --stencils fv_tp2d --iters 20 --config C720x137 --out ../gpurun_out/r01_sweep_fv_c720.json 2>&1 | cut -c100-330
timeout 200 python -m b200stencil.bench.sweep --stencils fv_tp2d --iters 20 --graph --sub 192,192,3,72 2>&1 | cut -c100-330
