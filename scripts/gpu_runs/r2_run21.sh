# round 2, run 21 (1 GPU): what the driver runs at round end, with the final defaults (exchange = handshake kernel + flat pull)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r2_run21_gpu_suite.log
cp gpurun_out/parity_pointwise.jsonl gpurun_out/r2_parity_pointwise.jsonl 2>/dev/null
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py --skip-cpu > gpurun_out/r2_final_bench_n1.json 2> gpurun_out/r2_final_bench_n1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_final_bench_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','halo_check')}, d['config']['step_launch'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['clocks'], {k:d['e2e'][k] for k in ('value','ms_per_step','frac_of_pcie','matches_resident_path')})
PY
tail -3 gpurun_out/r2_final_bench_n1.err | cut -c1-300
