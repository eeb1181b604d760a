# round 2, run 19 (1 GPU): the relayed handshake wait (block 0 polls the peers, the other blocks a local word) under the
# virtual-rank tests, then what the driver runs at round end: the -m gpu suite, smoke(), bench.py with its defaults
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_halo_device.py -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/r2_run19_halo_tests.log
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/r2_run19_gpu_suite.log
cp gpurun_out/parity_pointwise.jsonl gpurun_out/r2_parity_pointwise.jsonl 2>/dev/null
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/r2_final_bench_n1.json 2> gpurun_out/r2_final_bench_n1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_final_bench_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','halo_check')}, d['config']['step_launch'], d['roofline'], d['clocks'], {k:d['e2e'][k] for k in ('value','ms_per_step','frac_of_pcie','matches_resident_path')}, d['cpu_baseline'])
PY
timeout 200 python bench.py --impl reference --steps 5 --warmup 1 | cut -c1-400
