# round 2, run 4 (1 GPU): persistent exchange kernel -- tests, overlap probe, N = 1 bench lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_halo_device.py tests/test_c_abi_driver.py tests/test_gpu_parity.py tests/test_gpu_halo.py -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/r2_run04_tests.log
(
timeout 300 python scripts/overlap_probe.py --n 192 --variants 2,3
timeout 300 python scripts/overlap_probe.py --n 192 --variants 2 --option halo_blocks_per_sm=2
timeout 300 python scripts/overlap_probe.py --n 192 --variants 2 --option halo_blocks_per_sm=8
timeout 300 python scripts/overlap_probe.py --n 384 --variants 3
timeout 300 python scripts/overlap_probe.py --n 192 --variants 2,3 --dtype f32
) 2>&1 | tee gpurun_out/r2_overlap_probe_persistent.jsonl
timeout 300 python bench.py --steps 200 --warmup 10 --skip-cpu --skip-e2e > gpurun_out/r2_bench_n1_f64.json 2> gpurun_out/r2_bench_n1_f64.err; tail -c 600 gpurun_out/r2_bench_n1_f64.err
timeout 300 python bench.py --steps 200 --warmup 10 --overlap --skip-cpu --skip-e2e > gpurun_out/r2_bench_n1_f64_overlap.json 2> gpurun_out/r2_bench_n1_f64_overlap.err; tail -c 600 gpurun_out/r2_bench_n1_f64_overlap.err
python - <<'PY'
import json
for f in ("r2_bench_n1_f64","r2_bench_n1_f64_overlap"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "ms/step", round(d["ms_per_step"],4), "Gpts/s", round(d["value"]/1e9,1), "kernel_ms", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], "halo_ms", d["roofline"]["halo_exchange_ms"], "overlap", d["config"]["overlap_exchange"], "check", d["halo_check"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"], d["config"]["region_ms"])
    except Exception as e:
        print(f, "FAILED", e)
PY
