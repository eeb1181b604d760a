# round 2, run 7 (1 GPU): fused step (exchange + stencil in one launch) -- tests, probe, N = 1 bench in the three launch modes
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_halo_device.py tests/test_c_abi_driver.py tests/test_gpu_halo.py -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/r2_run07_tests.log
(
timeout 300 python scripts/overlap_probe.py --n 192 --variants 2,3
timeout 300 python scripts/overlap_probe.py --n 384 --variants 3
timeout 300 python scripts/overlap_probe.py --n 192 --variants 2 --dtype f32
) 2>&1 | tee gpurun_out/r2_overlap_probe_fused.jsonl
for mode in fused overlap serial; do
timeout 300 python bench.py --steps 200 --warmup 10 --step $mode --skip-cpu --skip-e2e > gpurun_out/r2_bench_n1_f64_$mode.json 2> gpurun_out/r2_bench_n1_f64_$mode.err; tail -c 600 gpurun_out/r2_bench_n1_f64_$mode.err
done
python - <<'PY'
import json
for f in ("r2_bench_n1_f64_fused","r2_bench_n1_f64_overlap","r2_bench_n1_f64_serial"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "ms/step", round(d["ms_per_step"],4), "Gpts/s", round(d["value"]/1e9,1), "kernel_ms", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], "halo_ms", d["roofline"]["halo_exchange_ms"], d["config"]["step_launch"], "check", d["halo_check"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"], d["config"]["region_ms"], d["gpu_launches"])
    except Exception as e:
        print(f, "FAILED", e)
PY
