# round 2, run 1 (1 GPU): new halo lifecycle + gated stencil tests, then the whole gpu suite, then bench lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_halo_device.py tests/test_c_abi_driver.py -x -q -m gpu 2>&1 | tail -25 | tee gpurun_out/r2_run01_newtests.log
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r2_run01_gpu_suite.log
timeout 300 python bench.py --steps 200 --warmup 10 > gpurun_out/r2_bench_n1_f64.json 2> gpurun_out/r2_bench_n1_f64.err; tail -c 600 gpurun_out/r2_bench_n1_f64.err
timeout 300 python bench.py --steps 200 --warmup 10 --overlap --skip-cpu --skip-e2e > gpurun_out/r2_bench_n1_f64_overlap.json 2> gpurun_out/r2_bench_n1_f64_overlap.err; tail -c 600 gpurun_out/r2_bench_n1_f64_overlap.err
timeout 300 python bench.py --steps 200 --warmup 10 --no-graph --skip-cpu --skip-e2e > gpurun_out/r2_bench_n1_f64_eager.json 2> gpurun_out/r2_bench_n1_f64_eager.err
timeout 300 python bench.py --workload patterns --steps 20 > gpurun_out/r2_patterns.json 2> gpurun_out/r2_patterns.err; tail -c 400 gpurun_out/r2_patterns.err
timeout 200 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
python - <<'PY'
import json
for f in ("r2_bench_n1_f64","r2_bench_n1_f64_overlap","r2_bench_n1_f64_eager"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "ms/step", round(d["ms_per_step"],4), "Gpts/s", round(d["value"]/1e9,1), "kernel_ms", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], "halo_ms", d["roofline"]["halo_exchange_ms"], d["config"]["launch"][:12], "overlap", d["config"]["overlap_exchange"], "check", d["halo_check"], "e2e", d["e2e"] and (round(d["e2e"]["value"]/1e9,2), d["e2e"]["frac_of_pcie"], d["e2e"]["pcie_gbs"], d["e2e"]["matches_resident_path"]), "cpu", d["cpu_baseline"] and (round(d["cpu_baseline"]["value"]/1e9,2), d["cpu_baseline"]["cores"]), d["clocks"]["sm_mhz"], d["clocks"]["reasons"], d["config"]["region_ms"])
    except Exception as e:
        print(f, "FAILED", e)
try:
    d=json.loads(open("gpurun_out/r2_patterns.json").read().strip().splitlines()[-1])
    for k,v in d["stencils"].items(): print(k, v)
except Exception as e: print("patterns FAILED", e)
try:
    d=json.loads(open("gpurun_out/r2_bench_reference.json").read().strip().splitlines()[-1]); print("reference", round(d["value"]/1e9,2), d["cpu_baseline"]["cores"], d["ms_per_step"])
except Exception as e: print("reference FAILED", e)
PY
