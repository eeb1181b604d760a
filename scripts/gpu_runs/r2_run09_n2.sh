# round 2, run 9 (2 GPUs): parallel handshake wait / more blocks -- traces at N = 2; tile kernel speed at the 8-GPU sub-domain shape; NUMA facts
mkdir -p gpurun_out
bash scripts/numa_probe.sh > gpurun_out/r2_numa_probe.txt 2>&1; tail -25 gpurun_out/r2_numa_probe.txt
timeout 300 PYTHONPATH=geosongpu-ci_b200 python -m b200stencil.bench.sweep --stencils fv_tp2d --dtypes f64 --sub 192,192,3,72 --graph --option fv_variant=2 2>&1 | tail -1
timeout 300 PYTHONPATH=geosongpu-ci_b200 python -m b200stencil.bench.sweep --stencils fv_tp2d --dtypes f64 --sub 192,192,3,72 --graph --option fv_variant=3 2>&1 | tail -1
timeout 300 PYTHONPATH=geosongpu-ci_b200 python -m b200stencil.bench.sweep --stencils fv_tp2d --dtypes f64 --sub 192,192,3,72 --option fv_variant=2 2>&1 | tail -1
timeout 300 python -m pytest tests/test_gpu_halo_device.py -x -q -m gpu 2>&1 | tail -2
N=2
show() { python - "$1" "$2" <<'PY'
import json,sys
f,label=sys.argv[1],sys.argv[2]
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(label, "us/step", round(d["ms_per_step"]*1e3,1), "kernel_us", round(d["roofline"]["kernel_ms"]*1e3,1), "halo_us", round(d["roofline"]["halo_exchange_ms"]*1e3,1), d["config"].get("step_launch"), "check", d["halo_check"], d["device_step_equals_nccl_step"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"], "trace", d.get("halo_trace_ns"))
except Exception as e:
    print(label, "FAILED", e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
}
for mode in "--step fused" "--step overlap" "--step serial"; do
  tag=$(echo "$mode" | tr -d ' -')
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 1000 --warmup 10 --skip-e2e $mode 2>gpurun_out/r2b_bench_n${N}_$tag.err > gpurun_out/r2b_bench_n${N}_$tag.json
  show gpurun_out/r2b_bench_n${N}_$tag.json "N=$N $mode"
done
