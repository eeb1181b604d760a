# round 2, run 14 (2 GPUs): gated tile kernel with the gate outside the item loop (tests + probe), then the exchange
# kernel version 1 against version 2 ACROSS GPUs: multigpu_check, and the three launch modes of the step at N = 2 with
# the device timeline of the last timed step
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_halo_device.py -x -q -m gpu 2>&1 | tail -3
PYTHONPATH=geosongpu-ci_b200 timeout 200 python scripts/overlap_probe.py --n 192 --variants 2,3 | tee gpurun_out/r2_run14_overlap_probe.jsonl | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print({k:v for k,v in d.items() if k.endswith('_us') or k in ('variant','cube','options')})"
N=2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/multigpu_check.py > gpurun_out/r2_run14_multigpu_check_n$N.log 2>&1; grep -v "^\*\*\*\|OMP_NUM\|^W1\|^$" gpurun_out/r2_run14_multigpu_check_n$N.log | tail -8
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
for hv in 2 1; do
for mode in serial overlap fused; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 500 --warmup 10 --skip-e2e --step $mode --option halo_variant=$hv 2>gpurun_out/r2d_bench_n${N}_${mode}_hv$hv.err > gpurun_out/r2d_bench_n${N}_${mode}_hv$hv.json
  show gpurun_out/r2d_bench_n${N}_${mode}_hv$hv.json "N=$N $mode hv=$hv"
done; done
