# round 2, run 16 (2 GPUs): STAGED pushes (crossing strips packed by their owner into the destination's staging area,
# unpacked there after the delivery flag): virtual-rank tests, multigpu_check, serial step at N = 2 staged / in place / pull
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_halo_device.py -x -q -m gpu 2>&1 | tail -3
N=2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/multigpu_check.py > gpurun_out/r2_run16_multigpu_check_n$N.log 2>&1; grep -v "^\*\*\*\|OMP_NUM\|^W1\|^$" gpurun_out/r2_run16_multigpu_check_n$N.log | tail -8
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
for push in staged inplace off staged; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 500 --warmup 10 --skip-e2e --step serial --push $push 2>gpurun_out/r2f_bench_n${N}_serial_$push.err > gpurun_out/r2f_bench_n${N}_serial_$push.json
  show gpurun_out/r2f_bench_n${N}_serial_$push.json "N=$N serial push=$push"
done
