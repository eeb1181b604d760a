# round 2, run 15 (2 GPUs): the mixed exchange (same-GPU strips pulled, strips that cross NVLink PUSHED by their owner,
# k_halo_exchange3): virtual-rank tests, multigpu_check, and the serial step at N = 2 with pushes against pull only,
# beside the overlapped step
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_halo_device.py tests/test_c_abi_driver.py -x -q -m gpu 2>&1 | tail -3
N=2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/multigpu_check.py > gpurun_out/r2_run15_multigpu_check_n$N.log 2>&1; grep -v "^\*\*\*\|OMP_NUM\|^W1\|^$" gpurun_out/r2_run15_multigpu_check_n$N.log | tail -8
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
for opt in "halo_push=1" "halo_push=0"; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 500 --warmup 10 --skip-e2e --step serial --option $opt 2>gpurun_out/r2e_bench_n${N}_serial_$opt.err > gpurun_out/r2e_bench_n${N}_serial_$opt.json
  show gpurun_out/r2e_bench_n${N}_serial_$opt.json "N=$N serial $opt"
done
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 500 --warmup 10 --skip-e2e --step overlap 2>gpurun_out/r2e_bench_n${N}_overlap.err > gpurun_out/r2e_bench_n${N}_overlap.json
show gpurun_out/r2e_bench_n${N}_overlap.json "N=$N overlap"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 500 --warmup 10 --skip-e2e --step serial --option halo_blocks_per_sm=2 2>gpurun_out/r2e_bench_n${N}_serial_bps2.err > gpurun_out/r2e_bench_n${N}_serial_bps2.json
show gpurun_out/r2e_bench_n${N}_serial_bps2.json "N=$N serial push 2 blocks/SM"
