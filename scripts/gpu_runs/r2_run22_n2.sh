# round 2, run 22 (2 GPUs): the final defaults across real GPUs: multigpu_check and bench.py --gpus 2 (halo_check, NCCL bit-equality)
mkdir -p gpurun_out
N=2
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/multigpu_check.py > gpurun_out/r2_run22_multigpu_check_n$N.log 2>&1; grep -v "^\*\*\*\|OMP_NUM\|^W1\|^$" gpurun_out/r2_run22_multigpu_check_n$N.log | tail -6
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 500 --warmup 10 2>gpurun_out/r2_final_bench_n2.err > gpurun_out/r2_final_bench_n2.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_final_bench_n2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','halo_check','device_step_equals_nccl_step')}, d['config']['step_launch'], d['roofline']['kernel_ms'], d['clocks']['sm_mhz'], d['e2e'] and {k:d['e2e'][k] for k in ('value','ms_per_step','matches_resident_path')})
PY
tail -2 gpurun_out/r2_final_bench_n2.err | cut -c1-300
