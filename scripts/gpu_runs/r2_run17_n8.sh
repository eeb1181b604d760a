# round 2, run 17 (8 GPUs): multigpu_check on 8 GPUs with the mixed / staged exchange, then the launch modes of the step at
# N = 8 and N = 4 (serial with staged pushes, serial pull-only, overlapped), and N = 1, 2 on the same box for the curve
mkdir -p gpurun_out
N=8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/multigpu_check.py > gpurun_out/r2_run17_multigpu_check_n$N.log 2>&1; grep -v "^\*\*\*\|OMP_NUM\|^W1\|^$" gpurun_out/r2_run17_multigpu_check_n$N.log | tail -8
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
run() { # n, tag, args...
  n=$1; tag=$2; shift 2
  if [ "$n" = "1" ]; then
    timeout 200 python bench.py --steps 1000 --warmup 10 --skip-e2e --skip-cpu "$@" 2>gpurun_out/r2g_bench_n${n}_$tag.err > gpurun_out/r2g_bench_n${n}_$tag.json
  else
    timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --steps 1000 --warmup 10 --skip-e2e "$@" 2>gpurun_out/r2g_bench_n${n}_$tag.err > gpurun_out/r2g_bench_n${n}_$tag.json
  fi
  show gpurun_out/r2g_bench_n${n}_$tag.json "N=$n $tag"
}
run 8 serial_pull --step serial --push off
run 8 serial_staged --step serial --push staged
run 8 overlap --step overlap
run 4 serial_pull --step serial --push off
run 2 serial_pull --step serial --push off
run 1 serial --step serial
