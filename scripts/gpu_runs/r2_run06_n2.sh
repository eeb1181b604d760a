# round 2, run 6 (2 GPUs): the library-owned exchange across processes (cudaIpc), multigpu_check, N = 2 bench lines
mkdir -p gpurun_out
N=${1:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/multigpu_check.py > gpurun_out/r2_multigpu_check_n$N.log 2>&1; grep -v "^\*\*\*\|OMP_NUM\|^W1\|^$" gpurun_out/r2_multigpu_check_n$N.log | tail -25
for mode in "" "--no-overlap" "--halo nccl --no-overlap" "--halo nccl"; do
  tag=$(echo "x$mode" | tr -d ' -')
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 400 --warmup 10 $mode 2>gpurun_out/r2_err_n${N}_$tag.log > gpurun_out/r2_bench_n${N}_$tag.json
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_bench_n${N}_$tag.json").read().strip().splitlines()[-1])
    print("N=$N mode='$mode'", "us/step", round(d["ms_per_step"]*1e3,1), "Gpts/s", round(d["value"]/1e9,1), "kernel_us", round(d["roofline"]["kernel_ms"]*1e3,1), "frac", d["roofline"]["frac"], "halo_us", round(d["roofline"]["halo_exchange_ms"]*1e3,1), "overlap", d["config"]["overlap_exchange"], "check", d["halo_check"], d["device_step_equals_nccl_step"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"], "e2e", d["e2e"] and (round(d["e2e"]["value"]/1e9,2), d["e2e"]["frac_of_pcie"], d["e2e"]["numa_node"], d["e2e"]["matches_resident_path"]), [round(x,1) for x in d["config"]["region_ms"]])
except Exception as e:
    print("N=$N mode='$mode' FAILED", e); print(open("gpurun_out/r2_err_n${N}_$tag.log").read()[-1500:])
PY
done
