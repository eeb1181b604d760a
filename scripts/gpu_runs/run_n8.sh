mkdir -p gpurun_out
N=${1:-8}
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/multigpu_check.py 2>&1 | grep -v "^\*\*\*\|OMP_NUM" | tail -25
for mode in "--halo p2p" "--halo nccl --no-overlap" "--halo nccl"; do
  tag=$(echo "$mode" | tr -d ' -')
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 1000 --warmup 10 $mode 2>gpurun_out/err_n${N}_$tag.log > gpurun_out/bench_n${N}_$tag.json
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_n${N}_$tag.json").read().strip().splitlines()[-1])
    print("N=$N mode='$mode'", "ms/step", round(d["ms_per_step"],4), "value", round(d["value"]/1e9,1), "Gpts/s kernel_ms", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], d["config"]["launch"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"], "e2e", d["e2e"] and round(d["e2e"]["value"]/1e9,2))
except Exception as e:
    print("N=$N mode='$mode' FAILED", e); print(open("gpurun_out/err_n${N}_$tag.log").read()[-1500:])
PY
done
