mkdir -p gpurun_out
N=${1:-2}
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/multigpu_check.py 2>&1 | grep -v "^\*\*\*\|OMP_NUM" | tail -6
for mode in "--halo p2p" "--halo p2p --no-graph" "--halo nccl" "--halo nccl --no-overlap"; do
  tag=$(echo "$mode" | tr -d ' -')
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 500 --warmup 10 --skip-e2e $mode 2>gpurun_out/err_n${N}_$tag.log > gpurun_out/bench_n${N}_$tag.json
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_n${N}_$tag.json").read().strip().splitlines()[-1])
    print("N=$N mode='$mode'", "ms/step", round(d["ms_per_step"],4), "value", round(d["value"]/1e9,1), "Gpts/s kernel_ms", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], d["config"]["launch"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print("N=$N mode='$mode' FAILED", e); print(open("gpurun_out/err_n${N}_$tag.log").read()[-1500:])
PY
done
