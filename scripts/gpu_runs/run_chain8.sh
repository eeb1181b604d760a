mkdir -p gpurun_out
for mode in "" "--fused-remap"; do
  tag=$(echo "$mode" | tr -d ' -')
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --workload chain --steps 100 --warmup 5 $mode > gpurun_out/r01_chain_n8_f64_$tag.json 2>gpurun_out/chain8_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r01_chain_n8_f64_$tag.json").read().strip().splitlines()[-1])
    print("chain N=8 '$mode'", round(d["ms_per_step"],4), "ms/step", round(d["value"]/1e9,1), "Gpts/s", d["roofline"]["achieved"], "GB/s per GPU", d["roofline"]["frac"], d["config"])
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/chain8_$tag.err").read()[-1500:])
PY
done
