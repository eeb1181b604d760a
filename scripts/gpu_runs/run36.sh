mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 400 --warmup 10 > gpurun_out/r01_scale_n2_stream.json 2> gpurun_out/scale_n2_stream.err
tail -c 400 gpurun_out/scale_n2_stream.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r01_scale_n2_stream.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus','roofline','clocks')}); print(d['config']['halo_exchange'], d['e2e'])
PY
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29572 scripts/multigpu_check.py 2>&1 | grep -v "^\*\*\*\|OMP_NUM" | tail -3
