# round 2, run 18 (8 GPUs): A/B of the step's launch modes and exchange-kernel options inside one job (scripts/step_modes_probe.py)
mkdir -p gpurun_out
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 scripts/step_modes_probe.py --steps 1000 --regions 3 2> gpurun_out/r2_run18_probe_n8.err | tee gpurun_out/r2_run18_probe_n8.jsonl | cut -c1-260
tail -3 gpurun_out/r2_run18_probe_n8.err | cut -c1-300
