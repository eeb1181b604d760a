# round 2, run 20 (8 GPUs): the relayed handshake wait against the per-block one, and exchange version 4 (one-block handshake
# kernel + flat-grid pull, two launches), inside one job
mkdir -p gpurun_out
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 scripts/step_modes_probe.py --steps 1000 --regions 3 --configs stencil,exchange,exchange:halo_variant=4,serial,serial:halo_handshake=1,serial:halo_variant=4,serial:halo_variant=1,serial:halo_variant=1:halo_handshake=1,overlap,overlap:halo_handshake=1,serial:halo_variant=4:halo_levels=4 2> gpurun_out/r2_run20_probe_n8.err | tee gpurun_out/r2_run20_probe_n8.jsonl | cut -c1-230
tail -3 gpurun_out/r2_run20_probe_n8.err | cut -c1-300
