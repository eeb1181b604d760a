# round 2, run 5 (1 GPU): device timeline of the overlapped step
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_halo_device.py tests/test_c_abi_driver.py -x -q -m gpu 2>&1 | tail -3
(
timeout 300 python scripts/overlap_probe.py --n 192 --variants 2,3
timeout 300 python scripts/overlap_probe.py --n 192 --variants 2 --option halo_blocks_per_sm=2
timeout 300 python scripts/overlap_probe.py --n 192 --variants 2 --option halo_blocks_per_sm=1
timeout 300 python scripts/overlap_probe.py --n 384 --variants 3 --option halo_blocks_per_sm=2
) 2>&1 | tee gpurun_out/r2_overlap_probe_trace.jsonl
