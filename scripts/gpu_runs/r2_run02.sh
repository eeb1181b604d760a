# round 2, run 2 (1 GPU): where the overlapped step's time goes, per kernel variant and sub-domain size
mkdir -p gpurun_out
(
timeout 300 python scripts/overlap_probe.py --n 192 --variants 2,3
timeout 300 python scripts/overlap_probe.py --n 192 --variants 2 --option fv_ti=32
timeout 300 python scripts/overlap_probe.py --n 192 --variants 3 --option fv_ti=64
timeout 300 python scripts/overlap_probe.py --n 384 --variants 3,2
timeout 300 python scripts/overlap_probe.py --n 384 --variants 3 --option fv_ti=64
timeout 300 python scripts/overlap_probe.py --n 192 --variants 2,3 --dtype f32
) 2>&1 | tee gpurun_out/r2_overlap_probe.jsonl
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
