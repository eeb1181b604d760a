mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "remap or vertical" 2>&1 | tail -5
timeout 300 python scripts/remap_variants.py --variants 3:8:1,2 --out gpurun_out/r01_remap_trimmed.json 2>&1 | tail -30
