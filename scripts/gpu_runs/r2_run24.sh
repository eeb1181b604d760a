# round 2, run 24 (1 GPU, last GPU-minutes of the round): the reference's unmodified pattern files on the device, then smoke()
timeout 100 python -m pytest tests/test_gpu_ref_patterns.py -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r2_run24_ref_patterns.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r2_run24_smoke.log
