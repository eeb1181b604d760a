# round 2, run 27 (1 GPU, what is left of the budget): the reordered collection (device-wait tests last) on the device
timeout 20 python -m pytest tests/test_c_abi_driver.py tests/test_gpu_golden.py tests/test_gpu_ref_patterns.py -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/r2_run27_reorder.log
