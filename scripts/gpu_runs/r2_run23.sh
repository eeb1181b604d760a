# round 2, run 23 (1 GPU): the halo lifecycle after the last host-side changes (segment stamp, plan upload clean-up)
timeout 400 python -m pytest tests/test_gpu_halo_device.py tests/test_c_abi_driver.py -x -q -m gpu 2>&1 | tail -3
