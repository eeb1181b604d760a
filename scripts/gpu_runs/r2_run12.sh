mkdir -p gpurun_out
for i in 1 2 3; do timeout 900 python -m pytest tests/test_gpu_halo_device.py -x -q -m gpu 2>&1 | tail -3; done | tee gpurun_out/r2_run12_tests.log
timeout 600 python -m pytest tests/test_c_abi_driver.py -x -q -m gpu 2>&1 | tail -3
