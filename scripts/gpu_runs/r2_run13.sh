# round 2, run 13 (1 GPU): virtual-rank halo tests three times (kernel preload at b2s_halo_init), exchange kernel version 1
# against version 2 (overlap probe, all-local links), step modes at N = 1, fresh ncu captures of the shipped
# fv_tp2d_split / remap_ppm kernels and of the exchange kernel
mkdir -p gpurun_out
for i in 1 2 3; do timeout 600 python -m pytest tests/test_gpu_halo_device.py -x -q -m gpu 2>&1 | tail -4; done | tee gpurun_out/r2_run13_halo_tests.log
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r2_run13_gpu_suite.log
cp gpurun_out/parity_pointwise.jsonl gpurun_out/r2_parity_pointwise.jsonl 2>/dev/null
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
export PYTHONPATH=geosongpu-ci_b200
for hv in 1 2; do
for n in 192 384; do
timeout 200 python scripts/overlap_probe.py --n $n --variants 2,3 --option halo_variant=$hv
done; done | tee gpurun_out/r2_run13_overlap_probe.jsonl | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print({k:v for k,v in d.items() if k.endswith('_us') or k in ('variant','cube','options')})"
for mode in serial overlap fused; do
timeout 300 python bench.py --steps 200 --warmup 10 --skip-cpu --skip-e2e --step $mode > gpurun_out/r2c_bench_n1_$mode.json 2> gpurun_out/r2c_bench_n1_$mode.err
python -c "
import json; d=json.loads(open('gpurun_out/r2c_bench_n1_$mode.json').read().strip().splitlines()[-1]); print('$mode', round(d['ms_per_step']*1e3,1), 'kernel', d['roofline']['kernel_ms'], 'halo', d['roofline']['halo_exchange_ms'], d['clocks']['sm_mhz'], d['clocks']['reasons'], d.get('halo_trace_ns'))"
done
cd geosongpu-ci_b200
CMD="python -m b200stencil.bench.sweep --stencils fv_tp2d_split --iters 5 --dtypes f64"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fv_split_stream -s 3 -c 1 -f -o ../gpurun_out/r02_fv_split_stream_f64 $CMD > ../gpurun_out/ncu_split2.log 2>&1; tail -1 ../gpurun_out/ncu_split2.log | cut -c1-300
CMD="python -m b200stencil.bench.sweep --stencils fv_tp2d_split --iters 5 --dtypes f32"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fv_split_stream -s 3 -c 1 -f -o ../gpurun_out/r02_fv_split_stream_f32 $CMD > ../gpurun_out/ncu_split2.log 2>&1; tail -1 ../gpurun_out/ncu_split2.log | cut -c1-300
for d in f64 f32; do
CMD="python -m b200stencil.bench.sweep --stencils remap_ppm --iters 4 --dtypes $d"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_remap_ppm -s 2 -c 1 -f -o ../gpurun_out/r02_remap_ppm_$d $CMD > ../gpurun_out/ncu_ppm2.log 2>&1; tail -1 ../gpurun_out/ncu_ppm2.log | cut -c1-300
done
cd ..
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_halo_exchange2 -s 4 -c 1 -f -o gpurun_out/r02_halo_exchange2 python scripts/overlap_probe.py --n 384 --variants 3 > gpurun_out/ncu_xchg2.log 2>&1; tail -1 gpurun_out/ncu_xchg2.log | cut -c1-200
