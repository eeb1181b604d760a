mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
cd geosongpu-ci_b200
show() { python -c "
import sys,json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); print(d['stencil'],d['config'],d['dtype'],d.get('options'),d['median_ms'],'ms',d['GBps'],'GB/s',d['frac_measured_peak'])
    else: print(line[:300])
"; }
timeout 100 python -m b200stencil.bench.sweep --stencils remap --iters 5 2>&1 | tail -2 | show
cd ..
timeout 200 python bench.py --workload chain --steps 20 --warmup 3 | cut -c1-900
timeout 200 python bench.py --workload chain --steps 20 --warmup 3 --dtype f32 | cut -c1-400
