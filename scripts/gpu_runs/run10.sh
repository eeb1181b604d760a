mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "vertical or chain" 2>&1 | tail -3
cd geosongpu-ci_b200
show() { python -c "
import sys,json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); print(d['stencil'],d['config'],d['dtype'],d.get('options'),d['median_ms'],'ms',d['GBps'],'GB/s',d['frac_measured_peak'])
    else: print(line[:300])
"; }
echo "== remap variants"
for v in 0 2 3 1; do timeout 100 python -m b200stencil.bench.sweep --stencils remap --iters 5 --option remap_variant=$v 2>&1 | tail -2 | show; done
