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
echo "== remap depth"
for d in 2 4 8; do timeout 100 python -m b200stencil.bench.sweep --stencils remap --iters 5 --option remap_depth=$d 2>&1 | tail -2 | show; done
echo "== saturation variants"
for u in 1 2 4; do for kc in 4 8 24 72; do timeout 100 python -m b200stencil.bench.sweep --stencils saturation_adjust --graph --iters 5 --option sat_unroll=$u --option sat_kchunk=$kc 2>&1 | tail -2 | show; done; done
echo "== while / patterns at C96 (graph) and C384"
timeout 100 python -m b200stencil.bench.sweep --stencils top_of_column,while_in_function,hybrid_index_2dout --graph --iters 10 2>&1 | tail -6 | show
timeout 100 python -m b200stencil.bench.sweep --stencils top_of_column,while_in_function,hybrid_index_2dout,find_klcl,cloud_top,saturation_adjust --config C384x72 --iters 10 2>&1 | tail -12 | show
