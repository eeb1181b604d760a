mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "fv_tp2d or region" 2>&1 | tail -3
cd geosongpu-ci_b200
show() { python -c "
import sys,json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); print(d['config'],d['dtype'],d.get('options'),d['median_ms'],'ms',d['GBps'],'GB/s',d['frac_measured_peak'])
    else: print(line[:300])
"; }
echo "== 192x192x3x72 (N=8 shape), graph timing"
for ti in 64 96 128 192; do for rs in "4 2" "8 2" "4 3"; do set -- $rs
  timeout 100 python -m b200stencil.bench.sweep --stencils fv_tp2d --sub 192,192,3,72 --graph --iters 10 --option fv_ti=$ti --option fv_rows=$1 --option fv_stages=$2 2>&1 | tail -2 | show
done; done
echo "== 384x96x3x72 (alt N=8 layout 1x4)"
timeout 100 python -m b200stencil.bench.sweep --stencils fv_tp2d --sub 384,96,3,72 --graph --iters 10 2>&1 | tail -2 | show
echo "== C384 full"
for ti in 96 128 192; do
  timeout 100 python -m b200stencil.bench.sweep --stencils fv_tp2d --iters 10 --option fv_ti=$ti 2>&1 | tail -2 | show
done
echo "== C720x137 ti sweep (f64)"
for ti in 96 128 192; do
  timeout 100 python -m b200stencil.bench.sweep --stencils fv_tp2d --config C720x137 --dtypes f64 --iters 5 --option fv_ti=$ti 2>&1 | tail -1 | show
done
