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
echo "== C720 fv sweep"
for ti in 128 192; do for rs in "4 2" "8 2" "8 3"; do set -- $rs
timeout 100 python -m b200stencil.bench.sweep --stencils fv_tp2d --config C720x137 --iters 4 --option fv_ti=$ti --option fv_rows=$1 --option fv_stages=$2 2>&1 | tail -2 | show
done; done
echo "== 360x360x3x137 (cfg5 at N=8)"
timeout 100 python -m b200stencil.bench.sweep --stencils fv_tp2d --sub 360,360,3,137 --iters 8 2>&1 | tail -2 | show
for ti in 96 128 192; do timeout 100 python -m b200stencil.bench.sweep --stencils fv_tp2d --sub 360,360,3,137 --dtypes f64 --iters 8 --option fv_ti=$ti 2>&1 | tail -1 | show; done
