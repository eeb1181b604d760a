cd geosongpu-ci_b200
run() { timeout 100 python -m b200stencil.bench.sweep --stencils fv_tp2d --iters 20 --dtypes $3 --config $1 $2 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$1 $3', d.get('options'), d['median_ms'], d['min_ms'], d['frac_measured_peak'])"; }
run C720x137 "" f32
run C720x137 "--option fv_jb=720" f32
run C720x137 "--option fv_jb=360" f32
run C720x137 "" f64
run C720x137 "--option fv_jb=720" f64
run C384x72 "" f32
run C384x72 "--option fv_jb=128" f32
run C384x72 "--option fv_jb=384" f32
