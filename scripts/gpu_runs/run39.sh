timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
cd geosongpu-ci_b200
run() { timeout 100 python -m b200stencil.bench.sweep --stencils fv_tp2d --iters 30 --dtypes $3 --graph --sub $1 $2 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$1 $3', d.get('options'), d['median_ms'], d['min_ms'], d['frac_measured_peak'])"; }
run 192,192,3,72 "--option fv_variant=2" f64
run 192,192,3,72 "--option fv_variant=3" f64
run 192,192,3,72 "--option fv_variant=3 --option fv_jb=96" f64
run 192,192,3,72 "--option fv_variant=2" f32
run 192,192,3,72 "--option fv_variant=3" f32
run 192,192,6,72 "--option fv_variant=2" f64
run 192,192,6,72 "--option fv_variant=3" f64
