cd geosongpu-ci_b200
run() { timeout 100 python -m b200stencil.bench.sweep --stencils fv_tp2d --iters 20 --dtypes f64 --graph --sub $1 $2 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$1', d.get('options'), d['median_ms'], d['frac_measured_peak'])"; }
for sub in 192,192,3,72 384,192,3,72 384,384,3,72; do
run $sub "--option fv_variant=2"
run $sub "--option fv_variant=3"
run $sub "--option fv_variant=3 --option fv_jb=192"
run $sub "--option fv_variant=3 --option fv_jb=64"
run $sub "--option fv_variant=3 --option fv_stages=4"
done
run 192,192,3,72 "--option fv_variant=3 --option fv_jb=48"
run 192,192,3,72 "--option fv_variant=3 --option fv_ti=128"
run 192,192,3,72 "--option fv_variant=3 --option fv_jb=192 --option fv_stages=2"
