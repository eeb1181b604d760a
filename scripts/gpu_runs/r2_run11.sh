# round 2, run 11 (1 GPU): A/B of the fv kernels, round-1 library (_old_r1 worktree of 30e86f7) against the current one, same box
for tree in _old_r1/geosongpu-ci_b200 geosongpu-ci_b200 _old_r1/geosongpu-ci_b200 geosongpu-ci_b200; do
export PYTHONPATH=$tree
echo "== $tree"
for v in 2 3; do
timeout 300 python -m b200stencil.bench.sweep --stencils fv_tp2d --dtypes f64 --sub 192,192,3,72 --graph --option fv_variant=$v 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['config'], d['options'], d['median_ms'], d['min_ms'])"
done
timeout 300 python -m b200stencil.bench.sweep --stencils fv_tp2d --dtypes f64 --config C384x72 --graph 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['config'], d['median_ms'], d['min_ms'])"
done
export PYTHONPATH=geosongpu-ci_b200
timeout 300 python scripts/overlap_probe.py --n 192 --variants 2,3 | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print({k:v for k,v in d.items() if k.endswith('_us') or k=='variant'})"
timeout 600 python -m pytest tests/test_gpu_halo_device.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
