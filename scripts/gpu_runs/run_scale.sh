# strong-scaling record on one box: N = 1, 2, 4, 8 back to back (what the driver does), p2p halo exchange
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 scripts/multigpu_check.py 2>&1 | grep "multigpu_check"
timeout 200 python bench.py --steps 1000 --warmup 10 --skip-cpu > gpurun_out/r01_scale_n1.json 2>gpurun_out/scale_n1.err
for N in 2 4 8; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 1000 --warmup 10 > gpurun_out/r01_scale_n$N.json 2>gpurun_out/scale_n$N.err
done
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 1000 --warmup 10 --halo nccl --no-overlap --skip-e2e > gpurun_out/r01_scale_n8_nccl.json 2>gpurun_out/scale_n8_nccl.err
python - <<'PY'
import json
t1=None
for n in (1,2,4,8,"8_nccl"):
    try:
        d=json.loads(open(f"gpurun_out/r01_scale_n{n}.json").read().strip().splitlines()[-1])
        if n==1: t1=d["ms_per_step"]
        N=d["n_gpus"]
        print(f"N={n}: {d['ms_per_step']*1e3:.1f} us/step  {d['value']/1e9:.1f} Gpts/s  eff={t1/(N*d['ms_per_step'])*100:.1f}%  kernel {d['roofline']['kernel_ms']*1e3:.1f} us frac {d['roofline']['frac']}  e2e {d['e2e'] and round(d['e2e']['value']/1e9,2)}  clocks {d['clocks']['sm_mhz']} {d['clocks']['reasons']}  {d['config']['halo_exchange'][:30]}")
    except Exception as e:
        print("N=",n,"FAILED",e); print(open(f"gpurun_out/scale_n{n}.err").read()[-800:])
PY
