# round 2, run 8 (8 GPUs): multigpu_check, the three launch modes of the device-exchange step + the NCCL baseline at N = 8,
# then the strong-scaling record N = 1, 2, 4, 8 in the default mode (what the driver does), and the cfg5 chain
mkdir -p gpurun_out
N=8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/multigpu_check.py > gpurun_out/r2_multigpu_check_n$N.log 2>&1; grep -v "^\*\*\*\|OMP_NUM\|^W1\|^$" gpurun_out/r2_multigpu_check_n$N.log | tail -8
show() { python - "$1" "$2" <<'PY'
import json,sys
f,label=sys.argv[1],sys.argv[2]
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(label, "us/step", round(d["ms_per_step"]*1e3,1), "Gpts/s", round(d["value"]/1e9,1), "kernel_us", round(d["roofline"]["kernel_ms"]*1e3,1), "frac", d["roofline"]["frac"], "halo_us", round(d["roofline"]["halo_exchange_ms"]*1e3,1), d["config"].get("step_launch"), "check", d["halo_check"], d["device_step_equals_nccl_step"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"], "e2e", d["e2e"] and (round(d["e2e"]["value"]/1e9,2), d["e2e"]["frac_of_pcie"], d["e2e"]["numa_node"], d["e2e"]["pcie_gbs"]["h2d"], d["e2e"]["matches_resident_path"]), [round(x,2) for x in d["config"]["region_ms"]])
except Exception as e:
    print(label, "FAILED", e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
}
for mode in "--step fused" "--step overlap" "--step serial" "--halo nccl --no-overlap"; do
  tag=$(echo "$mode" | tr -d ' -')
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 1000 --warmup 10 --skip-e2e $mode 2>gpurun_out/r2_bench_n${N}_$tag.err > gpurun_out/r2_bench_n${N}_$tag.json
  show gpurun_out/r2_bench_n${N}_$tag.json "N=$N $mode"
done
# the scaling record, default mode, e2e included
timeout 300 python bench.py --steps 1000 --warmup 10 --skip-cpu > gpurun_out/r2_scale_n1.json 2>gpurun_out/r2_scale_n1.err; show gpurun_out/r2_scale_n1.json "scale N=1"
for n in 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 1000 --warmup 10 > gpurun_out/r2_scale_n$n.json 2>gpurun_out/r2_scale_n$n.err; show gpurun_out/r2_scale_n$n.json "scale N=$n"
done
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 4 --steps 1000 --warmup 10 --skip-e2e --step serial > gpurun_out/r2_bench_n4_stepserial.json 2>gpurun_out/r2_bench_n4_stepserial.err; show gpurun_out/r2_bench_n4_stepserial.json "N=4 serial"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 1000 --warmup 10 --skip-e2e --step serial > gpurun_out/r2_bench_n2_stepserial.json 2>gpurun_out/r2_bench_n2_stepserial.err; show gpurun_out/r2_bench_n2_stepserial.json "N=2 serial"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 8 --workload chain --fused-remap --steps 200 --warmup 5 > gpurun_out/r2_chain_n8.json 2>gpurun_out/r2_chain_n8.err; tail -c 700 gpurun_out/r2_chain_n8.json
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29536 bench.py --gpus 8 --impl reference --steps 5 --warmup 1 > gpurun_out/r2_reference_n8.json 2>gpurun_out/r2_reference_n8.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_reference_n8.json').read().strip().splitlines()[-1]); print('reference under torchrun N=8:', round(d['value']/1e9,2), 'Gpts/s', d['cpu_baseline']['cores'], 'cores', round(d['ms_per_step'],1), 'ms/step')"
python - <<'PY'
import json
t1=None
for n in (1,2,4,8):
    try:
        d=json.loads(open(f"gpurun_out/r2_scale_n{n}.json").read().strip().splitlines()[-1])
        if n==1: t1=d["ms_per_step"]
        print(f"N={n}: {d['ms_per_step']*1e3:.1f} us/step eff={t1/(n*d['ms_per_step'])*100:.1f}%")
    except Exception as e: print(n, "FAILED", e)
PY
