"""Probe: does torch symmetric memory (peer-mapped buffers + device barrier) work on this box?"""
import os, sys, time
import torch, torch.distributed as dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
import torch.distributed._symmetric_memory as symm_mem
t = symm_mem.empty(1 << 20, dtype=torch.float64, device=dev)
t.fill_(float(rank))
hdl = symm_mem.rendezvous(t, group=dist.group.WORLD)
print(rank, "buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs][:4], "signal_pads", len(hdl.signal_pad_ptrs), flush=True)
hdl.barrier(channel=0)
peer = (rank + 1) % world
pt = hdl.get_buffer(peer, (1 << 20,), torch.float64)
torch.cuda.synchronize()
print(rank, "peer value", float(pt[5].item()), flush=True)
# timing of barrier
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200): hdl.barrier(channel=0)
e1.record(); torch.cuda.synchronize()
print(rank, "symm barrier us", e0.elapsed_time(e1) / 200 * 1e3, flush=True)
# peer copy bandwidth via torch copy from peer buffer
mine = torch.empty(1 << 20, dtype=torch.float64, device=dev)
e0.record()
for _ in range(50): mine.copy_(pt)
e1.record(); torch.cuda.synchronize()
print(rank, "peer pull GB/s", 8 * (1 << 20) * 50 / (e0.elapsed_time(e1) * 1e-3) / 1e9, flush=True)
sys.stdout.flush()
os._exit(0)
