# what the box says about GPU <-> NUMA placement (for hostio.bind_to_gpu_numa_node)
nvidia-smi topo -m 2>&1 | head -20
nvidia-smi --query-gpu=index,pci.bus_id --format=csv,noheader
for d in /sys/bus/pci/devices/*; do v=$(cat $d/vendor 2>/dev/null); c=$(cat $d/class 2>/dev/null); if [ "$v" = "0x10de" ] && [ "${c:0:6}" = "0x0302" ]; then echo "$d numa_node=$(cat $d/numa_node) local_cpulist=$(cat $d/local_cpulist)"; fi; done
ls /sys/devices/system/node/ | head; for n in /sys/devices/system/node/node*; do echo "$n cpulist=$(cat $n/cpulist)"; done
python - <<'PY'
import torch,os
p=torch.cuda.get_device_properties(0)
print([a for a in dir(p) if 'pci' in a], getattr(p,'pci_bus_id',None), getattr(p,'pci_device_id',None), getattr(p,'pci_domain_id',None))
print("affinity", len(os.sched_getaffinity(0)), "cpu_count", os.cpu_count())
PY
