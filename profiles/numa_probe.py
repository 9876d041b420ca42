import os, subprocess, torch
print(subprocess.run(["nvidia-smi","topo","-m"],capture_output=True,text=True).stdout[:1500])
print("cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for i in range(torch.cuda.device_count()):
    p = torch.cuda.get_device_properties(i)
    bdf = "%04x:%02x:%02x.0" % (getattr(p,'pci_domain_id',0), getattr(p,'pci_bus_id',0), getattr(p,'pci_device_id',0))
    for f in ("local_cpulist","numa_node"):
        path="/sys/bus/pci/devices/%s/%s"%(bdf,f)
        print(i, bdf, f, open(path).read().strip() if os.path.exists(path) else "n/a")
print(subprocess.run("lscpu | grep -i -E 'numa|socket|model name'", shell=True, capture_output=True, text=True).stdout)
