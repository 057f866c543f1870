"""PCIe probe on the GPU box: pinned H2D / D2H times for the BASELINE buffer sizes, by the NUMA node
the pinned buffer was allocated from."""
import os, subprocess, time, torch
print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout)
print(subprocess.run(["bash", "-c", "lscpu | grep -i numa; cat /sys/bus/pci/devices/*/numa_node 2>/dev/null | sort | uniq -c"],
                     capture_output=True, text=True).stdout)
torch.cuda.init()
def t(fn, reps=10):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
all_cpus = sorted(os.sched_getaffinity(0))
print("affinity: %d cpus %s..%s" % (len(all_cpus), all_cpus[0], all_cpus[-1]))
halves = {"first-half": set(all_cpus[:len(all_cpus)//2]), "second-half": set(all_cpus[len(all_cpus)//2:])}
for name, cpus in halves.items():
    os.sched_setaffinity(0, cpus)
    for size in (81_000_000, 106_000_000, 81_000_000):
        h = torch.empty(size, dtype=torch.uint8).pin_memory(); h.fill_(1)
        d = torch.empty(size, dtype=torch.uint8, device="cuda")
        print("%-11s %4d MB  H2D %.2f ms  D2H %.2f ms" % (name, size // 1_000_000,
              t(lambda: d.copy_(h, non_blocking=True)), t(lambda: h.copy_(d, non_blocking=True))))
        del h, d
