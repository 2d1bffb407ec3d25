"""H2D probe for the e2e leg: topology, NVML affinity, bandwidth per NUMA node, per tensor size, one vs many tensors."""
import os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench_extras
print(subprocess.run(['nvidia-smi', 'topo', '-m'], capture_output=True, text=True).stdout[:3000])
print('cpus allowed', len(os.sched_getaffinity(0)), 'cpu_count', os.cpu_count())
print(subprocess.run('lscpu | grep -i "numa\|model name\|socket"', shell=True, capture_output=True, text=True).stdout)
torch.cuda.init()
def bw(nbytes, reps=8, parts=1):
    hs = [torch.empty(nbytes // parts, dtype=torch.uint8).pin_memory() for _ in range(parts)]
    for h in hs: h.fill_(1)
    ds = [torch.empty(nbytes // parts, dtype=torch.uint8, device='cuda') for _ in range(parts)]
    for d, h in zip(ds, hs): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        for d, h in zip(ds, hs): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    return nbytes * reps / (time.perf_counter() - t0) / 1e9
print('default affinity: 512MB x1 %.1f GB/s, 8 parts %.1f, 64MB %.1f' % (bw(512 << 20), bw(512 << 20, parts=8), bw(64 << 20)))
print(bench_extras.bind_to_gpu_numa_node(0))
print('bound: 512MB x1 %.1f GB/s, 8 parts %.1f, 64MB %.1f' % (bw(512 << 20), bw(512 << 20, parts=8), bw(64 << 20)))
import glob
for n in sorted(glob.glob('/sys/devices/system/node/node[0-9]*')):
    txt = open(n + '/cpulist').read().strip()
    cpus = set()
    for part in txt.split(','):
        if '-' in part:
            a, b = part.split('-'); cpus |= set(range(int(a), int(b) + 1))
        elif part: cpus.add(int(part))
    try:
        os.sched_setaffinity(0, cpus)
        print(os.path.basename(n), len(cpus), 'cpus: %.1f GB/s' % bw(512 << 20))
    except Exception as e:
        print(os.path.basename(n), 'cannot bind', e)
