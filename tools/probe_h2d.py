"""Pinned H2D bandwidth with the process bound to each NUMA node (dev probe for the e2e leg of bench.py)."""
import glob
import os
import sys
import time

import torch


def cpulist(path):
    cpus = []
    for part in open(path).read().strip().split(','):
        if '-' in part:
            a, b = part.split('-')
            cpus += list(range(int(a), int(b) + 1))
        elif part:
            cpus.append(int(part))
    return cpus


def gpu_numa_node(dev=0):
    p = torch.cuda.get_device_properties(dev)
    bus = f'{getattr(p, "pci_domain_id", 0):04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0'
    path = f'/sys/bus/pci/devices/{bus}/numa_node'
    return bus, (int(open(path).read()) if os.path.exists(path) else None)


def bw(nbytes=512 << 20, reps=5):
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    host.fill_(1)
    dev = torch.empty(nbytes, dtype=torch.uint8, device='cuda')
    dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    return nbytes * reps / (time.perf_counter() - t0) / 1e9


def main():
    torch.cuda.init()
    print('gpu0', gpu_numa_node(0), 'affinity', len(os.sched_getaffinity(0)), 'cpus')
    nodes = sorted(glob.glob('/sys/devices/system/node/node[0-9]*'))
    print('default', round(bw(), 1), 'GB/s')
    allowed = os.sched_getaffinity(0)
    for n in nodes:
        cpus = set(cpulist(os.path.join(n, 'cpulist'))) & allowed
        if not cpus:
            print(os.path.basename(n), 'no allowed cpus')
            continue
        os.sched_setaffinity(0, cpus)
        print(os.path.basename(n), len(cpus), 'cpus', round(bw(), 1), 'GB/s')
    os.sched_setaffinity(0, allowed)


if __name__ == '__main__':
    main()
