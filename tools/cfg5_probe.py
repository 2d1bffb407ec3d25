"""Dev probe (torchrun): the cfg5 training step with several gradient-bucket sizes."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, '3d-vlm-gd_b200'))
import torch
import bench, bench_extras
rank, world, local = bench.dist_setup(0)
dev = torch.device('cuda', local)
class A: steps = 20
for mb in [int(x) for x in (sys.argv[1:] or ['25', '100', '400'])]:
    blk = bench_extras.cfg5_block(A, rank, world, local, dev, bucket_mb=mb)
    if rank == 0:
        print(mb, 'MB buckets:', json.dumps({k: blk[k] for k in ('step_ms', 'step_ms_without_allreduce', 'exposed_allreduce_ms', 'buckets')}), flush=True)
if world > 1:
    torch.distributed.destroy_process_group()
