"""Host-side ceiling of the end-to-end path: pinned-memory H2D / D2H bandwidth per GPU, alone and with all ranks
copying at once, plus the NUMA facts of the box (how many nodes, which node each GPU hangs off, which CPUs this
process may use).  Run under torchrun for N > 1:

    python tools/host_copy_probe.py                       # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        tools/host_copy_probe.py

Each rank optionally binds itself to the CPUs of its GPU's NUMA node BEFORE pinning (OFB_NUMA_BIND=1), which is what
bench.py does for its end-to-end arm.  Rank 0 prints one JSON object."""
import glob
import json
import os
import subprocess
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oflibpytorch_b200 import numa     # noqa: E402

rank = int(os.environ.get('RANK', '0'))
world = int(os.environ.get('WORLD_SIZE', '1'))
local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group('nccl', device_id=dev)

bound = None
if os.environ.get('OFB_NUMA_BIND', '1') == '1':
    bound = numa.bind_to_gpu_node(local)

GB = 1 << 30
h_in = torch.empty(GB, dtype=torch.uint8).pin_memory()
h_out = torch.empty(GB, dtype=torch.uint8).pin_memory()
d_a = torch.empty(GB, dtype=torch.uint8, device=dev)
d_b = torch.empty(GB, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def timed(fn, reps=4):
    fn()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    barrier()
    return e0.elapsed_time(e1) / reps


def h2d():
    d_a.copy_(h_in, non_blocking=True)


def d2h():
    h_out.copy_(d_b, non_blocking=True)


def both():
    s1.wait_stream(torch.cuda.current_stream())
    s2.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1)
    torch.cuda.current_stream().wait_stream(s2)


res = {'h2d_GBs': GB / timed(h2d) / 1e6, 'd2h_GBs': GB / timed(d2h) / 1e6, 'both_each_GBs': GB / timed(both) / 1e6}
vals = torch.tensor([res['h2d_GBs'], res['d2h_GBs'], res['both_each_GBs']], dtype=torch.float64, device=dev)
if world > 1:
    gathered = [torch.zeros_like(vals) for _ in range(world)]
    dist.all_gather(gathered, vals)
else:
    gathered = [vals]
if rank == 0:
    def sh(cmd):
        try:
            return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=20).stdout.strip()
        except Exception as exc:    # noqa: BLE001
            return 'unavailable: %s' % exc
    out = {
        'world': world, 'numa_bind': bound,
        'per_rank_GBs': [{'rank': r, 'h2d': round(float(g[0]), 1), 'd2h': round(float(g[1]), 1),
                          'h2d_and_d2h_at_once_each': round(float(g[2]), 1)} for r, g in enumerate(gathered)],
        'sum_h2d_GBs': round(sum(float(g[0]) for g in gathered), 1),
        'sum_d2h_GBs': round(sum(float(g[1]) for g in gathered), 1),
        'sum_both_each_GBs': round(sum(float(g[2]) for g in gathered), 1),
        'numa_nodes_online': sh('cat /sys/devices/system/node/online'),
        'cpus_allowed': sh("grep Cpus_allowed_list /proc/self/status"),
        'cpu_count': os.cpu_count(),
        'gpu_numa_nodes': numa.gpu_numa_nodes(),
        'lscpu_numa': sh("lscpu | grep -i numa"),
        'topo': sh('nvidia-smi topo -m'),
    }
    print(json.dumps(out, indent=1))
if world > 1:
    dist.destroy_process_group()
