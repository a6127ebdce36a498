"""Latency of the per-iteration exchange: all-reduce of the packed statistics (141 doubles at N=10) + read-back, as
bench.py's em_step does it.  torchrun --nproc-per-node 2 tools/allreduce_probe.py"""
import os
import time

import torch
import torch.distributed as td

local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
td.init_process_group('nccl', device_id=dev)
x = torch.zeros(141, dtype=torch.float64, device=dev)
for n in (141, 141 * 8):
    x = torch.ones(n, dtype=torch.float64, device=dev)
    for _ in range(20):
        td.all_reduce(x)
        x.cpu()
    torch.cuda.synchronize()
    td.barrier()
    t0 = time.perf_counter()
    for _ in range(200):
        td.all_reduce(x)
        y = x.cpu()
    t1 = time.perf_counter()
    # with a busy GPU in between (a 2 ms kernel), as in the real loop
    a = torch.randn(4096, 4096, device=dev)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    for _ in range(50):
        b = a @ a
        td.all_reduce(x)
        y = x.cpu()
    t3 = time.perf_counter()
    for _ in range(50):
        b = a @ a
        y = x.cpu()
    t4 = time.perf_counter()
    if td.get_rank() == 0:
        print('n=%d: all_reduce + D2H %.1f us per call; after a kernel: %.1f us more than the kernel alone'
              % (n, (t1 - t0) / 200 * 1e6, ((t3 - t2) - (t4 - t3)) / 50 * 1e6), flush=True)
td.destroy_process_group()
