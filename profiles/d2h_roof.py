#!/usr/bin/env python
"""Device -> pinned-host copy roof of the box, all ranks copying at once (the roof of bench.py's e2e path), with and without
binding every rank to the CPUs next to its GPU before the pinned buffer is allocated (first touch decides the NUMA node).
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 profiles/d2h_roof.py
Prints aggregate GB/s for: plain pinned memory / CPU-affinity-bound allocation; 27.8 MB chunks on two streams (what ssd_step_host
issues) and one 222 MB copy."""
import os
import re
import subprocess
import sys
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)


def gpu_cpus(index):
    """CPU list of `nvidia-smi topo -m` for this GPU, e.g. '0-15,32-47' -> set of ints."""
    try:
        out = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
        hdr = [l for l in out.splitlines() if "CPU Affinity" in l][0]
        col = hdr.split("\t").index([c for c in hdr.split("\t") if "CPU Affinity" in c][0])
        row = [l for l in out.splitlines() if l.startswith("GPU%d\t" % index) or l.startswith("GPU%d " % index)][0]
        spec = row.split("\t")[col].strip()
        cpus = set()
        for part in spec.split(","):
            m = re.match(r"(\d+)-(\d+)$", part)
            if m:
                cpus.update(range(int(m.group(1)), int(m.group(2)) + 1))
            elif part.isdigit():
                cpus.add(int(part))
        return cpus, spec
    except Exception as exc:  # noqa: BLE001
        return set(), "unavailable: %r" % (exc,)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)


def measure(nbytes, chunks, reps=10):
    src = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    dst = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    dst.fill_(1)  # touch every page from this (possibly bound) thread
    streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
    cuts = [(i * nbytes // chunks, (i + 1) * nbytes // chunks) for i in range(chunks)]

    def once():
        for k, (lo, hi) in enumerate(cuts):
            with torch.cuda.stream(streams[k & 1]):
                dst[lo:hi].copy_(src[lo:hi], non_blocking=True)
        for s in streams:
            s.synchronize()
    for _ in range(3):
        once()
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    return world * nbytes * reps / float(dt.item()) / 1e9


NB = 65536 * (3375 + 20)
cpus, spec = gpu_cpus(local)
res = {"plain_8chunks": measure(NB, 8), "plain_1copy": measure(NB, 1)}
if cpus:
    os.sched_setaffinity(0, cpus)
    res["bound_8chunks"] = measure(NB, 8)
    res["bound_1copy"] = measure(NB, 1)
if rank == 0:
    print("ranks %d  GPU0 cpu affinity %s  host cpus %d" % (world, spec, os.cpu_count()))
    for k, v in res.items():
        print("  %-16s %.1f GB/s aggregate" % (k, v))
    sys.stdout.flush()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
