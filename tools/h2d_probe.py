"""Host -> device feed of the end-to-end leg, per GPU ALONE and with all GPUs CONCURRENTLY (the e2e number of bench.py is
PCIe bound: 8.6 GB per step from page-locked memory).  Launch like the bench:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/h2d_probe.py
Every rank pins 2 GiB, copies it to its GPU 6 times (CUDA events); first one rank at a time (the others idle), then all at
once.  Rank 0 prints one JSON object and writes gpurun_out/h2d_probe.json; the topology the box reports goes next to it.
If the concurrent per-GPU rate collapses while the solo rate does not, the box's host memory / PCIe fabric is the cap of
the e2e scaling, not the library (there is no collective and no shared buffer on this path)."""
import json
import os
import subprocess
import sys

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)


def barrier():
    if world > 1:
        dist.barrier(device_ids=[local])
    torch.cuda.synchronize()


nbytes = 2 << 30
h = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
h.fill_(rank + 1)          # first touch by this rank's thread
d = torch.empty(nbytes, dtype=torch.uint8, device=dev)


def gbs(reps=6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    return nbytes * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9


solo = 0.0
for r in range(world):
    barrier()
    if r == rank:
        solo = gbs()
barrier()
conc = gbs()
barrier()
vals = torch.tensor([solo, conc], dtype=torch.float64, device=dev)
allv = [torch.zeros_like(vals) for _ in range(world)]
if world > 1:
    dist.all_gather(allv, vals)
else:
    allv = [vals]
if rank == 0:
    rows = [{"gpu": i, "solo_gbs": round(float(v[0]), 2), "concurrent_gbs": round(float(v[1]), 2)} for i, v in enumerate(allv)]
    out = {"n_gpus": world, "bytes_per_copy": nbytes, "per_gpu": rows, "sum_solo_gbs": round(sum(r["solo_gbs"] for r in rows), 1),
           "sum_concurrent_gbs": round(sum(r["concurrent_gbs"] for r in rows), 1),
           "cpu_affinity_of_rank0": sorted(os.sched_getaffinity(0))[:4] + ["..."] + [len(os.sched_getaffinity(0))],
           "numa_nodes": sorted(x for x in os.listdir("/sys/devices/system/node") if x.startswith("node")) if os.path.isdir("/sys/devices/system/node") else None}
    try:
        out["topo"] = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout.replace("\x1b[4m", "").replace("\x1b[0m", "")[:4000]
    except Exception as exc:
        out["topo"] = repr(exc)
    print(json.dumps(out))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(f"gpurun_out/h2d_probe_n{world}.json", "w"), indent=1)
if world > 1:
    dist.destroy_process_group()
