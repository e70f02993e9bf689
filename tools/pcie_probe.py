#!/usr/bin/env python
"""Host <-> device ceiling of the box, per GPU and in aggregate, with 1 / 2 / 4 / 8 GPUs copying at once.

The e2e leg of bench.py moves 8 MiB per block in each direction, so its bound is
min(H2D, D2H under concurrent load) / 8 MiB per GPU; VERDICT r01 found e2e flat from 2 to 8 GPUs.
This probe separates the platform from the library: pinned buffers allocated three ways
(cudaHostAllocDefault, cudaHostAllocWriteCombined for the H2D source, cudaHostAllocPortable) and, when
the guest exposes NUMA nodes, with the process pinned to the GPU's node before allocating.

    python tools/pcie_probe.py                                   # one GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_probe.py

Prints one JSON line (rank 0): per mode, the per-rank GB/s of H2D alone, D2H alone and both at once,
all ranks copying concurrently, and the sums."""
import ctypes as C
import json
import os

import torch
import torch.distributed as dist

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rt = C.CDLL("libcudart.so.12")
N = 1 << 30
FLAGS = {"default": 0, "portable": 1, "write_combined_h2d_source": 4}


def numa_of_gpu():
    try:
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = torch.cuda.get_device_properties(local).pci_domain_id
        dev = torch.cuda.get_device_properties(local).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        return int(open(path).read())
    except Exception:
        return None


def host_alloc(flags):
    p = C.c_void_p()
    assert rt.cudaHostAlloc(C.byref(p), C.c_size_t(N), C.c_uint(flags)) == 0
    C.memset(p, 1, N)
    return p


d_a = torch.empty(N, dtype=torch.uint8, device="cuda")
d_b = torch.empty(N, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h_in, h_out, do_in, do_out, reps=4):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    for _ in range(reps):
        if do_in:
            assert rt.cudaMemcpyAsync(C.c_void_p(d_a.data_ptr()), h_in, C.c_size_t(N), 1, C.c_void_p(s1.cuda_stream)) == 0
        if do_out:
            assert rt.cudaMemcpyAsync(h_out, C.c_void_p(d_b.data_ptr()), C.c_size_t(N), 2, C.c_void_p(s2.cuda_stream)) == 0
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    e1.record()
    torch.cuda.synchronize()
    return reps * N / (e0.elapsed_time(e1) * 1e-3) / 1e9


def gather(v):
    t = torch.tensor([v], device="cuda")
    if world == 1:
        return [v]
    out = [torch.zeros(1, device="cuda") for _ in range(world)]
    dist.all_gather(out, t)
    return [float(x.item()) for x in out]


res = {"n_gpus": world, "gib_per_copy": 1, "gpu_numa_node": numa_of_gpu(),
       "host_numa_nodes": len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")]) if os.path.isdir("/sys/devices/system/node") else None,
       "cpus": os.cpu_count(), "modes": {}}
node = res["gpu_numa_node"]
for mode, flag in FLAGS.items():
    h_in = host_alloc(flag)
    h_out = host_alloc(0 if mode == "write_combined_h2d_source" else flag)   # write-combined memory is for device READS only
    run(h_in, h_out, True, True, 1)
    r = {"h2d": gather(run(h_in, h_out, True, False)), "d2h": gather(run(h_in, h_out, False, True)),
         "both_each_direction": gather(run(h_in, h_out, True, True))}
    r["sum"] = {k: sum(v) for k, v in r.items()}
    res["modes"][mode] = r
    rt.cudaFreeHost(h_in); rt.cudaFreeHost(h_out)
if node is not None and node >= 0:
    try:  # pin to the GPU's node and allocate again (first touch by this process)
        cpus = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
        ids = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids.update(range(int(a), int(b or a) + 1))
        os.sched_setaffinity(0, ids)
        h_in, h_out = host_alloc(0), host_alloc(0)
        r = {"h2d": gather(run(h_in, h_out, True, False)), "d2h": gather(run(h_in, h_out, False, True)),
             "both_each_direction": gather(run(h_in, h_out, True, True))}
        r["sum"] = {k: sum(v) for k, v in r.items()}
        res["modes"]["default_pinned_to_gpu_numa_node"] = r
    except Exception as e:
        res["modes"]["default_pinned_to_gpu_numa_node"] = {"error": str(e)}
best = max(res["modes"].values(), key=lambda m: m.get("sum", {}).get("both_each_direction", 0))
res["e2e_blocks_per_s_bound_n4096_all_gpus"] = best["sum"]["both_each_direction"] * 1e9 / (8 << 20)
if rank == 0:
    print(json.dumps(res))
if world > 1:
    dist.destroy_process_group()
