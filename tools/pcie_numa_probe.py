"""Host <-> device copy rates of all ranks at once, with and without binding each rank (and its pinned buffers) to the NUMA
node of its GPU.  usage: torchrun --nproc-per-node N tools/pcie_numa_probe.py"""
import os, sys, time, glob
import torch, torch.distributed as dist
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
def gpu_numa(lr):
    bus = torch.cuda.get_device_properties(lr).pci_bus_id if hasattr(torch.cuda.get_device_properties(lr), "pci_bus_id") else None
    try:
        import pynvml
        pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(lr); b = pynvml.nvmlDeviceGetPciInfo(h).busId
        b = b.decode() if isinstance(b, bytes) else b
        p = "/sys/bus/pci/devices/" + b.lower()[-12:] + "/numa_node"
        return int(open(p).read()), b
    except Exception as e:
        return -1, str(e)
def bind(node):
    try:
        cpus = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
        s = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-"); s.update(range(int(a), int(b or a) + 1))
        os.sched_setaffinity(0, s); return len(s)
    except Exception as e:
        return str(e)
def measure(tag):
    n = 128 << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory(); h.fill_(1)
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    out = {}
    for name, f in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
        f(); torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        for _ in range(5): f()
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
        out[name] = n / dt / 1e9
        dist.barrier()
    t = torch.tensor([out["h2d"], out["d2h"]], device="cuda"); lo = t.clone(); dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(t, op=dist.ReduceOp.SUM)
    if rank == 0: print(f"{tag}: aggregate H2D {t[0].item():.0f} GB/s (slowest rank {lo[0].item():.1f}), D2H {t[1].item():.0f} GB/s (slowest rank {lo[1].item():.1f})", flush=True)
node, bus = gpu_numa(lr)
info = [None] * world
dist.all_gather_object(info, (rank, node, bus, len(os.sched_getaffinity(0))))
if rank == 0:
    print("rank, gpu numa node, bus, cpus allowed:", info, flush=True)
    os.system("lscpu | grep -i 'numa\\|socket\\|model name' | head -8")
measure("unbound")
if node >= 0:
    r = bind(node)
    measure(f"bound to the GPU's NUMA node (rank 0: node {node}, {r} cpus)")
dist.destroy_process_group()
