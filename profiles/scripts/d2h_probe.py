"""Concurrent device->host copy bandwidth per rank (torchrun): is the end-to-end number at N > 1
bound by the host platform?  Prints the PCIe/NUMA topology once and each rank's affinity."""
import os, subprocess, time
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
if rank == 0:
    print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout, flush=True)
before = len(os.sched_getaffinity(0))
note = ""
try:
    import pynvml
    pynvml.nvmlInit()
    pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
except Exception as e:  # noqa
    note = repr(e)
after = sorted(os.sched_getaffinity(0))
src = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
dst = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True)
for mode in ("alone", "together"):
    for r in range(world if mode == "alone" else 1):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if mode == "together" or r == rank:
            t0 = time.perf_counter()
            for _ in range(4):
                dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            gbs = 4 * (1 << 30) / (time.perf_counter() - t0) / 1e9
            print(f"rank {rank} {mode}: {gbs:.1f} GB/s  cpus {before}->{len(after)} [{after[0]}..{after[-1]}] {note}", flush=True)
        if world > 1:
            dist.barrier()
