"""Host->device bandwidth per rank when all ranks copy at once (run under torchrun on the GPU box)."""
import os, time, json
import torch, torch.distributed as dist
r = int(os.environ.get("LOCAL_RANK", 0)); w = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(r)
if w > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", r))
h = torch.empty(500_000_000, dtype=torch.uint8).pin_memory()
d = torch.empty_like(h, device="cuda")
for _ in range(2):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
if w > 1:
    dist.barrier()
t = time.perf_counter()
for _ in range(5):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t) / 5
out = torch.tensor([0.5 / dt], device="cuda")
if w > 1:
    lst = [torch.zeros_like(out) for _ in range(w)]
    dist.all_gather(lst, out)
    if r == 0:
        print(json.dumps({"ranks": w, "h2d_GBps_per_rank": [round(float(x), 1) for x in lst], "cpus": os.cpu_count()}))
    dist.destroy_process_group()
else:
    print(json.dumps({"ranks": 1, "h2d_GBps": round(float(out), 1)}))
