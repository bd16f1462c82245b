"""Time a 25 MB fp64 all-reduce (the sample-summed C2 gradient) under the current NCCL settings."""
import os, torch, torch.distributed as dist
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
t = torch.ones(3 * 1024 * 1024 + 8, dtype=torch.float64, device="cuda")
for _ in range(10): dist.all_reduce(t)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 50
a.record()
for _ in range(n): dist.all_reduce(t)
b.record(); torch.cuda.synchronize()
if dist.get_rank() == 0:
    print("ALGO=%s PROTO=%s  %.1f us per 25 MB all-reduce" % (os.environ.get("NCCL_ALGO"), os.environ.get("NCCL_PROTO"), a.elapsed_time(b) / n * 1e3))
dist.destroy_process_group()
