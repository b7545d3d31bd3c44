import os, torch, torch.distributed as dist, time
r=int(os.environ["RANK"]); w=int(os.environ["WORLD_SIZE"]); l=int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(l); dev=torch.device(f"cuda:{l}")
dist.init_process_group("nccl", device_id=dev)
for n in (26_000_000, 2_000_000, 260_000):
    x=torch.ones(n,device=dev)
    for _ in range(5): dist.all_reduce(x)
    torch.cuda.synchronize(); dist.barrier(); 
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): dist.all_reduce(x)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/20
    if r==0: print(f"allreduce {n*4/1e6:.1f} MB: {ms:.3f} ms  algbw {n*4/ms/1e6:.1f} GB/s", flush=True)
if r==0: print("p2p access 0->1:", torch.cuda.can_device_access_peer(0,1))
dist.destroy_process_group()
