import torch, time
dev=torch.device("cuda:0")
n=128*1024*1024
h1=torch.empty(n,dtype=torch.uint8).pin_memory(); h2=torch.empty(n,dtype=torch.uint8).pin_memory()
d1=torch.empty(n,dtype=torch.uint8,device=dev); d2=torch.empty(n,dtype=torch.uint8,device=dev)
s1,s2=torch.cuda.Stream(),torch.cuda.Stream()
def t(fn,reps=10):
    fn(); torch.cuda.synchronize(); t0=time.time()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.time()-t0)/reps
def h2d():
    with torch.cuda.stream(s1): d1.copy_(h1,non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h2.copy_(d2,non_blocking=True)
def both(): h2d(); d2h()
print("H2D GB/s", n/t(h2d)/1e9, "D2H GB/s", n/t(d2h)/1e9, "both: each GB/s", n/t(both)/1e9)
