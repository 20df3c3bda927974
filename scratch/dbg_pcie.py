import torch, time
n=64*1024*1024//8
h1=torch.zeros(n,dtype=torch.float64).pin_memory(); h2=torch.zeros(n,dtype=torch.float64).pin_memory()
d1=torch.zeros(n,dtype=torch.float64,device='cuda'); d2=torch.ones(n,dtype=torch.float64,device='cuda')
s1=torch.cuda.Stream(); s2=torch.cuda.Stream()
def t(f,reps=10):
    torch.cuda.synchronize(); a=time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter()-a)/reps
def up():
    with torch.cuda.stream(s1): d1.copy_(h1,non_blocking=True)
def down():
    with torch.cuda.stream(s2): h2.copy_(d2,non_blocking=True)
def both(): up(); down()
for f,name in ((up,'H2D'),(down,'D2H'),(both,'both')):
    f(); dt=t(f); print(name, "%.2f ms  %.1f GB/s (per direction)"%(dt*1e3, 64*1.048576/1e3/dt))
