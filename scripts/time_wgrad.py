import sys, time, torch
sys.path.insert(0, '.')
from einconv_b200 import evaluator
dev='cuda'
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    s=torch.cuda.Event(enable_timing=True); e=torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); e.synchronize()
    return s.elapsed_time(e)/n
cases = {
 'A: (14336,64,112) k3  [channel stride 224 B]': ((14336,64,112),(3,),(1,)),
 'B: (8,64,16,112,112) k(1,1,3) [channel stride 401 KB]': ((8,64,16,112,112),(1,1,3),(0,0,1)),
 'C: (8,64,16,112,112) k(3,3,3)': ((8,64,16,112,112),(3,3,3),(1,1,1)),
 'D: (8,64,16,112,112) k(1,1,1)': ((8,64,16,112,112),(1,1,1),(0,0,0)),
}
for name,(xs,k,p) in cases.items():
    x=torch.rand(*xs,device=dev).bfloat16(); v=torch.rand(*xs,device=dev).bfloat16()
    ms=t(lambda: evaluator.conv_weight_vjp_raw(x,v,k,1,p,1,1))
    flops=2*x.numel()*64*1.0
    kt=1
    for kk in k: kt*=kk
    print(f'{name}: {ms:.3f} ms  {flops*kt/ms/1e9:.1f} TFLOP/s')
