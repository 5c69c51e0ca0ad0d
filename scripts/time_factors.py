import sys, torch
sys.path.insert(0, '.')
from einconv_b200 import evaluator
from bench import RESNET50_LAYERS
dev='cuda'
evaluator.set_fp32_math(sys.argv[1] if len(sys.argv)>1 else 'tf32')
BATCH=int(sys.argv[2]) if len(sys.argv)>2 else 256
def t(fn, n=3):
    fn(); torch.cuda.synchronize()
    s=torch.cuda.Event(enable_timing=True); e=torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); e.synchronize()
    return s.elapsed_time(e)/n
tk=tr=0
print(f"{'layer':28s} {'count':>5s} {'kfc ms':>8s} {'TF/s':>7s} {'kfacr ms':>9s} {'GB/s':>7s}")
for c,h,k,s,p,count in RESNET50_LAYERS:
    x=torch.rand(BATCH,c,h,h,device=dev)
    o=(h+2*p-k)//s+1; a=c*k*k
    mk=t(lambda: evaluator.kfc(x,k,stride=s,padding=p))
    mr=t(lambda: evaluator.kfac_reduce(x,k,stride=s,padding=p))
    fl=2*BATCH*o*o*a*a
    print(f"C{c:<5d}H{h:<4d}k{k}s{s}p{p} A={a:<6d} {count:5d} {mk:8.3f} {fl/mk/1e9:7.1f} {mr:9.3f} {x.numel()*4/mr/1e6:7.1f}")
    tk+=mk*count; tr+=mr*count
print('total kfc ms', tk, 'kfac_reduce ms', tr)
