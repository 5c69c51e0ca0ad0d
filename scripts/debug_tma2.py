import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from einconv_b200 import evaluator, _native as nat
from oracle import direct
torch.manual_seed(0)
dev='cuda'
# args: mode B C H W Cout K pad stride
mode=sys.argv[1]; B,C,H,W,Co,K,pad,stride=map(int, sys.argv[2:10])
dtype = torch.bfloat16 if mode=='bf16' else torch.float32
evaluator.set_fp32_math('tf32' if mode=='tf32' else 'auto')
x=(torch.rand(B,C,H,W)-0.3).to(dtype)
Ho=(H+2*pad-K)//stride+1; Wo=(W+2*pad-K)//stride+1
v=(torch.rand(B,Co,Ho,Wo)-0.5).to(dtype)
want=direct.conv_weight_vjp(x.double().numpy(), v.double().numpy(), (K,K), padding=pad, stride=stride)
try:
    got=evaluator.conv_weight_vjp_raw(x.to(dev), v.to(dev), (K,K), 1, pad, stride, 1)
    torch.cuda.synchronize()
    got=got.double().cpu().numpy()
    err=np.abs(got-want)
    print(sys.argv[1:], 'max ref %.3f max err %.4f' % (np.abs(want).max(), err.max()), 'per-tap max err', np.round(err.reshape(Co,C,K*K).max(axis=(0,1)),3))
except Exception as e:
    print(sys.argv[1:], 'ERR', str(e)[:150])
