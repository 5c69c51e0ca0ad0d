import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from einconv_b200 import evaluator, _native as nat
from oracle import direct
torch.manual_seed(0)
dev='cuda'
mode = sys.argv[1] if len(sys.argv)>1 else 'bf16'
dtype = torch.bfloat16 if mode=='bf16' else torch.float32
evaluator.set_fp32_math('tf32' if mode=='tf32' else 'auto')
import os
W=int(os.environ.get('DBG_W','16'))
x=(torch.rand(1,8,8,W)-0.3).to(dtype); v=(torch.rand(1,8,8,W)-0.5).to(dtype)
want=direct.conv_weight_vjp(x.double().numpy(), v.double().numpy(), (3,3), padding=1)
g = nat.make_geom(1,1,8,8,(8,W),(3,3),(1,1),(1,1),(1,1),(8,W))
print('kernel:', nat.lib().einconv_selected_kernel(nat.OP_CONV_WEIGHT_VJP, g, nat.dtype_code(dtype), evaluator._math_for(dtype), 0))
t=time.time()
try:
    got=evaluator.conv_weight_vjp_raw(x.to(dev), v.to(dev), (3,3), 1, 1, 1, 1)
    torch.cuda.synchronize()
    print('time', time.time()-t)
    got=got.double().cpu().numpy()
    print('max ref', np.abs(want).max(), 'max err', np.abs(got-want).max())
    print(np.round(got[0,0],3)); print(np.round(want[0,0],3))
except Exception as e:
    print('time', time.time()-t, 'ERR', str(e)[:200])
