import sys, numpy as np, torch
sys.path.insert(0, '.')
from einconv_b200 import evaluator
from oracle import direct
torch.manual_seed(0)
dev='cuda'
def show(name, got, want):
    got=got.double().cpu().numpy(); err=np.abs(got-want)
    print(name, 'shape', got.shape, 'max|ref|', np.abs(want).max(), 'max err', err.max(), 'got absmax', np.abs(got).max(), 'nonzero frac', (got!=0).mean())
    flat=got.reshape(got.shape[0],-1); w=want.reshape(want.shape[0],-1)
    print('  got[0,:8]', np.round(flat[0,:8],3)); print('  ref[0,:8]', np.round(w[0,:8],3))
    print('  got[1,:8]', np.round(flat[1,:8],3)); print('  ref[1,:8]', np.round(w[1,:8],3))
    print('  got[5,:8]', np.round(flat[5,:8],3)); print('  ref[5,:8]', np.round(w[5,:8],3))
for mode in ('tf32',):
    evaluator.set_fp32_math(mode)
    # tiny: 1d, k=1: gw[co, ci] = sum_pos x[ci,pos] v[co,pos]
    x=torch.rand(1,8,40); v=torch.rand(1,8,40)
    want=direct.conv_weight_vjp(x.numpy(), v.numpy(), (1,))
    got=evaluator.conv_weight_vjp_raw(x.to(dev), v.to(dev), (1,), 1, 0, 1, 1)
    show('wgrad 1x1 tf32', got.reshape(8,8), want.reshape(8,8))
    # structured: x[ci,pos] = ci+1 ; v[co,pos] = 1 if pos==0
    x=torch.zeros(1,8,40); v=torch.zeros(1,8,40)
    for ci in range(8): x[0,ci,:]=ci+1
    v[0,:,0]=1
    for co in range(8): v[0,co,0]=10*(co+1)
    want=direct.conv_weight_vjp(x.numpy(), v.numpy(), (1,))
    got=evaluator.conv_weight_vjp_raw(x.to(dev), v.to(dev), (1,), 1, 0, 1, 1)
    print('structured got\n', np.round(got.reshape(8,8).cpu().numpy(),1)); print('want\n', want.reshape(8,8))
    # kfc
    x=torch.rand(1,8,40)
    want=direct.kfc(x.numpy(), (1,))
    got=evaluator.kfc(x.to(dev), (1,))
    show('kfc 1x1 tf32', got[0], want[0])
