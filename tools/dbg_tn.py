import torch, sys
sys.path.insert(0,'.')
from fieldconv_b200 import ops
DEV='cuda:0'
for (m,n,k) in [(128,96,32),(128,16,8),(128,32,64)]:
    g = torch.Generator().manual_seed(1)
    a = torch.randn(k, m, generator=g).to(DEV); b = torch.randn(k, n, generator=g).to(DEV)
    for mode in (2,1):
        c = ops.gemm(a,b,True,mode); torch.cuda.synchronize()
        ref = (a.double().t() @ b.double()).float()
        print(m,n,k,mode,'zeros frac', float((c==0).float().mean()), 'absmax', float(c.abs().max()), 'ref absmax', float(ref.abs().max()), 'err', float((c-ref).abs().max()))
        # is c a permutation/transposition of ref?
        if k<=64:
            print(' c[0,:6]', c[0,:6].tolist()); print(' ref[0,:6]', ref[0,:6].tolist()); print(' c[:6,0]', c[:6,0].tolist()); print(' ref[:6,0]', ref[:6,0].tolist())
