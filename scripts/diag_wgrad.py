import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dsgan_b200._lib import lib
L = lib()
def rel(a,b): return float((a.double()-b.double()).norm()/(b.double().norm()+1e-30))
for P,Co,Ci in [(64,128,64),(64,64,64),(16,128,64),(128,128,64),(256,128,64),(64,128,128),(64,256,256),(4096,64,128)]:
    g = torch.Generator().manual_seed(1)
    dY = torch.randn(P,Co,generator=g).bfloat16(); X = torch.randn(P,Ci,generator=g).bfloat16()
    ref = dY.float().t() @ X.float()
    dW = torch.zeros(Co,Ci,device='cuda')
    L.tc_wgrad(dY.cuda().data_ptr(), Co, X.cuda().data_ptr(), Ci, P, Co, Ci, dW.data_ptr(), Ci, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    out = dW.cpu()
    print(P,Co,Ci,'total %.3f'%rel(out,ref), 'rows0-63 %.3f'%rel(out[:64],ref[:64]), 'rows64+ %.3f'%(rel(out[64:],ref[64:]) if Co>64 else -1),
          'cols0-63 %.3f'%rel(out[:,:64],ref[:,:64]))
    # per-k-row contribution check: which k rows are used? use one-hot in k
    if P==64 and Co==128 and Ci==64:
        for kk in (0,1,7,8,15,16,17,63):
            dY1 = torch.zeros(P,Co).bfloat16(); dY1[kk]=dY[kk]
            ref1 = dY1.float().t() @ X.float()
            dW = torch.zeros(Co,Ci,device='cuda')
            L.tc_wgrad(dY1.cuda().data_ptr(), Co, X.cuda().data_ptr(), Ci, P, Co, Ci, dW.data_ptr(), Ci, torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize(); o=dW.cpu()
            # find which X row best explains output row 0
            best = max(range(P), key=lambda r: float((o[0]*X[r].float()).sum()/ (X[r].float().norm()*o[0].norm()+1e-9)))
            print('   one-hot k=%d: rel %.3f ; out row0 matches X row %d'%(kk, rel(o,ref1), best))
