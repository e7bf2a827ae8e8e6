import sys, torch
sys.path.insert(0, '.')
from codlad_b200 import synthetic, engine
from oracle import restate as R
for (L,F,K,seed) in [(64,2,64,11),(300,3,64,12)]:
    X = synthetic.ca_trace(F, L, seed)
    D_ref, I_ref = R.knn_graph(X, torch.ones(F, L), K)
    D, I = engine.knn_topk(X.cuda(), None, K)
    D = D.cpu(); I = I.cpu().long()
    ne = (D != D_ref)
    print(L, 'mismatch D', int(ne.sum()), 'of', D.numel(), 'idx mismatch', int((I != I_ref).sum()))
    if ne.any():
        w = ne.nonzero()[:10]
        for b,i,k in w.tolist():
            print(b,i,k, D[b,i,k].item(), D_ref[b,i,k].item(), (D[b,i,k].view(torch.int32)-D_ref[b,i,k].view(torch.int32)).item(), I[b,i,k].item(), I_ref[b,i,k].item())
    # recompute on GPU with torch ops
    Xc = X.cuda()
    d = Xc[:, None, :, :] - Xc[:, :, None, :]
    sq = d*d
    s = (sq[...,0]+sq[...,1])+sq[...,2]
    Dg = torch.sqrt(s+1e-6).cpu()
    d = X[:, None, :, :] - X[:, :, None, :]
    sq = d*d
    s = (sq[...,0]+sq[...,1])+sq[...,2]
    Dc = torch.sqrt(s+1e-6)
    print('torch gpu vs cpu full matrix mismatches', int((Dg != Dc).sum()))
    s2 = s + 1e-6
    s3 = s + torch.tensor(1e-6, dtype=torch.float32)
    print('scalar add variants differ', int((s2 != s3).sum()))
    print('sqrt cpu vs gpu on same input', int((torch.sqrt(s2) != torch.sqrt(s2.cuda()).cpu()).sum()))
