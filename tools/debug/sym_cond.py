"""Numerics experiment behind DESIGN.md 3.2a: posterior variance through the L-form (k** - ||Linv k*||^2) vs the symmetric
single-pass form (k** - k* . Kinv k*) against a long-double forward substitution, on the synthetic benchmark GP and on a BO-like
training set with clustered points.  CPU only (numpy / torch); nothing here is on the product path."""
import torch, math, sys
sys.path.insert(0,'/root/repo')
from bench_workload import make_problem, _matern52, D_CONT
import numpy as np
torch.manual_seed(0)
def run(n, noise, cluster=False, near=1e-3):
    g = torch.Generator().manual_seed(0)
    X = torch.rand(n, 13, generator=g, dtype=torch.float64)
    if cluster:  # BO-like: half the points cluster around one optimum
        X[n//2:, :3] = 0.5 + 0.01*torch.randn(n - n//2, 3, generator=g, dtype=torch.float64)
    X[:, 3:] = X[:, 3:].round()
    if cluster: X[n//2:, 3:] = X[n//2, 3:]
    ls_c = torch.tensor([0.4,0.6,0.7], dtype=torch.float64); ls_b = torch.full((10,),1.2,dtype=torch.float64)
    def kern(A,B): return 1.3*_matern52(A[:, :3],B[:, :3],ls_c)*_matern52(A[:,3:],B[:,3:],ls_b)
    K = kern(X,X); K.diagonal().add_(noise)
    ev = torch.linalg.eigvalsh(K)
    L = torch.linalg.cholesky(K)
    Linv = torch.linalg.solve_triangular(L, torch.eye(n,dtype=torch.float64), upper=False)
    Kinv = Linv.T @ Linv
    Kinv = 0.5*(Kinv+Kinv.T)
    # test points: random + near training points
    m = 512
    Xt = torch.rand(m,13,generator=g,dtype=torch.float64); Xt[:,3:] = Xt[:,3:].round()
    idx = torch.randint(0,n,(m,),generator=g)
    Xn = X[idx].clone(); Xn[:, :3] += near*torch.randn(m,3,generator=g,dtype=torch.float64)
    for name, P in (("random",Xt),("near",Xn)):
        ks = kern(P, X)  # m x n
        # truth in long double via numpy
        Lq = L.numpy().astype(np.longdouble); kq = ks.numpy().astype(np.longdouble)
        import scipy.linalg
        # forward substitution in longdouble (vectorized over rhs)
        v = np.zeros((n,m),dtype=np.longdouble)
        for i in range(n):
            v[i] = (kq[:,i] - Lq[i,:i] @ v[:i]) / Lq[i,i]
        truth = 1.3 - (v*v).sum(0)
        truth = torch.tensor(truth.astype(np.float64))
        vL = ks @ Linv.T
        s_tri = 1.3 - (vL*vL).sum(-1)
        w = ks @ Kinv
        s_sym = 1.3 - (ks*w).sum(-1)
        for nm, s in (("tri",s_tri),("sym",s_sym)):
            rel = ((s-truth).abs()/truth.abs()).max().item()
            ab = (s-truth).abs().max().item()
            print(f"n={n} noise={noise:g} cluster={cluster} cond={ev[-1]/ev[0]:.3g} {name:6s} {nm}: max rel {rel:.3g} max abs {ab:.3g} min var {truth.min().item():.3g}")
run(1000,1e-6)
run(1000,1e-6,cluster=True)
run(300,1e-6,cluster=True)
