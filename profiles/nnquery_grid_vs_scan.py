import sys, os, torch, time
sys.path.insert(0, '/root/repo'); os.chdir('/root/repo')
import sph3d_gcn_b200 as S
def t(fn, it=5):
    fn(); torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/it
g=torch.Generator().manual_seed(1)
for (B,N,K,r) in ((32,10000,64,0.1),(32,10000,64,0.1451),(8,8192,64,0.1),(4,65536,64,0.05),(4,65536,64,0.03)):
    xyz=torch.rand(B,N,3,generator=g).cuda()
    res={}
    for grid in ('1','0'):
        os.environ['SPH3D_NNQUERY_GRID']=grid
        ms=t(lambda: S.tf_nnquery.build_sphere_neighbor(xyz,xyz,radius=r,nnsample=K))
        idx,cnt,dst=S.tf_nnquery.build_sphere_neighbor(xyz,xyz,radius=r,nnsample=K)
        res[grid]=(ms,idx.clone(),cnt.clone(),dst.clone())
    same=all(torch.equal(res['1'][i],res['0'][i]) for i in (1,2,3))
    print(f"B={B} N={N} K={K} r={r}: grid {res['1'][0]:.3f} ms, scan {res['0'][0]:.3f} ms, identical={same}, mean cnt {res['1'][2].float().mean():.1f}")
