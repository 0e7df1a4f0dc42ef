import sys, os
sys.path[:0]=[os.path.dirname(os.path.dirname(os.path.abspath(__file__)))+'/mcmc-ammsb-gpu_b200',os.path.dirname(os.path.dirname(os.path.abspath(__file__)))+'/oracle',os.path.dirname(os.path.dirname(os.path.abspath(__file__)))+'/tests']
import numpy as np, pymcmc, pyoracle, pyammsb as A
from util import make_edges, rel_err
from test_gpu_learner import OracleLearner, make_cfg
orc=pyoracle.Oracle()
N,K,n=1200,64,16
for resync_until in (0, 12):
    cfg=make_cfg(N=N,K=K,n=n,strategy="Node")
    lrn=pymcmc.Learner(cfg,0); ol=OracleLearner(orc,cfg,lrn,N,K,n)
    print("resync_until",resync_until, "ppx", lrn.heldout_perplexity(), ol.perplexity())
    for it in range(40):
        edges,nodes,nbrs,w=lrn.peek(n)
        want=ol.iterate(edges,nodes,w)
        assert np.array_equal(nbrs,want)
        lrn.run(1)
        pi,phi,beta,theta=lrn.read(N,K)
        epi=rel_err(pi,ol.pi); ephi=rel_err(phi,ol.phi); eth=rel_err(theta,ol.theta)
        mask=np.ones(N,bool); mask[nodes]=False
        print("it %2d V=%5d E=%5d w=%8.1f  pi max %.2e (untouched max %.2e, rows>1e-4: %d) phi max %.2e theta max %.2e frac>1e-5 %.3f"%(
            it,len(nodes),len(edges),w,epi.max(),epi[mask].max() if mask.any() else 0,(epi.max(axis=1)>1e-4).sum(),ephi.max(),eth.max(),(eth>1e-5).mean()))
        if it<resync_until: ol.pi,ol.phi,ol.beta,ol.theta=pi,phi,beta,theta
    print("final ppx", lrn.heldout_perplexity(), ol.perplexity())
    lrn.close(); cfg.close()
