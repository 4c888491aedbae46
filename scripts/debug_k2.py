import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch
from helpers import *
from oracle import pin_oracle as po
from pin_slam_b200 import ops

def run(F,K,L,wf,oc,n):
    m = synthetic_map(n_surface=30000, seed=F + K + L, resolution=0.4, buffer_size=200003, feature_dim=F,
                      after_pgo=False, local_radius=14.0, diff_td=3.0)
    sig = oc > 1
    dec = po.make_decoder(F + 3, 64, L, oc, 0.044, seed=L)
    q = queries_near(m, n, seed=9)
    dl = torch.randn(n, oc, generator=torch.Generator().manual_seed(1))
    def reference(mm, dd, qq, dll):
        mm.local_geo_features.requires_grad_(True); dd.requires_grad_(True)
        vec, _, w, _, _ = po.query_feature(mm, qq, None, K, wf, training_mode=False)
        if wf:
            out = po.decoder_color(dd, vec) if sig else po.decoder_sdf(dd, vec).unsqueeze(1)
        else:
            flat = vec.reshape(-1, F + 3)
            o = po.decoder_color(dd, flat) if sig else po.decoder_sdf(dd, flat).unsqueeze(1)
            out = (o.view(qq.shape[0], K, oc) * w).sum(1)
        (out * dll).sum().backward()
        return mm.local_geo_features.grad, torch.cat([p.grad.reshape(-1) for p in dd.tensors()])
    gf_ref, gd_ref = reference(m.clone(), dec.clone(), q, dl)
    mh = map_handle_from_oracle(m, True); dh = decoder_handle_from_oracle(dec, sigmoid_out=sig)
    idx, _, w, _ = ops.knn_search(mh, q.cuda(), K)[:4]
    gfeat = torch.zeros_like(mh.keep["geo_feat"]); gdec = torch.zeros(dh.param_count(), device="cuda")
    ops.train_backward(mh, dh, mh.keep["geo_feat"], q.cuda(), idx, w, dl.cuda(), wf, gfeat, gdec)
    ef = (gfeat.cpu()-gf_ref).abs(); ed=(gdec.cpu()-gd_ref).abs()
    print(f"F{F} K{K} L{L} wf{int(wf)} oc{oc} n{n}: gfeat err {ef.max():.2e}/{gf_ref.abs().max():.2e} badrows {(ef.max(1)[0] > 1e-3*gf_ref.abs().max()).sum()}  gdec err {ed.max():.2e}/{gd_ref.abs().max():.2e}", flush=True)

for cfg in [(16,5,2,False,1,7001),(16,5,2,False,1,1500),(16,5,1,False,1,7001),(16,6,2,False,1,7001),(8,5,2,False,1,7001),(16,5,2,True,1,7001),
            (8,6,1,False,1,40000),(32,8,2,True,1,30000),(64,8,1,False,1,7001),(64,3,2,True,3,7001),(8,6,2,False,3,7001)]:
    run(*cfg)
