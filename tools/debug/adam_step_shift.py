"""What ONE Adam step does to the policy's mean action (debugging aid, CPU only).

A 48-env oracle rollout of LidarSpread n = 3 with the initial policy; the PPO surrogate with (a) constant advantages
-8, (b) N(0, 1) noise, each with the reference's zero-carry chunks and with the true carries (ratio == 1); then the
first Adam step (lr * sign(g)) and the change of the mean action over all samples.  Result on record: the shift is
0.03 common to all agents in every case - the size of the step does not depend on the advantage signal."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from dgppo_b200.algo import params as P, update as U
from oracle import algo_np, env_np, nn_np, reset_np
F=np.float32
N,N_OBS,T,RS,B=3,3,128,16,48
cfg=env_np.EnvCfg(env_np.LIDAR_SPREAD,n=N,n_obs=N_OBS,max_step=T)
rays=env_np.ray_table(cfg.n_rays,cfg.comm_radius)
rng=np.random.default_rng(0)
pol=P.init_policy_params(cfg.node_dim,cfg.edge_dim,2,2,seed=0)
ag,gl,recs=zip(*[reset_np.reset_states(cfg,int(k))[:3] for k in rng.integers(0,2**31-1,size=B)])
rec=np.stack(recs)
obst=dict(center=rec[...,0:2],width=rec[...,2],height=rec[...,3],theta=rec[...,4],cos=rec[...,5],sin=rec[...,6],points=rec[...,8:16].reshape(rec.shape[:-1]+(4,2)))
g0=env_np.reset_graph(cfg,np.stack(ag),np.stack(gl),obst,None,rays)
eps=rng.standard_normal((B,T,N,2)).astype(F)
ro=algo_np.rollout(cfg,pol,g0,obst,eps,T,rays=rays)
gi=U.GraphIndex(N,cfg.n_ag,cfg.n_ao,cfg.n_nodes,torch.device("cpu"))
tt=lambda a: torch.tensor(np.ascontiguousarray(a))
a=[tt(ro[k][:,:T]) for k in ("nodes","edges","receivers","senders")]
g=U.prep_graphs(a[0].reshape((B*T,)+a[0].shape[2:]),a[1].reshape((B*T,)+a[1].shape[2:]),a[2].reshape(B*T,-1),a[3].reshape(B*T,-1),gi,torch.float32)
st=U.NetTrainState(pol,"cpu",3e-4)
def loss_policy(params,adv,init_carry):
    C=T//RS
    emb=U.gnn(params["params"]["PolicyNet_0"]["GraphTransformerGNN_0"],g,gi,2).reshape(B,C,RS,N,-1)
    act=tt(ro["actions"]).reshape(B,C,RS,N,2)
    h=init_carry
    lps=[]; means=[]
    for t in range(RS):
        mean,std,h=U.policy_step(params,emb[:,:,t],h)
        lps.append(U.tanh_normal_log_prob(act[:,:,t],mean,std)); means.append(mean)
    lp=torch.stack(lps,2).reshape(B,T,N)
    ratio=torch.exp(lp-tt(ro["log_pis"]))
    l1=-ratio*adv; l2=-torch.clamp(ratio,0.75,1.25)*adv
    return torch.maximum(l1,l2).mean(), torch.stack(means,2).reshape(B,T,N,2), ratio
zero=torch.zeros((B,T//RS,N,64))
true=tt(ro["rnn_states"][:,:T].reshape(B,T//RS,RS,N,64)[:,:,0])
for name,carry in (("zero-carry chunks (reference)",zero),("true carries",true)):
    for Aname,adv in (("A=-8 const",torch.full((B,T,N),-8.0)),("A=N(0,1) noise",torch.randn((B,T,N)))):
        loss,means,ratio=loss_policy(st.tree(),adv,carry)
        (gr,)=torch.autograd.grad(loss,[st.flat])
        # effect of one sign-step of Adam (first step = lr*sign(g)) on the mean action of the deterministic policy
        with torch.no_grad():
            old=st.flat.clone(); st.flat-=3e-4*torch.sign(gr)
        _,means2,_=loss_policy(st.tree(),adv,carry)
        with torch.no_grad(): st.flat.copy_(old)
        dm=(means2-means).detach()
        print(f"{name:32s} {Aname:16s} |grad| {float(gr.norm()):.4f}  mean|ratio-1| {float((ratio-1).abs().mean()):.4f}  d(mean action) after one sign step: mean {dm.mean(dim=(0,1)).numpy().round(4).tolist()} rms {float(dm.pow(2).mean().sqrt()):.4f}")
