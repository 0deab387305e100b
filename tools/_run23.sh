python - <<'PY'
import numpy as np, mipego_b200 as mb
from oracle import gp_oracle as O
rng=np.random.default_rng(0)
n,d=300,5
X=rng.uniform(-5,5,(n,d)); y=(X**2).sum(1); y=(y-y.mean())/y.std()
par=np.r_[np.full(d,0.05),0.8]
st=O.fit_state(X,y,par,"squared_exponential",True,0.0,"noisy",1e-6)
Xs=rng.uniform(-5,5,(3000,d))
rm,rs=O.predict(st,Xs)
for dev in (1,0,1):
    gp=mb.GaussianProcess(mean=mb.constant_trend(d,beta=None),corr="squared_exponential",theta0=[0.05]*d,thetaL=[1e-4]*d,thetaU=[10.0]*d,nugget=1e-6,device=dev)
    gp.fit_fixed(X,y,par)
    m,s=gp.predict(Xs,eval_MSE=True)
    llf,g,stt=gp.log_likelihood_batch(np.tile(par,(3,1)))
    print("device",dev,"mean err %.2e mse err %.2e llf %s"%(np.abs(m-rm).max(),np.abs(s-rs).max(),llf))
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/dist_fit_check.py 2>&1 | grep -v "^\*\*\*\|OMP_NUM\|^$" | tail -3
