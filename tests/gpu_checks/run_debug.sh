export PYTHONPATH=.
python - <<'PY'
import torch
from f5_tts_mlx_b200 import ops
from f5_tts_mlx_b200.dit import rope_table
dev="cuda"; g=torch.Generator().manual_seed(0)
def rnd(*s, scale=1.0): return (torch.randn(*s, generator=g)*scale).to(dev)
B,NF,D=2,937,1024; M=B*NF
a=rnd(M,D).bfloat16(); w=rnd(3*D,D,scale=D**-0.5).bfloat16(); bias=rnd(3*D); rope=rope_table(NF).to(dev)
for tile,var in ((192,2),(0,0)):
    out=torch.empty(M,3*D,device=dev,dtype=torch.bfloat16)
    ops.gemm(a,w,out,bias=bias,rope=rope,rope_cols=2*D,q_scale=0.125,q_cols=D,rows_per_batch=NF,num_batches=B,tile_n=tile,variant=var)
    torch.cuda.synchronize()
    ref=(a.float()@w.float().T+bias).view(B,NF,3*D//64,32,2)
    c,s=rope[None,:,None,:,0],rope[None,:,None,:,1]
    rot=torch.stack([ref[...,0]*c-ref[...,1]*s, ref[...,1]*c+ref[...,0]*s],dim=-1)
    ref2=ref.clone(); ref2[:,:,:2*D//64]=rot[:,:,:2*D//64]; ref2=ref2.reshape(M,3*D).clone(); ref2[:,:D]*=0.125
    print("qkv tile",tile,"var",var,"rel",((out.float()-ref2).norm()/ref2.norm()).item())
# 6-stage single-wave v1 (120 CTAs) with resid/gate
a=rnd(M,2048).bfloat16(); w=rnd(D,2048,scale=2048**-0.5).bfloat16(); bias=rnd(D); gate=rnd(6*D); x=rnd(M,D); x0=x.clone()
ops.gemm(a,w,x,bias=bias,resid=x,gate=gate[2*D:3*D],rows_per_batch=NF,num_batches=B)
torch.cuda.synchronize()
ref=x0+gate[2*D:3*D]*(a.float()@w.float().T+bias)
print("ff2 v1 6-stage rel",((x-ref).norm()/ref.norm()).item())
PY
python tests/gpu_checks/check_insitu.py 2>&1 | tail -12
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench5.json 2> gpurun_out/bench5.err; tail -2 gpurun_out/bench5.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench5.json').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'])
PY
