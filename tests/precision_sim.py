"""TEST INFRASTRUCTURE — where does the 16-bit tower's error come from?  

A CPU emulation (torch fp32) of the tower's roundings on the flagship 20x256 synthetic net, 16 positions: operands are
rounded to fp16 / bf16 exactly where the tensor-core path rounds them (weights once, activations at every layer input,
the residual stream kept as a hi+lo pair), accumulation and epilogue in fp32, heads from the unrounded last layer.
Rounding the weights only, or the activations only, shows that for bf16 each side alone already exceeds the 5e-3 value
tolerance — which is why compensating one operand (2 MMAs per product) cannot rescue bf16, and 3 MMAs are too slow.
    python tests/precision_sim.py [seed]      (about a minute on 8 cores)"""
import sys, os, tempfile, time
sys.path.insert(0,'/root/repo')
import numpy as np, torch, torch.nn.functional as F
from unrealgo_b200 import tfformat
import oracle
from oracle import tf_reader
torch.set_num_threads(8)
d=tempfile.mkdtemp()
seed=int(sys.argv[1]) if len(sys.argv)>1 else 20260119
meta,prefix=tfformat.write_synthetic_checkpoint(d,20,256,seed=seed)
V={k:torch.from_numpy(v) for k,v in tf_reader.read_bundle(prefix).items()}
packed,planes=oracle.synthetic_positions(16,seed=20260119)
x0=torch.from_numpy(oracle.transform_feature(planes,0,True)).float().permute(0,3,1,2)
def fold(vs):
    w=V[f"{vs}/conv/kernel"]; g=V[f"{vs}/bn/gamma"]; b=V[f"{vs}/bn/beta"]; m=V[f"{vs}/bn/moving_mean"]; v=V[f"{vs}/bn/moving_variance"]
    s=g/torch.sqrt(v+1e-5)
    return (w*s).permute(3,2,0,1).contiguous(), b-m*s
def q(t,mode):
    if mode=="fp32": return t
    if mode=="bf16": return t.to(torch.bfloat16).float()
    if mode=="fp16": return t.to(torch.float16).float()
def run(wmode, amode, split=True):
    def conv(x,vs,k=3):
        w,b=fold(vs)
        return F.conv2d(q(x,amode), q(w,wmode), padding=k//2)+b.view(1,-1,1,1)
    s=torch.relu(conv(x0,"resnet/init_conv"))
    def stream(t):  # residual stream storage: hi+lo pair (split) or single
        if amode=="fp32": return t
        hi=q(t,amode); 
        return hi+q(t-hi,amode) if split else hi
    s=stream(s)
    for k in range(20):
        t=torch.relu(conv(s,f"resnet/unit_res_{k}/sub1/conv1"))
        y=conv(t,f"resnet/unit_res_{k}/sub2/conv2")+s
        s=stream(torch.relu(y))
    # heads in fp32 from fp32 accumulators (fused)
    def head(vs): 
        w,b=fold(vs); return torch.relu(F.conv2d(s,w)+b.view(1,-1,1,1))
    p=head("resnet/policy_head/conv1x1").permute(0,2,3,1).reshape(x0.shape[0],-1)
    p=torch.softmax(p@V["resnet/policy_head/dense/kernel"]+V["resnet/policy_head/dense/bias"],-1)
    v=head("resnet/value_head/conv1x1").permute(0,2,3,1).reshape(x0.shape[0],-1)
    v=torch.relu(v@V["resnet/value_head/dense1/kernel"]+V["resnet/value_head/dense1/bias"])
    v=torch.tanh(v@V["resnet/value_head/dense2/kernel"]+V["resnet/value_head/dense2/bias"]).reshape(-1)
    return p,v
with torch.no_grad():
    rp,rv=run("fp32","fp32")
    for wm,am in [("fp16","fp16"),("bf16","bf16"),("bf16","fp32"),("fp32","bf16"),("fp16","bf16"),("bf16","fp16")]:
        p,v=run(wm,am)
        print(f"seed {seed} weights {wm:5s} activations {am:5s}: max|dp| {float((p-rp).abs().max()):.2e}  max|dv| {float((v-rv).abs().max()):.2e}")
