"""Error of the REFERENCE ALGORITHM itself under torch.autocast (CPU, bf16 / fp16) against its fp32 run, on the golden
fixtures: the yardstick for what a 16-bit-operand implementation can reach.  python tools/autocast_reference_error.py"""
import sys, torch, numpy as np
import os; ROOT=os.path.abspath(os.path.join(os.path.dirname(__file__),'..')); sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,'tests'))
from conftest import load_golden, T
from oracle import rgfm_oracle as O
from ratio_guided_multimodal_fm_b200 import synthetic as S
torch.set_num_threads(8)
for name in ["sampler_pair28_small","sampler_mnist_svhn_small","sampler_pair28_rulsif"]:
    g=load_golden(name); pair=str(g["pair"]); kx,ky,kr,_,_=O.PAIR_KINDS[pair]
    sx,sy,sr=(int(s) for s in g["seeds"])
    args=(S.state_dict(kx,sx),S.state_dict(ky,sy),S.state_dict(kr,sr),pair,str(g["method"]),float(g["gamma"]),int(g["steps"]),T(g["x0"]),T(g["y0"]),T(g["mc_x0"]),T(g["mc_y0"]),str(g["loss_type"]))
    for dt in (torch.bfloat16, torch.float16):
        try:
            with torch.autocast("cpu",dtype=dt):
                x1,y1=O.sample_pair(*args)
            print(name, dt, (x1.float()-T(g["x1"])).abs().max().item(), (y1.float()-T(g["y1"])).abs().max().item(), flush=True)
        except Exception as e:
            print(name, dt, "ERR", str(e)[:100])
