import sys, torch
sys.path.insert(0, '.')
from oracle import rgfm_oracle as O
from ratio_guided_multimodal_fm_b200 import RatioTrainPlan, synthetic
DEV='cuda:0'
for kind, loss_type in (("ratio28","disc"),("ratio_mnist_svhn","rulsif")):
    sd = synthetic.state_dict(kind, 21 if kind=="ratio28" else 22)
    hw, cy, widths = (28,1,(512,256)) if kind=="ratio28" else (32,3,(512,512))
    gen = torch.Generator().manual_seed(31)
    params = {k: v.to(DEV).contiguous() for k,v in sd.items() if v.dtype.is_floating_point}
    plan = RatioTrainPlan(kind, params, max_batch=128, device=DEV)
    for B in (128, 16):
        x, y = torch.randn(B,1,hw,hw,generator=gen), torch.randn(B,cy,hw,hw,generator=gen)
        r = (torch.rand(B,generator=gen)<0.6).float()
        masks = [torch.bernoulli(torch.full((B,w),0.9),generator=gen)/0.9 for w in widths]
        oloss,_,og = O.ratio_train_grads(sd,kind,x,y,r,loss_type,0.1,masks)
        # float64 yardstick: the same oracle in double
        sd64 = {k:(v.double() if v.dtype.is_floating_point else v) for k,v in sd.items()}
        _,_,og64 = O.ratio_train_grads(sd64,kind,x.double(),y.double(),r,loss_type,0.1,[m.double() for m in masks])
        out8, g = plan.step(x.to(DEV),y.to(DEV),r.to(DEV),loss_type,0.1,[m.to(DEV) for m in masks])
        torch.cuda.synchronize()
        print(kind, B, 'loss', out8[0].item(), oloss.item())
        for k,ref in og.items():
            if 'conv' in k and 'weight' in k or 'score_net.0.weight' in k:
                r64 = og64[k]
                e_ours = ((g[k].cpu().double()-r64).norm()/r64.norm()).item()
                e_orc = ((ref.double()-r64).norm()/r64.norm()).item()
                print(f"  {k:34s} ours-vs-f64 {e_ours:.2e}   torch-f32-vs-f64 {e_orc:.2e}   ours-vs-torch-f32 {((g[k].cpu()-ref).norm()/ref.norm()).item():.2e}")
