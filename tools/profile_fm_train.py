import os, sys, copy
sys.path.insert(0, os.getcwd())
import torch
from torch.profiler import profile, ProfilerActivity
from oracle.unet_module import UNet
from ratio_guided_multimodal_fm_b200 import CFMSchedule, synthetic as S
from ratio_guided_multimodal_fm_b200.fm_train import fm_step
dev = torch.device("cuda", 0)
kind = sys.argv[1] if len(sys.argv) > 1 else "unet_svhn"
c, hw = (3, 32) if kind == "unet_svhn" else (1, 32)
m = UNet(kind); m.load_state_dict(S.state_dict(kind, 12), strict=True); m = m.to(dev).train()
opt = torch.optim.Adam(m.parameters(), lr=1e-4)
x_1 = torch.randn(128, c, hw, hw, device=dev); sched = CFMSchedule()
def step():
    opt.zero_grad(); loss = fm_step(m, x_1, sched); opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))
