import os, sys, torch
sys.path.insert(0, os.getcwd())
from bench import NET_KW, H, W
from image_restoration_b200 import GFPGANv1OCR
from image_restoration_b200.backward import unet_forward
from torch.profiler import profile, ProfilerActivity
B = 64
torch.manual_seed(0)
net = GFPGANv1OCR(**NET_KW)
names = ('conv_body_first', 'conv_body_down', 'final_conv', 'final_linear', 'conv_body_up', 'condition_scale', 'condition_shift')
sd = {k: v.detach().clone().cuda().requires_grad_() for k, v in net.state_dict().items() if k.split('.')[0] in names}
x = torch.rand(B, 3, H, W, device='cuda') * 2 - 1
def step():
    style, conds = unet_forward(sd, x, different_w=True, num_style_feat=NET_KW['num_style_feat'])
    torch.autograd.backward([style] + conds, [torch.ones_like(t) for t in [style] + conds])
for _ in range(2): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=22, max_name_column_width=60))
