import os, sys, cProfile, pstats, torch
sys.path.insert(0, os.getcwd())
from bench import NET_KW, H, W
from image_restoration_b200 import GFPGANv1OCR
from image_restoration_b200.backward import unet_forward
B = 8
torch.manual_seed(0)
net = GFPGANv1OCR(**NET_KW)
names = ('conv_body_first', 'conv_body_down', 'final_conv', 'final_linear', 'conv_body_up', 'condition_scale', 'condition_shift')
sd = {k: v.detach().clone().cuda().requires_grad_() for k, v in net.state_dict().items() if k.split('.')[0] in names}
x = torch.rand(B, 3, H, W, device='cuda') * 2 - 1
cots = []
def step():
    for v in sd.values(): v.grad = None
    s, c = unet_forward(sd, x, different_w=True, num_style_feat=NET_KW['num_style_feat'])
    if not cots: cots.extend(torch.randn_like(t) for t in [s] + c)
    torch.autograd.backward([s] + c, cots)
for _ in range(3): step()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(5): step()
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(22)
