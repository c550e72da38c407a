"""One eager (no CUDA graph) CFG denoise step at the bench workload, for `ncu` launch lists / kernel captures."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from animate3d_b200.pipeline import AnimateDiffMVI2VPipeline, get_camera
from animate3d_b200.scheduler import DDIMScheduler
from animate3d_b200.unet import MVUNetMotionModel
from animate3d_b200.unet_config import UNetConfig
from animate3d_b200.weights import random_state_dict

NV, NF, LAT = 4, 16, 32
cfg = UNetConfig()
model = MVUNetMotionModel(cfg)
model.use_cuda_graph = False
model.load_state_dict(random_state_dict(cfg, 0, "cuda"))
model._prepare()
model.drop_reference_weights()
sched = DDIMScheduler()
sched.set_timesteps(25)
pipe = AnimateDiffMVI2VPipeline(unet=model, scheduler=sched)
g = torch.Generator(device="cuda").manual_seed(0)
lat = torch.randn(NV, 4, NF, LAT, LAT, device="cuda", generator=g)
first = lat[:, :, :1].clone()
pe = torch.randn(2 * NV, 77, 768, device="cuda", generator=g)
ie = torch.randn(2 * NV, 1024, device="cuda", generator=g)
cam = get_camera(NV).cuda()
steps = int(os.environ.get("A3D_STEPS", "1"))
for i in range(steps):
    pipe.denoise_step(lat, 961, pe, torch.cat([cam, cam]), ie, first, 7.5, num_views=NV)
torch.cuda.synchronize()
print("one_step done; launches per forward:", model.launches_per_forward)
