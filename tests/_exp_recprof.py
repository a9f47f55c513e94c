import sys, torch
sys.path.insert(0, '.')
import bench
from multi_view_stereonet_b200 import MultiViewStereoNet, synthetic
sd,_ = bench.load_state(); net = MultiViewStereoNet(); net.load_state_dict(sd); net=net.cuda().eval()
inp = synthetic.to_device(synthetic.make_inputs(512,640,1,1),"cuda")
net.set_option("recurrence_profile", 1)
with torch.no_grad():
    for _ in range(3): net(*inp,64,True,[True]*5)
torch.cuda.synchronize()
prof = net.get_stage("recurrence_profile", torch.int64).view(16,12).cpu()
names = ["W stage","MMA0","E0","barA","S1","MMA1","E1","barC","S2","MMA2","E2","barE"]
for r in (0,5,10):
    print("rank", r, " ".join(f"{n}={prof[r,i].item()/63:.0f}" for i,n in enumerate(names)), " total/step", prof[r].sum().item()/63)
