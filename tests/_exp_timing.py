import sys, time, torch
sys.path.insert(0, '.')
import bench
from multi_view_stereonet_b200 import MultiViewStereoNet, synthetic
sd,_ = bench.load_state(); net = MultiViewStereoNet(); net.load_state_dict(sd); net=net.cuda().eval()
inp = synthetic.to_device(synthetic.make_inputs(512,640,1,1),"cuda")
flags=(64,True,[True]*5)
flush = torch.empty(256<<20, dtype=torch.uint8, device="cuda")
def run(label, do_flush, sync_each, steps=20):
    with torch.no_grad():
        for _ in range(3): net(*inp,*flags)
        torch.cuda.synchronize()
        st=[torch.cuda.Event(enable_timing=True) for _ in range(steps)]; en=[torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        t0=time.perf_counter()
        cpu=0
        for i in range(steps):
            if do_flush: flush.zero_()
            st[i].record()
            c0=time.perf_counter(); net(*inp,*flags); cpu+=time.perf_counter()-c0
            en[i].record()
            if sync_each: torch.cuda.synchronize()
        torch.cuda.synchronize()
        wall=time.perf_counter()-t0
    ms=[a.elapsed_time(b) for a,b in zip(st,en)]
    print(f"{label}: event mean {sum(ms)/steps:.3f} ms (min {min(ms):.3f} max {max(ms):.3f}), wall/step {1e3*wall/steps:.3f} ms, cpu launch time/step {1e3*cpu/steps:.3f} ms", flush=True)
run("noflush async", False, False)
run("noflush sync ", False, True)
run("flush async  ", True, False)
run("flush sync   ", True, True)
net.set_option("tensor_cores", 0)
run("v0 noflush async", False, False)
