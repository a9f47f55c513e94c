#!/usr/bin/env python
"""Benchmark of the MultiViewStereoNet depth-inference hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch-per-gpu B]

Metric (BASELINE.json): depthmaps/sec at 512x640, 2-view (1 comparison view), 64 idepth
hypotheses.  One "step" is one MultiViewStereoNet.forward over one batch of synthetic image groups
(BASELINE cfg2: batch 1 per GPU).  For N > 1 the driver launches this file under torchrun; every
rank runs the same per-GPU workload on its own seeded items (weak scaling, no data-path collective)
and rank 0 prints ONE JSON line.

--impl reference times the reference's CPU implementation of the path: the reference is pure
PyTorch and cannot travel to the GPU box, so this is the oracle port (oracle/mvsnet_oracle.py, pinned
to the reference's outputs by tests/test_oracle.py) on all host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

# Rank 0 must print exactly one JSON line on stdout: NCCL's version banner (NCCL_DEBUG=VERSION) goes there too.
if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"

import torch  # noqa: E402

ROWS, COLS, VIEWS, HYPS = 512, 640, 1, 64       # BASELINE cfg2 / cfg4 per-item shape
METRIC = "depthmaps/sec at 512x640, 2-view, 64 hyp"
UNIT = "depthmaps/s"
E2E_BLOCKS = 5    # the end-to-end leg times this many blocks of K steps and reports the median block


_REAL_STDOUT = None


def reserve_stdout():
    """Keeps the process's stdout for the one JSON line: everything else that writes to fd 1 (NCCL's banner, library
    chatter) is sent to stderr."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def load_state():
    """Pretrained GTA-SfM weights if the fixture is present, else seeded random weights of the same
    architecture (timing does not depend on the values)."""
    from multi_view_stereonet_b200 import weights
    path = os.path.join(REPO, "tests", "golden", "gta_sfm_150epochs_state.npz")
    if os.path.exists(path):
        sd = weights.load_state_npz(path)
        fe = "left_feature_extractor."
        for k in [k for k in sd if k.startswith(fe)]:
            sd["right_feature_extractor.feature_extractor." + k[len(fe):]] = sd[k]
        return sd, "pretrained gta_sfm_150epochs (reference fixture)"
    return weights.seeded_random_state(0), "seeded random"


def algorithmic_work(rows, cols, views, hyps):
    """MACs and layerwise-compulsory fp32 bytes per depthmap."""
    h, w = [rows], [cols]
    for _ in range(4):
        h.append((h[-1] + 1) // 2)
        w.append((w[-1] + 1) // 2)
    P = [a * b for a, b in zip(h, w)]
    V, D = views, hyps
    mac = (1 + V) * (P[1] * 3 * 32 * 25 + (P[2] + P[3] + P[4]) * 32 * 32 * 25 + 7 * P[4] * 32 * 32 * 9)
    mac += V * (D - 1) * P[4] * (35 * 32 * 9 + 2 * 32 * 32 * 9)
    mac += V * D * P[4] * (4 * 32 * 32 * 27 + 32 * 27)
    byt = (1 + V) * 4 * ((3 * P[0] + 32 * P[1]) + 32 * (P[1] + P[2]) + 32 * (P[2] + P[3]) + 32 * (P[3] + P[4]) + 7 * 64 * P[4])
    byt += V * 4 * (6 * P[0] + 3 * P[4] + 3 * D * P[4])
    byt += V * (D - 1) * 4 * P[4] * (64 + 67 + 64 + 64)
    byt += V * 4 * (32 * P[4] + 64 * D * P[4])
    byt += V * 4 * D * P[4] * (4 * 64 + 33)
    byt += V * 4 * (D * P[4] + P[4])
    for lvl in range(5):
        cin = 4 if lvl == 0 else 36
        mult = V if lvl == 4 else 1
        mac += mult * P[lvl] * (cin * 32 * 9 + 6 * 32 * 32 * 9 + 32 * 9)
        byt += mult * 4 * P[lvl] * ((cin + 32) + 6 * 64 + 33)
    return mac, byt, P


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons while the timed region runs: NVML in-process (the queries behind
    `nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.*`), every 10 ms.  Spawning nvidia-smi
    itself next to the timed loop stalled single steps by several ms on these boxes; it is only the fallback
    when NVML cannot be loaded."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []          # (sm_mhz, sm_max_mhz, [active reasons])
        self.stop_flag = threading.Event()
        self.proc = None
        self.source = None
        self.nvml = None
        self.handle = None
        try:
            import pynvml
            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(index).uuid)
            if not uuid.startswith("GPU-"):
                uuid = "GPU-" + uuid
            try:
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(uuid)
            self.nvml = pynvml
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.source = "nvml"
        except Exception as e:      # noqa: BLE001
            log("clock sampler: NVML unavailable (%s), falling back to nvidia-smi" % (e,))
            self.nvml = None
            self.source = "nvidia-smi"

    def _run_nvml(self):
        n = self.nvml
        bits = [(n.nvmlClocksEventReasonHwSlowdown, "hw_slowdown"),
                (n.nvmlClocksEventReasonHwThermalSlowdown, "hw_thermal_slowdown"),
                (n.nvmlClocksEventReasonSwThermalSlowdown, "sw_thermal_slowdown"),
                (n.nvmlClocksEventReasonSwPowerCap, "sw_power_cap")]
        while not self.stop_flag.is_set():
            try:
                sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                r = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                self.samples.append((sm, self.sm_max, [name for bit, name in bits if r & bit]))
            except Exception:       # noqa: BLE001
                pass
            self.stop_flag.wait(0.01)

    def _run_smi(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag.is_set():
                    break
                parts = [p.strip() for p in line.split(",")]
                if len(parts) >= 6:
                    try:
                        self.samples.append((float(parts[0]), float(parts[1]),
                                             [name for name, val in zip(self.NAMES, parts[2:6])
                                              if val.lower().startswith("active")]))
                    except ValueError:
                        continue
        except Exception:           # noqa: BLE001
            pass

    def run(self):
        if self.nvml is not None:
            self._run_nvml()
        else:
            self._run_smi()

    def mark(self):
        """Index of the next sample: lets the caller keep only the samples taken inside the timed region."""
        return len(self.samples)

    def stop(self):
        self.stop_flag.set()
        if self.proc is not None:
            self.proc.terminate()

    def shutdown(self):
        """Stops the thread, waits for it and releases NVML."""
        self.stop()
        if self.is_alive():
            self.join(timeout=2.0)
        if self.proc is not None:
            try:
                self.proc.wait(timeout=2.0)
            except Exception:       # noqa: BLE001
                self.proc.kill()
        if self.nvml is not None:
            try:
                self.nvml.nvmlShutdown()
            except Exception:       # noqa: BLE001
                pass
            self.nvml = None

    def summary(self, first=0, last=None):
        window = self.samples[first:last]
        sm = [s[0] for s in window]
        mx = max([s[1] for s in window], default=0.0)
        reasons = set()
        for s in window:
            reasons.update(s[2])
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm), "source": self.source}


def cpu_oracle_throughput(state, steps, warmup, budget_s=25.0):
    """The oracle port on all host cores; returns (depthmaps/s, ms per depthmap, cores, runs)."""
    from oracle import mvsnet_oracle as oracle
    from multi_view_stereonet_b200 import synthetic
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    inputs = synthetic.make_inputs(ROWS, COLS, VIEWS, 1)
    times = []
    with torch.no_grad():
        for _ in range(warmup):
            oracle.forward(state, *inputs, HYPS, True, (True,) * 5)
        t_begin = time.perf_counter()
        for _ in range(steps):
            t0 = time.perf_counter()
            oracle.forward(state, *inputs, HYPS, True, (True,) * 5)
            times.append(time.perf_counter() - t0)
            if time.perf_counter() - t_begin > budget_s:
                break
    ms = 1e3 * sum(times) / len(times)
    return 1e3 / ms, ms, cores, len(times)


# BASELINE.json's throughput configurations as one GPU sees them (per-GPU share of the global batch); `value` stays on
# cfg2 / batch 1, these are reported next to it in the same JSON line ("configs").
EXTRA_CONFIGS = [
    ("cfg4_share", dict(rows=512, cols=640, views=1, hyps=64, batch=8),
     "cfg4: 512x640, 1 comparison view, 64 hypotheses, batch 64 over 8 GPUs = 8 per GPU"),
    ("cfg3", dict(rows=512, cols=640, views=4, hyps=64, batch=8),
     "cfg3: 512x640, 4 comparison views, 64 hypotheses, batch 8 on one GPU"),
    ("cfg5_share", dict(rows=1024, cols=1280, views=4, hyps=128, batch=4),
     "cfg5: 1024x1280, 4 comparison views, 128 hypotheses, batch 32 over 8 GPUs = 4 per GPU"),
]


def run_extra_configs(net, dev, rank, world, flush, barrier, hbm_peak, tf32_peak_tflops, steps=5, warmup=4):
    """Times the per-GPU share of cfg4 / cfg3 / cfg5 device-resident (every rank runs its own seeded items, max over
    ranks) and one profiled forward each for the per-stage breakdown.  Returns {name: {...}} on every rank."""
    from multi_view_stereonet_b200 import sharding, synthetic
    out = {}
    for name, c, what in EXTRA_CONFIGS:
        B = c["batch"]
        first_item, _ = sharding.shard_range(world * B, rank, world)
        inputs = synthetic.to_device(synthetic.make_inputs(c["rows"], c["cols"], c["views"], B, first_item=first_item), dev)
        flags = (c["hyps"], True, [True] * 5)
        with torch.no_grad():
            # warm-up and timed steps run in ONE loop with the same event / output-lifetime pattern (a first timed step
            # that allocates differently from the warm-up stalled in cudaMalloc for ~30 ms once); the last `steps`
            # iterations are the timed ones
            res = None
            starts = [torch.cuda.Event(enable_timing=True) for _ in range(warmup + steps)]
            ends = [torch.cuda.Event(enable_timing=True) for _ in range(warmup + steps)]
            barrier()   # ranks start together; no host synchronisation between the warm-up and the timed steps (the
            #             first forward behind a synchronise ran 1-9 ms long in the two-lane configuration)
            for i in range(warmup + steps):
                flush.zero_()
                starts[i].record()
                res = net(*inputs, *flags)
                ends[i].record()
            barrier()
            launches = net.last_launch_count()
            step_ms = [s.elapsed_time(e) for s, e in zip(starts[warmup:], ends[warmup:])]
            # per-stage breakdown: one extra forward with events at the stage boundaries (not part of the timing above)
            net.set_option("stage_profile", 1)
            net(*inputs, *flags)
            torch.cuda.synchronize(dev)
            stages = net.last_stage_profile()
            net.set_option("stage_profile", 0)
            del res
        allt = sharding.gather_timings([sum(step_ms)], device=dev)
        _, worst_ms = sharding.aggregate_throughput(B, steps, allt[:, 0])
        mac, byt, _ = algorithmic_work(c["rows"], c["cols"], c["views"], c["hyps"])
        t_roof_us = max(byt / (hbm_peak * 1e9), 2 * mac / (tf32_peak_tflops * 1e12)) * 1e6
        ms_step = worst_ms / steps
        out[name] = {
            "workload": what, "batch_per_gpu": B, "global_batch": world * B, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_step, "depthmaps_per_s": world * B / (ms_step * 1e-3),
            "roofline_us_per_depthmap": t_roof_us, "frac_of_roofline": t_roof_us / (ms_step / B * 1e3),
            "gpu_launches_per_step": launches, "step_ms_rank0": [round(t, 3) for t in step_ms],
            "stage_us_rank0": {k: round(v, 1) for k, v in stages.items()},
        }
        del inputs
        torch.cuda.empty_cache()
    return out


def run_reference(args, rank):
    if rank != 0:
        return
    state, wsrc = load_state()
    dmps, ms, cores, runs = cpu_oracle_throughput(state, args.steps, max(1, min(args.warmup, 2)), budget_s=150.0)
    sample = f"{runs} full forwards of one 512x640 / 1 comparison view / 64 hypotheses image group (batch 1)"
    line = {
        "impl": "reference", "metric": METRIC, "value": dmps, "unit": UNIT, "n_gpus": args.gpus, "steps": runs,
        "warmup": max(1, min(args.warmup, 2)), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg2: 512x640, 1 comparison view, 64 idepth hypotheses, batch 1", "weights": wsrc,
                   "device": "host CPU (torch %s, %d threads)" % (torch.__version__, cores)},
        "cpu_baseline": {"value": dmps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": dmps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    from multi_view_stereonet_b200 import MultiViewStereoNet, sharding, synthetic

    assert torch.cuda.is_available(), "bench.py --impl ours needs a CUDA device (there is no CPU path)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    state, wsrc = load_state()
    net = MultiViewStereoNet()
    net.load_state_dict(state, strict=True)
    net = net.to(dev).eval()

    B = args.batch_per_gpu
    first_item, _ = sharding.shard_range(world * B, rank, world)
    cpu_inputs = synthetic.make_inputs(ROWS, COLS, VIEWS, B, first_item=first_item)
    inputs = synthetic.to_device(cpu_inputs, dev)
    flags = (HYPS, True, [True] * 5)
    flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)   # 256 MiB > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident throughput ("value") ----
    with torch.no_grad():
        net(*inputs, *flags)                   # builds the native handle and the workspace
        launches_per_step = net.last_launch_count()
        if not os.environ.get("BENCH_NO_PROBE"):
            net.probe_select("refine_conv32_l0")
        # nvidia-smi starts polling before the warm-up so that its start-up is not inside the timed region, and the
        # warm-up steps run back to back with the timed ones (same L2 flush between them)
        sampler = ClockSampler(local_rank)
        if not os.environ.get("BENCH_NO_SAMPLER"):
            sampler.start()
            time.sleep(float(os.environ.get("BENCH_SLEEP", "0.2")))
        out = None
        for _ in range(max(args.warmup, 3)):
            flush.zero_()
            out = net(*inputs, *flags)         # held like in the timed loop: the caching allocator reaches its steady
                                               # state (two live output sets) here, not in the second timed step
        if not os.environ.get("BENCH_NO_PROBE"):
            net.probe_read()                   # drop the warm-up launches from the kernel probe
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        barrier()
        mark0 = sampler.mark()
        wall0 = time.perf_counter()
        for i in range(args.steps):
            flush.zero_()                      # evict L2 between timed iterations (outside the event pair)
            starts[i].record()
            out = net(*inputs, *flags)
            ends[i].record()
        barrier()
        wall = time.perf_counter() - wall0
        mark1 = sampler.mark()
        step_ms = [s.elapsed_time(e) for s, e in zip(starts, ends)]
        probe_ms, probe_launches = net.probe_read()
        net.probe_select("none")
        time.sleep(0.2)
        sampler.stop()
        # the longest single kernel (latency bound: 63 dependent steps) -- probed in a short untimed pass
        rec_ms, rec_launches = 0.0, 0
        if not os.environ.get("BENCH_NO_PROBE"):
            net.probe_select("recurrence")
            for _ in range(5):
                net(*inputs, *flags)
            torch.cuda.synchronize(dev)
            rec_ms, rec_launches = net.probe_read()
            net.probe_select("none")
    total_ms = sum(step_ms)

    # ---- end to end through the host entry: pinned inputs -> H2D -> forward -> D2H ----
    pin = lambda t: t.pin_memory()
    host_inputs = ([pin(t) for t in cpu_inputs[0]], [pin(t) for t in cpu_inputs[1]], [pin(t) for t in cpu_inputs[2]],
                   [[pin(t) for t in p] for p in cpu_inputs[3]])
    host_out = {"left_idepthmap_pyr": [torch.empty((B, 1) + tuple(t.shape[-2:]), dtype=torch.float32).pin_memory()
                                       for t in cpu_inputs[0]]}
    net.set_host_outputs(host_out)
    e2e_steps = args.steps
    with torch.no_grad():
        # the host side of this path (pinned staging, driver queues, Python call overhead) keeps getting faster for
        # the first dozens of calls: warm it up longer than the device leg needs (untimed, ~40 ms)
        for _ in range(max(args.warmup, 20)):
            net(*host_inputs, *flags)
        barrier()
        # A block of K steps is ~30 ms of wall clock at this workload: one host hiccup on a fresh box moves it by 10-20 %
        # (a 565 next to 650-670 depthmaps/s in otherwise identical runs).  The block is therefore timed E2E_BLOCKS
        # times back to back and the MEDIAN block is reported; every block is listed in the JSON line.
        e2e_blocks_s = []
        for _ in range(E2E_BLOCKS):
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                net(*host_inputs, *flags)          # synchronous: returns after the D2H copy
            torch.cuda.synchronize(dev)
            e2e_blocks_s.append(time.perf_counter() - t0)
        e2e_s = sorted(e2e_blocks_s)[len(e2e_blocks_s) // 2]
    h2d, d2h = net.last_h2d_bytes, net.last_d2h_bytes
    # the same call with EVERYTHING the module returns downloaded (raw priors and the five dense mask volumes as well,
    # +28 MB at this workload): reported next to `e2e` as `e2e_all_outputs` (rank 0's own time, not part of `value`)
    shapes = [tuple(t.shape[-2:]) for t in cpu_inputs[0]]
    host_all = {"left_idepthmap_pyr": host_out["left_idepthmap_pyr"],
                "left_idepthmap_raw_pyr": [torch.empty((B, 1) + sh, dtype=torch.float32).pin_memory() for sh in shapes],
                "left_idepthmap_mask_pyr": [torch.empty((B, HYPS) + sh, dtype=torch.uint8).pin_memory() for sh in shapes]}
    net.set_host_outputs(host_all)
    with torch.no_grad():
        for _ in range(5):
            net(*host_inputs, *flags)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            net(*host_inputs, *flags)
        torch.cuda.synchronize(dev)
        e2e_all_s = time.perf_counter() - t0
    d2h_all = net.last_d2h_bytes
    net.set_host_outputs(None)

    # ---- the other BASELINE configurations (per-GPU share), device-resident, every rank ----
    peaks_all = {}
    if os.path.exists(os.path.join(REPO, "MEASURED_PEAKS.json")):
        peaks_all = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
    extra = None
    if not args.no_configs:
        extra = run_extra_configs(net, dev, rank, world, flush, barrier, float(peaks_all.get("hbm_gbs", 6650.0)),
                                  0.5 * float(peaks_all.get("bf16_tflops", 1590.0)))

    # ---- max over ranks (the path's only collective: one all_gather of per-rank timings) ----
    allt = sharding.gather_timings([total_ms, e2e_s * 1e3], device=dev)
    _, worst_ms = sharding.aggregate_throughput(B, args.steps, allt[:, 0])
    _, worst_e2e_ms = sharding.aggregate_throughput(B, e2e_steps, allt[:, 1])

    if rank == 0:
        mac, byt, P = algorithmic_work(ROWS, COLS, VIEWS, HYPS)
        peaks = {}
        ppath = os.path.join(REPO, "MEASURED_PEAKS.json")
        peak_src = "fallback (B200_PROFILING.md)"
        hbm_peak = 6650.0
        if os.path.exists(ppath):
            peaks = json.load(open(ppath))
            hbm_peak = float(peaks.get("hbm_gbs", hbm_peak))
            peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
        # dominant kernel: the 3x3 32->32 refiner convolution at level 0; algorithmic bytes per launch
        # = one fp32 read + one fp32 write of a (B, 512, 640, 32) activation
        k_bytes = B * P[0] * 64 * 4
        traffic = None
        tpath = os.path.join(REPO, "profiles", "r2_traffic.json")
        if os.path.exists(tpath) and B == 1:
            # dram__bytes_read.sum + dram__bytes_write.sum per launch of this kernel from the committed ncu --set full capture
            traffic = json.load(open(tpath)).get("traffic_bytes_per_launch")
        k_ms = probe_ms / max(probe_launches, 1)
        achieved = k_bytes / (k_ms * 1e-3) / 1e9 if k_ms > 0 else None
        value = world * B * args.steps / (worst_ms * 1e-3)
        e2e_value = world * B * e2e_steps / (worst_e2e_ms * 1e-3)
        t_roof_us = max(byt / (hbm_peak * 1e9), 2 * mac / (0.5 * float(peaks.get("bf16_tflops", 1590.0)) * 1e12)) * 1e6

        cpu = None
        if world == 1:      # the CPU baseline is timed on rank 0 at N = 1 only (the other ranks would sit in a
                            # NCCL barrier, spinning on their GPUs, for as long as it runs)
            dmps, ms, cores, runs = cpu_oracle_throughput(state, 20, 1, budget_s=15.0)
            cpu = {"value": dmps, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"{runs} full forwards of one cfg2 image group (oracle/mvsnet_oracle.py, torch CPU, "
                             f"{cores} threads), {ms:.0f} ms each"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": worst_ms / args.steps, "ms_per_step_median_rank0": statistics.median(step_ms),
            "step_ms_rank0": [round(x, 4) for x in step_ms],
            "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"cfg2: 512x640, 1 comparison view, 64 idepth hypotheses, batch {B} per GPU",
                       "global_batch": world * B, "parallelism": f"dp{world} (independent image groups, no data-path "
                       "collective; one all_gather of timings)", "weights": wsrc,
                       "l2": "256 MiB buffer written between timed steps (outside the event pairs)",
                       "timing": "CUDA events per step on the launch stream, max over ranks of the per-rank sum",
                       "wall_s_incl_flush": wall},
            "clocks": sampler.summary(mark0, max(mark1, mark0 + 1)),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": worst_e2e_ms / e2e_steps,
                    "estimator": f"median of {E2E_BLOCKS} back-to-back blocks of {e2e_steps} steps (wall clock), max over ranks",
                    "blocks_ms_per_step_rank0": [round(1e3 * b / e2e_steps, 4) for b in e2e_blocks_s],
                    "what": "MultiViewStereoNet.forward on pinned CPU tensors -> b200mvs_forward_host: H2D of the "
                            "image pyramids/K/T, full path incl. mask volumes on device, D2H of the 5-level idepth "
                            "pyramid"},
            "e2e_all_outputs": {"value": B * e2e_steps / e2e_all_s, "unit": UNIT, "ms_per_step": 1e3 * e2e_all_s / e2e_steps,
                                "d2h_bytes_per_step": d2h_all, "scope": "rank 0, one GPU's share",
                                "what": "the same call downloading idepth + raw prior pyramids and the five dense mask "
                                        "volumes (what forward() on CPU tensors returns by default)"},
            "gpu_launches": launches_per_step * args.steps,
            "gpu_launches_per_step": launches_per_step,
            "roofline": {"bound": "hbm", "kernel": "conv3x3_ws_kernel (refiner0 residual 3x3 32->32 convs, level 0)",
                         "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": (achieved / hbm_peak) if achieved else None, "traffic": traffic,
                         "peak_source": peak_src, "kernel_ms": k_ms, "kernel_launches": probe_launches,
                         "algorithmic_bytes_per_launch": k_bytes,
                         "kernel_share_of_step": probe_ms / total_ms if total_ms > 0 else None},
            "latency_kernel": {"kernel": "recurrence_kernel (persistent cluster, D-1 dependent steps)",
                               "kernel_ms": rec_ms / max(rec_launches, 1), "kernel_launches": rec_launches,
                               "us_per_step": 1e3 * rec_ms / max(rec_launches, 1) / max(B * (HYPS - 1), 1) * B,
                               "share_of_step": (rec_ms / max(rec_launches, 1)) / (total_ms / args.steps) if total_ms > 0 else None},
            "whole_path": {"algorithmic_gflop_per_depthmap": 2 * mac / 1e9, "algorithmic_mb_per_depthmap": byt / 1e6,
                           "roofline_us_per_depthmap": t_roof_us,
                           "frac_of_roofline": t_roof_us / (worst_ms / args.steps / B * 1e3)},
            "cpu_baseline": cpu,
            "configs": extra,
        }
        emit(line)
    # leave the device idle and every helper stopped before the process goes away
    sampler.shutdown()
    net.set_host_outputs(None)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize(dev)
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch-per-gpu", type=int, default=1)
    ap.add_argument("--no-configs", action="store_true",
                    help="skip the cfg3 / cfg4-share / cfg5-share legs (the `configs` block of the JSON line)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    reserve_stdout()
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world == 1 and args.gpus > 1:
        # Launched without torchrun: re-exec under it so that `python bench.py --gpus N` works too.
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29517"), __file__,
               "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup),
               "--batch-per-gpu", str(args.batch_per_gpu)] + (["--no-configs"] if args.no_configs else [])
        sys.exit(subprocess.call(cmd))
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
