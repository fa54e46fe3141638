#!/usr/bin/env python
"""bench.py -- objects/s of the per-object pose-generation hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # B200 arm (libgenpose_b200.so)
    python bench.py --impl reference --gpus N ...             # reference's CPU path on the host cores

A "step" is one pass of the full path (PointNet++ encoder x2 -> RK45 ScoreNet sampling, 50
hypotheses/object -> EnergyNet scoring -> aggregation -> ScaleNet) over one batch of synthetic
objects.  Workload = BASELINE.json configs[1]: 64 objects x 50 hypotheses per GPU, T0 = 0.55
(scripts/eval_single.sh), random-init weights, synthetic clouds.  Multi-GPU: objects are sharded
by rank, no collective on the data path, weak scaling (64 objects per GPU).

`value`   : objects/s with the clouds already resident in HBM (CUDA events, max over ranks).
`e2e`     : objects/s through the same public call with HOST (pinned) clouds: H2D copy of the clouds
            and D2H read of the poses + lengths inside the timed region.
`roofline`: the dominant kernel of this library (the fused ScoreNet/RK45 integrator), algorithmic FLOPs
            (hoisted figure) / CUDA-event duration, against MEASURED_PEAKS.json.
`cpu_baseline`: the oracle port (oracle/pose_oracle.py: the reference's torch-CPU + scipy path,
            restated) timed on this box's host cores on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "objects_per_sec_50hyp"
UNIT = "objects/s"
OBJECTS_PER_GPU = 64
REPEAT = 50
T0 = 0.55
NUM_POINTS = 1024
L2_FLUSH_BYTES = 256 << 20

# dram__bytes_read.sum + dram__bytes_write.sum of the integrator kernel, `ncu --set full` captures in profiles/README.md
NCU_DRAM_BYTES_PER_LAUNCH = {"fp32_ffma": 2270464 + 22784, "bf16": 1686784 + 58368, "fp32": 2249472 + 72960}

# algorithmic work of the ScoreNet RHS after hoisting, in FLOP
ROW_EVAL_FLOP = 2 * 266752
STAGE_SHARED_FLOP = 2 * 114688
OBJECT_ONCE_FLOP = 2 * 786432


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", type=str, default="b200", choices=["b200", "reference"])
    ap.add_argument("--objects", type=int, default=OBJECTS_PER_GPU, help="objects per GPU per step")
    ap.add_argument("--mlp_mode", type=str, default="fp32", choices=["fp32", "fp32_ffma", "bf16"])
    ap.add_argument("--cpu_sample_objects", type=int, default=8)
    ap.add_argument("--no_cpu_baseline", action="store_true")
    ap.add_argument("--single_mode", action="store_true", help="skip the measurement of the other mlp_mode")
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"],
                    bf16_tflops_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


_SAMPLER_CODE = r"""
import sys, time
idx = int(sys.argv[1])
try:
    import pynvml as n
    n.nvmlInit()
    h = n.nvmlDeviceGetHandleByIndex(idx)
    get = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
    mx = n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM)
    print("ready nvml", flush=True)
    while True:
        print(time.time(), n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM), mx, int(get(h)), flush=True)
        time.sleep(0.004)
except Exception as e:
    import subprocess
    print("ready nvidia-smi", flush=True)
    q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    while True:
        t = time.time()
        out = subprocess.run(["nvidia-smi", "-i", str(idx), "--query-gpu=" + q, "--format=csv,noheader,nounits"], capture_output=True, text=True).stdout
        p = [x.strip() for x in out.strip().split(",")]
        if len(p) >= 6:
            bits = sum(b for b, v in zip((0x8, 0x40, 0x20, 0x4), p[2:6]) if v.lower().startswith("active"))
            print(t, p[0], p[1], bits, flush=True)
"""


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region by a helper process (NVML every ~4 ms, or
    `nvidia-smi` when NVML cannot be initialised): a thread in this process would have to fight the step loop for
    the GIL.  The helper is started at construction (before the warm-up); only its samples with a wall-clock stamp
    inside [start(), stop()] are kept."""
    BITS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index):
        phys = index
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:  # torch's index counts visible devices
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if index < len(ids) and ids[index].isdigit():
                phys = int(ids[index])
        self.proc = subprocess.Popen([sys.executable, "-c", _SAMPLER_CODE, str(phys)], stdout=subprocess.PIPE, text=True)
        self.source = (self.proc.stdout.readline().split() + ["?", "?"])[1]
        self.t0 = self.t1 = None

    def start(self):
        self.t0 = time.time()

    def stop(self):
        self.t1 = time.time()
        time.sleep(0.02)
        self.proc.terminate()
        try:
            out = self.proc.communicate(timeout=5)[0]
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], 0.0, set()
        for line in out.splitlines():
            p = line.split()
            if len(p) != 4:
                continue
            try:
                t, clk, mxc, bits = float(p[0]), float(p[1]), float(p[2]), int(p[3])
            except ValueError:
                continue
            if not (self.t0 <= t <= self.t1):
                continue
            sm.append(clk)
            mx = max(mx, mxc)
            reasons.update(name for name, b in self.BITS if bits & b)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm), "source": self.source}


# --------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores
# --------------------------------------------------------------------------------------------------
def cpu_reference_step(state, n_objects):
    """One bounded sample of the workload on the CPU: n_objects x 50 hypotheses through the oracle
    port of the reference path (encoder x2, scipy RK45 sampler, energy, aggregation with sklearn
    DBSCAN, ScaleNet).  Returns seconds."""
    import torch
    from oracle import pose_oracle as po
    pts, center = state["pts"][:n_objects], state["center"][:n_objects]
    torch.manual_seed(1)
    noise = po.ve_prior((n_objects * REPEAT, 9), T=T0)
    t0 = time.perf_counter()
    po.full_pipeline(state["score_sd"], state["energy_sd"], state["scale_sd"], pts, center, noise,
                     repeat_num=REPEAT, T0=T0, integrator="scipy")
    return time.perf_counter() - t0


def cpu_state(n_objects):
    import torch
    from genpose2_b200 import synthetic
    torch.set_num_threads(os.cpu_count() or 1)
    pts, center = synthetic.make_point_clouds(max(n_objects, 1), NUM_POINTS, seed=0)
    return dict(pts=pts, center=center, score_sd=synthetic.random_gfobjectpose_state_dict(100),
                energy_sd=synthetic.random_gfobjectpose_state_dict(200),
                scale_sd=synthetic.random_scalenet_state_dict(300))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # rank 0 alone runs the CPU reference; other ranks exit without work
    import torch
    n = args.cpu_sample_objects
    st = cpu_state(n)
    for _ in range(min(args.warmup, 1)):
        cpu_reference_step(st, n)
    times = [cpu_reference_step(st, n) for _ in range(args.steps)]
    total = sum(times)
    value = n * len(times) / total
    cores = torch.get_num_threads()
    sample = f"{n} objects x {REPEAT} hypotheses per step (full path incl. both encoders), {len(times)} steps"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"C2: {OBJECTS_PER_GPU} objects x {REPEAT} hypotheses, full path, T0={T0}",
                   "sampled_objects_per_step": n},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    from genpose2_b200 import _lib, samplers, synthetic
    from genpose2_b200.pipeline import PosePipeline

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (B200 arm) needs a GPU; there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    B = args.objects
    # every rank owns its own contiguous slice of the global object list (weak scaling)
    pts_all, center_all = synthetic.make_point_clouds(B * world, NUM_POINTS, seed=0)
    pts_h = pts_all[rank * B:(rank + 1) * B].contiguous().pin_memory()
    center_h = center_all[rank * B:(rank + 1) * B].contiguous().pin_memory()
    pts_d, center_d = pts_h.to(dev), center_h.to(dev)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    out_pose_h = torch.empty((B, 4, 4), dtype=torch.float32).pin_memory()
    out_len_h = torch.empty((B, 3), dtype=torch.float32).pin_memory()
    peaks = measured_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, sampler=None):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        if sampler:
            sampler.start()
        for s, e in evs:
            flush.zero_()  # L2 flush between timed iterations (inputs are far smaller than the 126 MB L2)
            s.record()
            fn()
            e.record()
        torch.cuda.synchronize()
        clocks = sampler.stop() if sampler else None
        barrier()
        total_ms = sum(s.elapsed_time(e) for s, e in evs)
        if world > 1:
            t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms = float(t.item())
        return total_ms, clocks

    def measure(mlp_mode, with_e2e, with_clocks):
        pipe = PosePipeline(device=f"cuda:{local_rank}", mlp_mode=mlp_mode).load_synthetic_weights((100, 200, 300))

        def step_resident():
            return pipe({"pts": pts_d, "pts_center": center_d}, repeat_num=REPEAT, T0=T0)

        def step_e2e():
            p = pts_h.to(dev, non_blocking=True)
            c = center_h.to(dev, non_blocking=True)
            pose, length = pipe({"pts": p, "pts_center": c}, repeat_num=REPEAT, T0=T0)
            out_pose_h.copy_(pose, non_blocking=True)
            out_len_h.copy_(length, non_blocking=True)
            return pose, length

        torch.manual_seed(1234 + rank)
        sampler = ClockSampler(local_rank) if with_clocks else None   # helper process: up and sampling before the timed region
        for _ in range(max(args.warmup, 3)):
            step_resident()
        torch.cuda.synchronize()
        _lib.reset_launch_count()
        total_ms, clocks = timed(step_resident, args.steps, sampler)
        res = {"launches": _lib.launch_count(), "total_ms": total_ms, "clocks": clocks,
               "value": world * B * args.steps / (total_ms * 1e-3)}
        if with_e2e:
            for _ in range(2):
                step_e2e()
            e2e_ms, _ = timed(step_e2e, args.steps)
            res["e2e_ms"] = e2e_ms
            res["e2e_value"] = world * B * args.steps / (e2e_ms * 1e-3)

        # ---- roofline of the dominant kernel of this library: the fused ScoreNet / RK45 integrator ----
        score_net = pipe.score_agent.net
        feat = score_net(dict(pts=pts_d, pts_center=center_d), mode="pts_feature")
        N = B * REPEAT
        sdata = {"pts": torch.empty(N, 0), "pts_center": center_d.unsqueeze(1).expand(B, REPEAT, 3).reshape(N, 3).contiguous(),
                 "_gp_pts_feat_obj": feat, "_gp_rows_per_object": REPEAT}
        torch.manual_seed(99)
        noise = score_net.prior_fn((N, 9), T=T0)
        prior = lambda shape, T=1.0: noise

        def sampler_only():
            return samplers.cond_ode_sampler(score_net, sdata, prior, score_net.sde_fn, device=dev, eps=1e-5, T=T0,
                                             pose_mode="rot_matrix", return_trajectory=False)

        for _ in range(3):
            sampler_only()
        torch.cuda.synchronize()
        k_ms = []
        for _ in range(max(5, min(args.steps, 20))):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            sampler_only()  # projection kernel (tiny) + the persistent integrator
            e.record()
            torch.cuda.synchronize()
            k_ms.append(s.elapsed_time(e))
        st = samplers.ode_stats()
        nfev_total = st["nfev"] + 1  # + the denoise evaluation
        flop = N * nfev_total * ROW_EVAL_FLOP + nfev_total * STAGE_SHARED_FLOP + B * OBJECT_ONCE_FLOP
        k_med = sorted(k_ms)[len(k_ms) // 2]
        achieved = flop / (k_med * 1e-3) / 1e12
        peak = peaks["bf16_tflops_sustained"]
        ffma_peak = 148 * 128 * 2 * 1.965e9 / 1e12  # 148 SMs x 128 FP32 lanes x 2 FLOP x max SM clock
        res["roofline"] = {
            "bound": "tensor", "kernel": "ode_rk45_kernel<%s> (fused ScoreNet RHS + Dormand-Prince controller)"
                                         % {"bf16": "TcEval<1>: tcgen05 bf16", "fp32": "TcEval<3>: tcgen05 split-bf16 x3",
                                            "fp32_ffma": "SimtEval: FFMA fp32"}[mlp_mode],
            "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            # dram__bytes_read.sum + dram__bytes_write.sum of this kernel, one `ncu --set full` capture of the same
            # workload (profiles/README.md, round 1): weights and state are L2-resident, the kernel is not HBM bound
            "traffic": NCU_DRAM_BYTES_PER_LAUNCH[mlp_mode],
            "peak_source": f"{peaks['source']} bf16 dense, sustained",
            "mlp_mode": mlp_mode, "kernel_ms": k_med, "nfev": nfev_total, "accepted": st["accepted"],
            "rejected": st["rejected"], "algorithmic_flop_per_launch": flop,
            "hyp_evals_per_s": N * nfev_total / (k_med * 1e-3),
            "note": {"fp32_ffma": "MLPs on FP32 FFMA (no tensor cores): also quoted against the FP32 FFMA peak",
                     "fp32": "split-bf16 x3 on tcgen05 (3 MMAs per product, fp32-class accuracy); achieved counts the "
                             "algorithmic FLOPs once, so the tensor pipe does 3x that",
                     "bf16": "bf16 operands on tcgen05, fp32 accumulation in TMEM"}[mlp_mode],
            "ffma_peak_tflops": ffma_peak, "frac_of_ffma_peak": achieved / ffma_peak,
        }
        return res

    main_res = measure(args.mlp_mode, True, True)
    other_mode = "bf16" if args.mlp_mode != "bf16" else "fp32"
    other_res = measure(other_mode, False, False) if not args.single_mode else None
    value, total_ms, clocks, launches = main_res["value"], main_res["total_ms"], main_res["clocks"], main_res["launches"]
    e2e_value, e2e_ms, roofline = main_res["e2e_value"], main_res["e2e_ms"], main_res["roofline"]
    h2d = pts_h.numel() * 4 + center_h.numel() * 4
    d2h = out_pose_h.numel() * 4 + out_len_h.numel() * 4

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n = args.cpu_sample_objects
        stt = cpu_state(n)
        cpu_reference_step(stt, min(n, 2))  # warm-up
        reps = [cpu_reference_step(stt, n) for _ in range(3)]
        med = sorted(reps)[1]
        cpu_baseline = {"value": n / med, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                        "sample": f"{n} objects x {REPEAT} hypotheses, full path incl. both encoders, median of 3 "
                                  f"(oracle port: torch-CPU nets + scipy RK45 + sklearn DBSCAN)",
                        "os_cpu_count": os.cpu_count()}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.mlp_mode == "bf16" else ("f32 (split-bf16 x3 tensor-core products, fp32 accumulation)" if args.mlp_mode == "fp32" else "f32"), "data": "synthetic",
            "config": {"workload": f"C2: {B} objects x {REPEAT} hypotheses per GPU, full path "
                                   f"(encoder x2 + RK45 ScoreNet sampling + EnergyNet + aggregation + ScaleNet), "
                                   f"T0={T0}, rtol=atol=1e-5, {NUM_POINTS} pts/object, random-init weights",
                       "objects_per_gpu": B, "hypotheses": REPEAT, "T0": T0, "l2": "flushed between timed steps "
                       f"({L2_FLUSH_BYTES >> 20} MiB memset)", "parallelism": f"object-sharded x{world}, no collective on the data path",
                       "streams": "the energy encoder runs on a second stream beside the cooperative sampler launch"},
            "clocks": clocks, "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                                      "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu_baseline,
        }
        if other_res is not None:
            line["other_mode"] = {"mlp_mode": other_mode, "value": other_res["value"], "unit": UNIT,
                                  "ms_per_step": other_res["total_ms"] / args.steps, "roofline": other_res["roofline"]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
