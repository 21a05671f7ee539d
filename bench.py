#!/usr/bin/env python
"""Benchmark of the UniDepthV2.infer() hot path (see BASELINE.json / SURVEY.md section 8d).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" = one `infer` pass over one batch of synthetic uint8 RGB (ViT-L/14, 8 x 3x480x640 per GPU).
Rank 0 prints ONE JSON line.  `value` = images/s with inputs resident in HBM (whole job, max over
ranks); `e2e` = images/s through the public API with pinned-host input -> H2D -> infer -> D2H of
depth + intrinsics inside the timed region.  `--impl reference` times the reference algorithm's CPU
implementation (the torch-fp32 oracle port; /root/reference does not exist on the GPU box).
"""
from __future__ import annotations

import argparse
import copy
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

FLOPS_PER_IMAGE = 1648.56e9     # SURVEY.md section 8d / BASELINE.md section 2 (ViT-L/14 @ 490x644)
WORKLOAD = dict(model="UniDepthV2 ViT-L/14", batch_per_gpu=8, input="3x480x640 uint8", net_input="490x644")


def load_config():
    return json.load(open(os.path.join(ROOT, "tests", "golden", "config_v2_vitl14.json")))


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1446.4), d.get("bf16_tflops", 1695.9), d.get("hbm_gbs", 6555.5), "measured"
    return 1400.0, 1590.0, 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.lines = []
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu_index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], 0, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx = max(mx, float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def run_reference(args, rank, world):
    """CPU arm: the oracle port of the reference's infer on the host cores (bounded sample:
    one image of the workload per step)."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import unidepth_oracle as O
    from fixture import make_state_dict
    cfg = load_config()
    cores = min(32, os.cpu_count())   # torch CPU ops stop scaling (and regress) beyond a few dozen threads
    torch.set_num_threads(cores)
    sd = make_state_dict(cfg, 0)
    g = torch.Generator().manual_seed(0)
    rgb = torch.randint(0, 256, (1, 3, 480, 640), dtype=torch.uint8, generator=g)
    for _ in range(max(1, min(args.warmup, 2))):
        O.infer_v2(sd, copy.deepcopy(cfg), rgb)
    steps = max(1, args.steps)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.infer_v2(sd, copy.deepcopy(cfg), rgb)
    dt = time.perf_counter() - t0
    val = steps / dt
    line = {
        "impl": "reference", "metric": "images/sec UniDepthV2.infer ViT-L/14 480x640", "value": val, "unit": "images/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": 1000 * dt / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": "1 image per step (batch 1) of the 8-image batch"},
        "cpu_baseline": {"value": val, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": f"{steps} x batch-1 infer, torch fp32, {cores} threads of {os.cpu_count()} cores"},
        "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="images per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    warmup = max(3, args.warmup)

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from unidepth_b200 import UniDepthV2
    from unidepth_b200 import _cabi
    from unidepth_b200.synthetic import synthetic_state_dict
    from unidepth_b200.parallel import gather_outputs

    cfg = load_config()
    model = UniDepthV2(copy.deepcopy(cfg))
    model.load_state_dict(synthetic_state_dict(cfg, 0, device=dev), strict=True)   # same seed on every rank
    model = model.to(dev).eval()
    model.resolution_level = None
    B = args.batch
    g = torch.Generator().manual_seed(rank)
    rgb_host = torch.randint(0, 256, (B, 3, 480, 640), dtype=torch.uint8, generator=g).pin_memory()
    rgb_dev = rgb_host.to(dev)
    import warnings
    warnings.simplefilter("ignore")

    # N > 1: one all-gather of the packed outputs per step.  Default: synchronous NCCL gather (an NCCL kernel left
    # in flight under the next step steals SMs from the persistent GEMM CTAs and costs more than it hides:
    # 18.6 vs 18.0 ms/step at N=2).  UDB_GATHER=p2p: copy-engine pulls over NVLink peer memory, left in flight
    # under the next step (depth-1 pipeline); every gather is waited for inside the timed region (flush()).
    pending = {"dev": None, "e2e": None}
    pipelined = os.environ.get("UDB_GATHER", "nccl") == "p2p"

    def step_device():
        out = model.infer(rgb_dev)
        if world > 1 and pipelined:
            nxt = gather_outputs(out, world, async_op=True)
            if pending["dev"] is not None:
                pending["dev"].wait()
            pending["dev"] = nxt
        elif world > 1:
            out = gather_outputs(out, world)
        return out

    depth_host = torch.empty((B, 1, 480, 640), dtype=torch.float32).pin_memory()
    k_host = torch.empty((B, 3, 3), dtype=torch.float32).pin_memory()

    def _d2h(out):
        lo = rank * B if world > 1 else 0
        depth_host.copy_(out["depth"][lo:lo + B], non_blocking=True)
        k_host.copy_(out["intrinsics"][lo:lo + B], non_blocking=True)

    def step_e2e():
        x = rgb_host.to(dev, non_blocking=True)
        out = model.infer(x)
        if world > 1 and pipelined:
            nxt = gather_outputs(out, world, async_op=True)
            if pending["e2e"] is not None:
                _d2h(pending["e2e"].wait())
            pending["e2e"] = nxt
        else:
            if world > 1:
                out = gather_outputs(out, world)
            _d2h(out)
        return out

    def gather_kind():
        from unidepth_b200 import parallel
        if parallel._p2p_cache:
            return ("one packed all-gather of the per-rank outputs per step: copy-engine pulls over NVLink peer memory "
                    "(torch symmetric memory) on a side stream, left in flight under the next step's compute (depth-1 "
                    "pipeline); every gather completes inside the timed region")
        return "one packed NCCL all_gather_into_tensor of the per-rank outputs per step, synchronous, inside the timed region"

    def flush():
        if pending["dev"] is not None:
            pending["dev"].wait()
            pending["dev"] = None
        if pending["e2e"] is not None:
            _d2h(pending["e2e"].wait())
            pending["e2e"] = None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            fn()
        flush()
        e.record()
        barrier()
        ms = torch.tensor([s.elapsed_time(e)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    l0 = _cabi.launch_count()
    model.use_cuda_graph = False
    step_device()                                   # eager once: counts our launches per forward
    torch.cuda.synchronize()
    launches_per_step = _cabi.launch_count() - l0
    model.use_cuda_graph = True
    for _ in range(warmup):
        step_device()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms = timed(step_device, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)

    # per-kernel roofline for the dominant kernel (tcgen05 GEMM): instrumented eager pass
    roof = None
    if rank == 0:
        from unidepth_b200 import ops
        model.use_cuda_graph = False
        model.use_engine = False      # same kernels scheduled from Python so that each launch can be bracketed by events
        ops.PROFILE = []
        # keep the GPU busy while the host enqueues the whole eager pass (launches + event records),
        # so the events bracket back-to-back kernel executions, not host launch gaps
        torch.cuda._sleep(int(0.12 * 1.9e9))
        model.infer(rgb_dev)
        torch.cuda.synchronize()
        prof, ops.PROFILE = ops.PROFILE, None
        model.use_cuda_graph = True
        model.use_engine = True
        agg = {}
        for name, flops, s, e in prof:
            a = agg.setdefault(name, [0.0, 0.0, 0])
            a[0] += flops
            a[1] += s.elapsed_time(e)
            a[2] += 1
        sustained, burst, hbm, how = measured_peaks()
        tot_ms = sum(a[1] for a in agg.values())
        kern = {k: {"launches": a[2], "ms": round(a[1], 3), "tflops": round(a[0] / a[1] / 1e9, 1) if a[1] > 0 else None,
                    "share": round(a[1] / tot_ms, 3)} for k, a in agg.items()}
        gm = agg.get("gemm_f16_kernel", [0.0, 1.0, 1])   # ops.py labels both GEMM kernels with this key
        ach = gm[0] / gm[1] / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "gemm_traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get("bytes_per_launch_avg")   # ncu capture of the 4 encoder GEMM flavours
        roof = {"bound": "tensor", "kernel": "gemm_f16_kernel / gemm2_f16_kernel (linear + conv3x3 + convT launches)",
                "achieved": round(ach, 1), "peak": sustained, "unit": "TFLOP/s", "frac": round(ach / sustained, 4),
                "peak_source": f"{how} bf16_tflops_sustained (MEASURED_PEAKS.json)", "traffic": traffic,
                "launches": gm[2], "avg_launch_us": round(1000 * gm[1] / max(gm[2], 1), 2),
                "flops_per_launch_avg": round(gm[0] / max(gm[2], 1) / 1e9, 2),
                "step_tflops": round(B * FLOPS_PER_IMAGE / (ms / args.steps) / 1e9, 1),
                "step_frac": round(B * FLOPS_PER_IMAGE / (ms / args.steps) / 1e9 / sustained, 4),
                "kernels": kern}

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import unidepth_oracle as O
        cores = min(32, os.cpu_count())
        torch.set_num_threads(cores)
        sd_cpu = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        one = rgb_host[:1].clone()
        O.infer_v2(sd_cpu, copy.deepcopy(cfg), one)
        n = 3
        t0 = time.perf_counter()
        for _ in range(n):
            ref = O.infer_v2(sd_cpu, copy.deepcopy(cfg), one)
        dt = time.perf_counter() - t0
        got = model.infer(rgb_dev[:1])
        d, dr = got["depth"].cpu(), ref["depth"]
        arel = ((d - dr).abs() / dr).mean().item()
        cpu_base = {"value": n / dt, "unit": "images/s", "cores": cores, "kind": "port",
                    "sample": f"{n} x batch-1 infer of the same weights/input (torch fp32 oracle, {cores} threads of {os.cpu_count()} cores)",
                    "depth_arel_vs_cpu": arel}

    if rank == 0:
        total_images = B * world * args.steps
        line = {
            "metric": "images/sec UniDepthV2.infer ViT-L/14 480x640", "value": total_images / (ms / 1000.0),
            "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16 operands, f32 accumulate/residual", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": B * world, "parallelism": f"dp{world}",
                       "l2": "per-step working set (weights 0.7 GB + activations > 4 GB) exceeds the 126 MB L2",
                       "cuda_graph": True, "engine": "udb_infer_v2 (one C call per infer)",
                       **({"collective": gather_kind()} if world > 1 else {})},
            "e2e": {"value": total_images / (ms_e2e / 1000.0), "unit": "images/s",
                    "h2d_bytes_per_step": rgb_host.numel(), "d2h_bytes_per_step": depth_host.numel() * 4 + k_host.numel() * 4},
            "gpu_launches": int(launches_per_step * args.steps),
            "clocks": clocks, "roofline": roof, "cpu_baseline": cpu_base,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
