#!/usr/bin/env python
"""bench.py -- ESRGAN 4x G/D training-step throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference|reference-cudnn]

Workload (BASELINE.json configs[1]): RRDBNet 23 blocks nf=64 + VGG19-conv5_4 perceptual loss +
Discriminator_VGG(256) vanilla RaGAN, 16 images / GPU, LR 64x64 -> HR 256x256, bf16 compute with
fp32 master weights, Adam.  One "step" = feed_data + SRModel.optimize_parameters (G update + D
update).  Synthetic data (torch.rand, seed 1234 + rank), random-init weights (no network here).

Prints ONE JSON line (rank 0): value = HR-pixels/s over all GPUs with inputs resident in HBM;
e2e = the same metric through the public API with pinned HOST batches (H2D inside the timed
region, loss read back every step); roofline = tensor-pipe roofline of the dominant kernel
(tcgen05 implicit-GEMM conv) from CUDA-event timings of every launch in one instrumented step;
cpu_baseline = the reference algorithm (oracle port, fp32 PyTorch CPU) on this box's host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

GFLOP_PER_IMAGE_STEP = 757.7  # SURVEY.md 8d / BASELINE.md 2: algorithmic conv FLOPs of one step, per image


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "reference-cudnn"])
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--nb", type=int, default=23)
    ap.add_argument("--hr", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "hbm_gbs": d["hbm_gbs"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0,
            "source": "fallback (B200_PROFILING.md)"}


def vgg_checkpoint():
    """Seeded random-init torchvision VGG19 (no pretrained download possible offline)."""
    import torchvision
    path = os.path.join(tempfile.gettempdir(), "b200_bench_vgg19_seed7.pth")
    if not os.path.exists(path):
        g = torch.Generator().manual_seed(7)
        net = torchvision.models.vgg19(weights=None)
        sd = net.state_dict()
        for k, v in sd.items():
            if v.dim() > 1:
                fan_in = v[0].numel()
                v.copy_(torch.randn(v.shape, generator=g) * (2.0 / fan_in) ** 0.5)
            else:
                v.zero_()
        torch.save(sd, path + ".tmp%d" % os.getpid())
        os.replace(path + ".tmp%d" % os.getpid(), path)
    return path


def make_opt(args, vgg_path):
    return {"model": "sr", "scale": 4, "is_train": True, "datasets": {"train": {"crop_size": args.hr}},
            "network_G": {"type": "esrgan", "nb": args.nb, "nf": 64, "gc": 32, "gaussian": False,
                          "upsample_mode": "upconv"},
            "network_D": {"type": "discriminator_vgg"},
            "train": {"pixel_criterion": "l1", "pixel_weight": 1e-2, "feature_criterion": "l1",
                      "feature_weight": 1, "gan_type": "vanilla", "gan_weight": 5e-3, "lr_G": 1e-4, "lr_D": 1e-4,
                      "perceptual_opt": {"pretrained_path": vgg_path}}}


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def timed_steps(fn, steps, warmup, world):
    import torch.distributed as dist
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
        ms = float(t.item())
    return ms / steps


def cpu_reference_steps(args, steps, warmup, threads=None):
    """The reference algorithm on host cores: oracle port (fp32 PyTorch CPU), batch 1 per step."""
    from collections import OrderedDict
    from oracle import esrgan_oracle as O
    import torchvision
    from trainner_b200 import networks
    from trainner_b200.architectures import discriminators, RRDBNet_arch
    # oneDNN convolutions at batch 1 stop scaling (and oversubscribe badly) past a few dozen threads
    torch.set_num_threads(threads or min(os.cpu_count(), 32))
    torch.manual_seed(0)
    g = RRDBNet_arch.RRDBNet(3, 3, 64, args.nb)
    networks.init_weights(g, "kaiming", 0.1)
    d = discriminators.Discriminator_VGG(args.hr, 3, 64)
    networks.init_weights(d, "kaiming", 0.1)
    tv = torch.load(vgg_checkpoint())
    vgg_sd = O.torchvision_vgg_to_feature_net(tv)
    orc = O.ESRGANStepOracle(OrderedDict(g.state_dict()), args.nb, OrderedDict(d.state_dict()), args.hr, vgg_sd)
    gen = torch.Generator().manual_seed(1234)
    lr_img = torch.rand(1, 3, args.hr // 4, args.hr // 4, generator=gen)
    hr_img = torch.rand(1, 3, args.hr, args.hr, generator=gen)
    for _ in range(warmup):
        orc.optimize_parameters(lr_img, hr_img)
    t0 = time.perf_counter()
    for _ in range(steps):
        orc.optimize_parameters(lr_img, hr_img)
    dt = (time.perf_counter() - t0) / steps
    return 1.0 / dt, dt, torch.get_num_threads()


def run_reference_cpu(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 8))
    ips, dt, threads = cpu_reference_steps(args, steps, max(1, min(args.warmup, 1)))
    val = ips * args.hr * args.hr
    sample = "batch 1 per step, %d timed steps (bounded sample of the %d-image step)" % (steps, args.batch)
    line = {"impl": "reference", "metric": "hr_pixels_per_sec", "value": val, "unit": "HR-px/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": 1, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "images_per_sec": ips,
            "config": {"workload": "ESRGAN 4x G/D step nb=%d HR %d^2 (oracle port of the reference, CPU)" % (args.nb, args.hr),
                       "global_batch": 1},
            "cpu_baseline": {"value": val, "unit": "HR-px/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "HR-px/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def run_reference_cudnn(args):
    """Context row (not the driver's reference arm): the reference's op sequence (oracle modules)
    on the GPU through stock PyTorch/cuDNN under bf16 autocast -- what the reference does on a B200."""
    from collections import OrderedDict
    from oracle import esrgan_oracle as O
    from trainner_b200 import networks
    from trainner_b200.architectures import discriminators, RRDBNet_arch
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(0)
    g = RRDBNet_arch.RRDBNet(3, 3, 64, args.nb)
    networks.init_weights(g, "kaiming", 0.1)
    d = discriminators.Discriminator_VGG(args.hr, 3, 64)
    networks.init_weights(d, "kaiming", 0.1)
    vgg_sd = O.torchvision_vgg_to_feature_net(torch.load(vgg_checkpoint()))
    orc = O.ESRGANStepOracle(OrderedDict(g.state_dict()), args.nb, OrderedDict(d.state_dict()), args.hr, vgg_sd,
                             device="cuda")
    gen = torch.Generator().manual_seed(1234)
    lr_img = torch.rand(args.batch, 3, args.hr // 4, args.hr // 4, generator=gen).cuda()
    hr_img = torch.rand(args.batch, 3, args.hr, args.hr, generator=gen).cuda()

    def step():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            orc.optimize_parameters(lr_img, hr_img)

    ms = timed_steps(step, args.steps, args.warmup, 1)
    ips = args.batch / (ms / 1e3)
    print(json.dumps({"impl": "reference-cudnn", "metric": "hr_pixels_per_sec", "value": ips * args.hr ** 2,
                      "unit": "HR-px/s", "images_per_sec": ips, "ms_per_step": ms, "n_gpus": 1, "steps": args.steps,
                      "warmup": args.warmup, "dtype": "bf16 autocast", "data": "synthetic",
                      "config": {"workload": "reference op sequence via PyTorch/cuDNN, nb=%d batch %d HR %d^2" %
                                 (args.nb, args.batch, args.hr)}}))


def ncu_traffic_per_launch(tag):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu
    capture of one training step (profiles/r01_v5_conv_flat_dram_traffic.txt); None for other kernels."""
    path = os.path.join(ROOT, "profiles", "r01_v5_conv_flat_dram_traffic.txt")
    if tag != "conv_flat" or not os.path.exists(path):
        return None, None
    import re
    m = re.search(r"DRAM read ([0-9.]+) MB, DRAM write ([0-9.]+) MB", open(path).read())
    if not m:
        return None, None
    return (float(m.group(1)) + float(m.group(2))) * 1e6, "profiles/r01_v5_conv_flat_dram_traffic.txt"


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference_cpu(args)
    if args.impl == "reference-cudnn":
        return run_reference_cudnn(args)

    from trainner_b200 import _lib
    from trainner_b200.models.sr_model import create_model
    from trainner_b200.parallel import init_distributed
    import torch.distributed as dist

    world = init_distributed()
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    torch.manual_seed(0)
    vgg_path = vgg_checkpoint() if rank == 0 or world == 1 else None
    if world > 1:
        dist.barrier()
        vgg_path = vgg_checkpoint()
    model = create_model(make_opt(args, vgg_path), device="cuda")
    gen = torch.Generator().manual_seed(1234 + rank)
    host = {"LR": torch.rand(args.batch, 3, args.hr // 4, args.hr // 4, generator=gen).pin_memory(),
            "HR": torch.rand(args.batch, 3, args.hr, args.hr, generator=gen).pin_memory()}
    dev = {k: v.cuda() for k, v in host.items()}
    counter = {"n": 0}

    def step_resident():
        counter["n"] += 1
        model.feed_data(dev)
        model.optimize_parameters(counter["n"])

    sink = {}

    def step_e2e():
        counter["n"] += 1
        model.feed_data(host)                     # pinned host -> device inside the timed region
        model.optimize_parameters(counter["n"])
        sink["log"] = model.get_current_log()     # device -> host read of the step's losses

    sampler = ClockSampler(local) if rank == 0 else None
    ms = timed_steps(step_resident, args.steps, args.warmup, world)
    ms_e2e = timed_steps(step_e2e, args.steps, max(1, args.warmup // 2), world)
    clocks = sampler.stop() if sampler else None

    gb = args.batch * world
    px = args.hr * args.hr
    ips, ips_e2e = gb / (ms / 1e3), gb / (ms_e2e / 1e3)
    pk = peaks()

    # ---- roofline leg: one instrumented step, CUDA events around every kernel launch of the plans
    roof = None
    cpu_base = None
    if True:  # every rank runs the instrumented step (it contains the gradient all-reduce); rank 0 reports
        from trainner_b200 import runtime
        rows = []
        orig_run = runtime.Plan.run

        def timed_run(self):
            rows.extend(self.run_timed())

        runtime.Plan.run = timed_run
        detail = [] if os.environ.get("B200_BENCH_DETAIL") else None
        runtime.Plan.detail_sink = detail
        l0 = _lib.launch_count()
        try:
            step_resident()
        finally:
            runtime.Plan.run = orig_run
            runtime.Plan.detail_sink = None
        # kernels of libtrainner_b200.so per step, counted at the C entry points during this eagerly executed
        # step (the timed steps replay the same launches from CUDA graphs, which bypass the host-side counter)
        launches = (_lib.launch_count() - l0) * args.steps
        if detail is not None and rank == 0:
            grp = {}
            for tag, info, t_ms, fl in detail:
                a = grp.setdefault((tag, info), [0, 0.0, 0.0])
                a[0] += 1; a[1] += t_ms; a[2] += fl
            with open(os.environ["B200_BENCH_DETAIL"], "w") as fh:
                for (tag, info), (c, t, fl) in sorted(grp.items(), key=lambda kv: -kv[1][1]):
                    fh.write("%-26s %-44s n=%4d  %8.3f ms  %7.1f TF/s\n" % (tag, info, c, t, fl / (t * 1e-3) / 1e12 if t > 0 else 0))
        torch.cuda.synchronize()
        agg = {}
        for tag, t_ms, fl in rows:
            a = agg.setdefault(tag, [0, 0.0, 0.0])
            a[0] += 1
            a[1] += t_ms
            a[2] += fl
        total_ms = sum(a[1] for a in agg.values())
        dom = max(agg.items(), key=lambda kv: kv[1][1])
        tag, (cnt, t_ms, fl) = dom
        achieved = fl / (t_ms * 1e-3) / 1e12 if t_ms > 0 else 0.0
        peak = pk["bf16_tflops_sustained"]
        traffic, traffic_src = ncu_traffic_per_launch(tag)
        roof = {"bound": "tensor", "kernel": tag, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_unit": "DRAM bytes per launch (ncu, mean)",
                "traffic_source": traffic_src, "launches_per_step": cnt,
                "avg_launch_ms": t_ms / cnt, "share_of_kernel_time": t_ms / total_ms if total_ms else None,
                "peak_source": pk["source"] + ", sustained figure (kernel timed inside a long step)",
                "per_kernel": {k: {"launches": v[0], "ms": round(v[1], 3),
                                   "tflops": round(v[2] / (v[1] * 1e-3) / 1e12, 1) if v[1] > 0 and v[2] > 0 else None}
                               for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])},
                "step_algorithmic_tflops": GFLOP_PER_IMAGE_STEP * 1e-3 * args.batch,
                "step_frac_of_peak": (GFLOP_PER_IMAGE_STEP * 1e9 * (args.batch / (ms / 1e3))) / (peak * 1e12)}
        if not args.no_cpu_baseline and world == 1 and rank == 0:
            c_ips, c_dt, threads = cpu_reference_steps(args, 2, 1)
            cpu_base = {"value": c_ips * px, "unit": "HR-px/s", "images_per_sec": c_ips, "cores": threads,
                        "kind": "port", "sample": "2 timed steps at batch 1 (nb=%d, HR %d^2), fp32" % (args.nb, args.hr)}
    if world > 1:
        dist.barrier()
    if rank == 0:
        h2d = sum(v.numel() * v.element_size() for v in host.values())
        line = {"metric": "hr_pixels_per_sec", "value": ips * px, "unit": "HR-px/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "images_per_sec": ips,
                "config": {"workload": "ESRGAN 4x G/D step: RRDBNet nb=%d nf=64 + VGG19 conv5_4 L1 + Discriminator_VGG(%d) "
                                       "RaGAN, L1 pixel; LR %d^2 -> HR %d^2" % (args.nb, args.hr, args.hr // 4, args.hr),
                           "global_batch": gb, "per_gpu_batch": args.batch, "parallelism": "dp%d" % world,
                           "l2": "per-step working set (>2 GB of activations) exceeds the 126 MB L2; no explicit flush"},
                "e2e": {"value": ips_e2e * px, "unit": "HR-px/s", "images_per_sec": ips_e2e, "ms_per_step": ms_e2e,
                        "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4 * len(sink.get("log", {}))},
                "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu_base}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
