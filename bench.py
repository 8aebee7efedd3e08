#!/usr/bin/env python
"""bench.py -- frame-pairs/s of the CMFlow forward hot path (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch 256] [--points 256] [--model cmflow|cmflow_t|raflow]
                    [--precision fp16x3|tf32x3|fp32] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = one CMFlow.forward over a batch of synthetic radar frame pairs (BASELINE.json configs[1]:
N=256 points, batch=256 per GPU; weak scaling: every rank runs its own 256 pairs, no data-path collective --
SURVEY.md 8e).  Rank 0 prints ONE JSON line on stdout (everything else, e.g. NCCL's banner, goes to stderr).
The other BASELINE.json configurations are reachable for the record (--model cmflow_t: a step is the three forwards
of a clip batch; --points 4096 --batch 64: the dense cloud), but the bench line is configs[1].

  value  : device-resident throughput (inputs in HBM before the timed region), CUDA events, max over ranks
  e2e    : same metric through the public host-buffer call (cmf_model_forward_host): pinned H2D of the four
           input tensors + forward + D2H of the four outputs inside the timed region, every step
  roofline / kernels : per-kernel-category device time measured live with CUDA events on the launching stream
           in a separate profiled pass of the same workload (event pairs around every launch).  roofline = the
           dominant kernel (set-conv #2 layer 2): algorithmic TFLOP/s against the measured cuBLAS bf16 rate
           (`frac`), against that rate / 3 (`frac_of_ceiling`: the 3-MMA split), `traffic` = DRAM bytes per launch
           from the committed ncu capture, and `hbm` = SURVEY 8d's algorithmic bytes over the launch time
  cpu_baseline : the oracle (CPU port of the reference's PyTorch path) on a bounded sample, rank 0, N=1 only
  ref_cuda_baseline : informational -- the reference's own ball-query kernel (compiled unmodified) under the
           PyTorch-eager unfused model on the same GPU (north_star's "reference's own lib/src CUDA build")
  --impl reference : the reference arm = that same CPU port timed with all host threads (the reference's
           Python cannot travel to the GPU box and ships no CPU kernels of its own -- SURVEY.md 8c)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained"),
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 50 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def time_cpu_port(pairs, points, steps, warmup, threads):
    """The oracle (CPU port of the reference forward) on `pairs` synthetic pairs per step."""
    from cmflow_b200.synth import make_pairs, synthetic_state_dict
    from oracle import cmflow_oracle as O
    torch.set_num_threads(threads)
    sd = synthetic_state_dict(0)
    pc1, pc2, ft1, ft2, _ = make_pairs(pairs, points, seed=1234)
    with torch.no_grad():
        for _ in range(warmup):
            O.cmflow_forward(sd, pc1, pc2, ft1, ft2)
        t0 = time.perf_counter()
        for _ in range(steps):
            O.cmflow_forward(sd, pc1, pc2, ft1, ft2)
        dt = time.perf_counter() - t0
    return pairs * steps / dt, dt / steps


def time_ref_cuda(pairs, points, steps, warmup, dev):
    """Informational GPU baseline (north_star's "reference's own lib/src CUDA build"): the reference's ball-query kernel compiled UNMODIFIED
    (oracle/_ref/libpointnet2_ref.so) under the PyTorch-eager, unfused model (oracle/cmflow_oracle.py evaluated on CUDA tensors: cuBLAS fp32
    for the 1x1 convs, torch.gather grouping, square_distance + topk k-NN as radarflow_util.py:8-30,88-99).  Part of the baseline leg: rank 0,
    N=1, bounded sample; the oracle stays the checker, never the product path."""
    import types
    from cmflow_b200.synth import make_pairs, synthetic_state_dict
    from oracle import cmflow_oracle as O
    from oracle import refcuda as R
    if not R.available():
        return None

    def knn_point(nsample, xyz, new_xyz):            # radarflow_util.py:88-99 (square_distance + topk)
        d = -2 * torch.matmul(new_xyz, xyz.permute(0, 2, 1))
        d = d + torch.sum(new_xyz ** 2, -1).unsqueeze(-1) + torch.sum(xyz ** 2, -1).unsqueeze(1)
        dist, idx = torch.topk(torch.clamp(d, min=0.0), nsample, dim=-1, largest=False, sorted=False)
        return idx.int(), dist

    sd = {k: v.to(dev) for k, v in synthetic_state_dict(0).items()}
    pc1, pc2, ft1, ft2 = (t.to(dev) for t in make_pairs(pairs, points, seed=1234)[:4])
    saved = O.P
    O.P = types.SimpleNamespace(ball_query=R.ball_query, knn_point=knn_point)
    try:
        with torch.no_grad():
            for _ in range(warmup):
                O.cmflow_forward(sd, pc1, pc2, ft1, ft2)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(steps):
                O.cmflow_forward(sd, pc1, pc2, ft1, ft2)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
    finally:
        O.P = saved
    return {"value": pairs / (ms * 1e-3), "unit": "frame-pairs/s", "ms_per_step": ms, "kind": "reference lib/src ball-query kernel (unmodified, sm_100a) + PyTorch-eager fp32 model on the same GPU",
            "sample": f"{pairs} pairs/step x {steps} steps (N={points}); the unfused model materialises 34 MB/pair of grouped tensors at K=32"}


def emit(line):
    """The ONE JSON line goes to the real stdout; everything else that libraries print (NCCL's version banner...) was diverted to stderr."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=256, help="frame pairs per GPU per step")
    ap.add_argument("--points", type=int, default=256)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="cmflow", choices=["cmflow", "cmflow_t", "raflow"],
                    help="cmflow = the headline workload (BASELINE.json configs[1-2]); cmflow_t = 3-frame clips with the GRU state carried "
                         "(configs[3]: a step is the three forwards of a clip batch); raflow = models/raflow.py.  --points 4096 --batch 64 is configs[4]")
    ap.add_argument("--precision", default=os.environ.get("CMF_BENCH_PRECISION", "fp16x3"), choices=["fp32", "tf32x3", "fp16x3"],
                    help="fp32 = strict fp32 FMA kernels; tf32x3 / fp16x3 = tcgen05 tensor cores with a 22-bit hi/lo operand split "
                         "(3 MMAs per product, fp32 accumulate, fp32-class accuracy) in kind::tf32 or kind::f16")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ncu", action="store_true", help="profiler harness: W+K device forwards only, prints no bench line")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    W = max(args.warmup, 3) if (args.impl == "ours" and not args.ncu) else args.warmup
    K = args.steps
    # host threads for the CPU arms: all cores up to 32 -- beyond that the small per-pair matmuls of this workload slow down
    # (measured on the 128-core GPU host: 0.37 pairs/s with 128 threads vs ~10 with 32); the count used is reported in `cores`
    cores = min(os.cpu_count() or 1, 32)
    mname = {"cmflow": "CMFlow forward", "cmflow_t": "CMFlow-T temporal forward, 3-frame clips,", "raflow": "RaFlow forward"}[args.model]
    workload = f"{mname} synthetic radar pairs N={args.points}, batch={args.batch}/GPU, {args.gpus}xB200"

    if args.impl == "reference":
        if rank != 0:
            return
        # bounded sample: ~1 s of CPU work per step at N=256 (the whole --steps K run stays within a few minutes)
        sample = max(1, min(16, (16 * 256 * 256) // (args.points * args.points)))
        v, spp = time_cpu_port(sample, args.points, K, min(W, 1), cores)
        emit(({
            "impl": "reference", "metric": "frame-pairs/sec CMFlow forward", "value": v, "unit": "frame-pairs/s", "n_gpus": args.gpus,
            "steps": K, "warmup": min(W, 1), "ms_per_step": spp * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic radar pairs, seeded random-init weights",
            "config": {"workload": workload, "points": args.points, "pairs_per_step": sample},
            "cpu_baseline": {"value": v, "unit": "frame-pairs/s", "cores": cores, "kind": "port",
                             "sample": f"{sample} pairs/step x {K} steps, oracle/cmflow_oracle.py (CPU port of the reference forward), {cores} threads"},
            "e2e": {"value": v, "unit": "frame-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    assert torch.cuda.is_available(), "bench.py --impl ours needs a GPU (there is no CPU path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    from cmflow_b200.cmflow import CMFlow, CMFlow_T, RaFlow
    from cmflow_b200.synth import make_pairs, raflow_state_dict, synthetic_state_dict

    class A:
        num_points = args.points
        stat_thres = 0.5
        rigid_thres = 0.15

    if args.model == "cmflow_t":
        net = CMFlow_T(A()); net.load_state_dict(synthetic_state_dict(0, temporal=True))
    elif args.model == "raflow":
        net = RaFlow(A()); net.load_state_dict(raflow_state_dict(0))
    else:
        net = CMFlow(A()); net.load_state_dict(synthetic_state_dict(0))
    net = net.to(dev)
    FRAMES = 3 if args.model == "cmflow_t" else 1             # forwards per step
    net.set_precision(args.precision)
    B, N = args.batch, args.points
    NSETS = 4                                                  # rotate distinct input batches
    host_sets = [tuple(t.pin_memory() for t in make_pairs(B, N, seed=1234 + 97 * rank + s, dense=(N >= 2048))[:4]) for s in range(NSETS)]
    dev_sets = [tuple(t.to(dev) for t in hs) for hs in host_sets]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    interval = torch.full((args.batch,), 0.1, device=dev)

    def fwd_dev(i):
        with torch.no_grad():
            if args.model == "cmflow_t":                      # one clip step: gfeat None -> carry -> carry (clip_util.py:218-233)
                g = None
                for f in range(FRAMES):
                    out = net(*dev_sets[(i + f) % NSETS], None, "test", g)
                    g = out[4]
                return out
            if args.model == "raflow":
                return net(*dev_sets[i % NSETS], interval)
            return net(*dev_sets[i % NSETS], None, "test")

    out_host = None

    def fwd_host(i):
        nonlocal out_host
        if args.model == "raflow":                            # no host-buffer entry point for RaFlow: explicit pinned copies around the device call
            ins = [t.to(dev, non_blocking=True) for t in host_sets[i % NSETS]]
            with torch.no_grad():
                res = net(*ins, interval)
            out_host = [r.to("cpu", non_blocking=False) for r in res]
            return out_host
        g = None
        for f in range(FRAMES):
            out_host = net.forward_host(*host_sets[(i + f) % NSETS], gfeat=g, out=out_host)
            g = out_host["gfeat"] if args.model == "cmflow_t" else None
        return out_host

    def timed(fn, steps):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        for i in range(steps):
            flush.zero_()                                       # L2 flush between timed iterations (outside the events)
            ev[i][0].record()
            fn(i)
            ev[i][1].record()
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in ev)
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    for i in range(W):
        fwd_dev(i)
    torch.cuda.synchronize()
    if args.ncu:                      # under a profiler: just run K more forwards and leave (never a bench value)
        for i in range(K):
            fwd_dev(i)
        torch.cuda.synchronize()
        emit({"ncu_harness": True, "launches_per_step": net.launches_per_forward()})
        return
    sampler = ClockSampler(local) if rank == 0 else None
    ms_dev = timed(fwd_dev, K)
    launches = net.launches_per_forward()
    for i in range(W):
        fwd_host(i)
    ms_host = timed(fwd_host, K)
    clocks = sampler.stop() if sampler else None
    total_pairs = B * world * FRAMES

    # profiled pass: per-category device time (CUDA events around every launch, same stream)
    prof, cpu = None, None
    if rank == 0:
        net.set_profiling(True)
        acc = {}
        PR = 3
        for i in range(PR):
            flush.zero_()
            fwd_dev(i)
            for k, (ms, cnt, work) in net.read_profile().items():
                a = acc.setdefault(k, [0.0, 0, 0.0])
                a[0] += ms; a[1] += cnt; a[2] += work
        net.set_profiling(False)
        prof = {k: {"ms_per_step": v[0] / PR, "launches_per_step": v[1] // PR, "gflop_per_step": v[2] / PR / 1e9} for k, v in acc.items()}
    if dist is not None:
        dist.barrier()

    if rank == 0:
        peaks = load_peaks()
        tot_ms = sum(v["ms_per_step"] for v in prof.values())
        for v in prof.values():
            v["share"] = v["ms_per_step"] / tot_ms
            if v["gflop_per_step"] > 0:
                v["tflops"] = v["gflop_per_step"] / v["ms_per_step"]
        dom = "gemm_setconv2_l2"
        d = prof[dom]
        achieved = d["gflop_per_step"] / d["ms_per_step"]       # GFLOP/ms == TFLOP/s (algorithmic FLOPs: 2*M*K*cols, split passes not counted)
        tc = args.precision != "fp32"
        split = {"tf32x3": ("3xTF32", 6), "fp16x3": ("3xFP16", 3)}.get(args.precision)
        peak_tf = peaks["bf16_tflops_sustained"] or peaks["bf16_tflops"]
        traffic, traffic_src = None, None
        tj = os.path.join(ROOT, "profiles", "r01f_tc_gemm2_traffic.json")
        if args.precision == "fp16x3" and B == 256 and N == 256 and os.path.exists(tj):      # the ncu capture is of exactly this workload
            tjd = json.load(open(tj))
            traffic, traffic_src = tjd["traffic_bytes_per_launch_avg"], tjd["source"]
        # algorithmic bytes of the kernel: every (point, neighbour) column gathers one 512-channel fp32 layer-1 row (2 KB) and
        # writes 256 channels (1 KB); 60 columns per point over the four scales, 4 launches
        alg_bytes_launch = B * N * 60 * 3072 / max(1, d["launches_per_step"])
        roofline = {"kernel": (f"tc_gemm2_kernel<SC2_Y1> tcgen05 {split[0]}" if tc else "gemm_nt_kernel<128> fp32 FMA") +
                              " (set-conv #2 layer 2, 512->256 over N*K neighbour columns, gather fused)" ,
                    "bound": "tensor", "achieved": achieved, "peak": peak_tf,
                    "unit": "TFLOP/s", "frac": achieved / peak_tf,
                    "peak_source": peaks["source"] + " cuBLAS bf16 (sustained)", "traffic": traffic, "traffic_source": traffic_src,
                    "ceiling": peak_tf / split[1] if tc else 74.5, "frac_of_ceiling": achieved / (peak_tf / split[1] if tc else 74.5),
                    "hbm": {"algorithmic_bytes_per_launch": alg_bytes_launch,
                            "achieved_gbs": alg_bytes_launch / (d["ms_per_step"] / max(1, d["launches_per_step"]) * 1e-3) / 1e9,
                            "peak_gbs": peaks["hbm_gbs"], "frac": alg_bytes_launch / (d["ms_per_step"] / max(1, d["launches_per_step"]) * 1e-3) / 1e9 / peaks["hbm_gbs"],
                            "note": "not the binding bound (86 flop/B); gathered rows are served by L2, see profiles/r01b_tc_kernels_ncu_full.md"},
                    "launches_per_step": d["launches_per_step"], "avg_launch_ms": d["ms_per_step"] / max(1, d["launches_per_step"]),
                    "note": (f"{split[0]}: three MMAs per algorithmic MAC" + (" (kind::tf32 runs at half the bf16 rate)" if args.precision == "tf32x3" else "") + f" => ceiling = bf16 peak / {split[1]}" if tc else
                             "strict-fp32 FMA build (no tensor cores): the chip's fp32 FMA ceiling is 74.5 TFLOP/s, ~1/19 of this peak")}
        if not args.no_cpu_baseline and world == 1:
            cs = max(1, min(16, (16 * 256 * 256) // (N * N)))
            v, spp = time_cpu_port(cs, N, 12, 1, cores)          # ~10-20 s of CPU work
            cpu = {"value": v, "unit": "frame-pairs/s", "cores": cores, "kind": "port",
                   "sample": f"{cs} pairs/step x 12 steps (N={N}), oracle/cmflow_oracle.py CPU port of the reference forward, {cores} threads"}
        ref_cuda = None
        if not args.no_cpu_baseline and world == 1:
            try:
                ref_cuda = time_ref_cuda(max(1, min(32, (32 * 256 * 256) // (N * N))), N, 5, 2, dev)
            except Exception as e:                      # informational only
                ref_cuda = {"unavailable": repr(e)[:200]}
        h2d = FRAMES * 4 * B * 3 * N * 4
        d2h = FRAMES * (B * 3 * N * 4 + B * N * 4 + B * 16 * 4 + B * N)
        line = {
            "metric": "frame-pairs/sec CMFlow forward", "value": total_pairs * K / (ms_dev / 1e3), "unit": "frame-pairs/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": (f"f32 ({split[0]} split on tcgen05 tensor cores, fp32 accumulate)" if tc else "f32"), "data": "synthetic radar pairs (cmflow_b200/synth.py), seeded random-init weights of the CMFlow architecture",
            "config": {"workload": workload, "points": N, "pairs_per_gpu": B, "global_batch": total_pairs, "parallelism": f"dp{world}",
                       "l2": "256 MB flush between timed steps; 4 rotating input batches", "precision_mode": args.precision},
            "e2e": {"value": total_pairs * K / (ms_host / 1e3), "unit": "frame-pairs/s", "ms_per_step": ms_host / K,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches * K * FRAMES, "launches_per_step": launches * FRAMES,
            "clocks": clocks, "roofline": roofline, "kernels": prof, "cpu_baseline": cpu, "ref_cuda_baseline": ref_cuda,
            "workspace_bytes": net.workspace_bytes(),
        }
        emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
