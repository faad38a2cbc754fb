#!/usr/bin/env python
"""Benchmark of the legacy per-image scoring pass (BASELINE.json metric: images/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]

One "step" = one pass of the hot path over one batch of B synthetic 24 MP frames per GPU:
technical metrics (csrc/tech_stats.cu) + perceptual hash (csrc/phash.cu) + CLIP preprocess
(csrc/preprocess.cu, csrc/resample_tc.cu) + ViT-L/14 tower,
aesthetic head and tag similarities (csrc/gemm.cu, csrc/vit.cu) + the similarity stage on the
step's embeddings (all-gather across ranks, cosine pairs, csrc/gemm.cu + csrc/similarity.cu).
`value` is the whole-job rate with the frames already resident in HBM; `e2e` is the same pass
through the host-buffer pipeline (pinned host frames, H2D and D2H inside the timed region).
Prints ONE JSON line on rank 0.  `--impl reference` times the reference's CPU path (oracle port;
/root/reference does not exist on the GPU box) on a bounded sample instead.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W = 4000, 6000
GFLOP_PER_IMAGE = 162.0          # ViT-L/14 224 px incl. attention and patch embedding
GEMM_GFLOP_PER_IMAGE = 24 * (1.617 + 0.539 + 4.312) + 0.308   # the part the tcgen05 GEMM kernel executes
TECH_BYTES_PER_IMAGE = 3 * H * W


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=128, help="24 MP frames per GPU per step")
    ap.add_argument("--impl", default="facet_b200", choices=["facet_b200", "reference"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--clock-period-ms", type=float, default=50.0, help="NVML sampling period during the timed region (0 = off)")
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "source": "MEASURED_PEAKS.json (sustained bf16, copy HBM)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1400.0, "source": "fallback of B200_PROFILING.md (MEASURED_PEAKS.json absent)"}


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.  NVML is queried in-process
    (nvidia_ml_py): one clock read + one reasons read per sample cost microseconds, whereas an `nvidia-smi -lms`
    poller stalls the driver for milliseconds per sample and slows every launch of the step it lands in
    (measured: +20 % on a 10-step run).  Falls back to nvidia-smi if NVML is not importable."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index, period_s=0.02):
        self.idx = gpu_index
        self.period = period_s
        self.samples = []          # (sm_mhz, reasons bitmask)
        self.max_mhz = None
        self.proc = None
        self.lines = []
        self.stop_flag = threading.Event()
        self.thread = None
        self.nvml = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "250"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read_smi, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _poll_nvml(self):
        n = self.nvml
        while not self.stop_flag.is_set():
            try:
                mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                reasons = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.samples.append((mhz, reasons))
            except Exception:
                pass
            self.stop_flag.wait(self.period)

    def reset(self):
        """Drop what was sampled so far (called right before the timed region starts)."""
        self.samples = []

    def _read_smi(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        self.stop_flag.set()
        if self.nvml is not None:
            self.thread.join(timeout=2)
            n = self.nvml
            names = {"hw_slowdown": getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                     "hw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                     "sw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                     "sw_power_cap": getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
            sm = sorted(s[0] for s in self.samples)
            mask = 0
            for _, r in self.samples:
                mask |= r
            reasons = sorted(k for k, bit in names.items() if mask & bit)
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_min_mhz": sm[0] if sm else None,
                    "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(sm), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1]))
                mx = float(parts[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        # median of the upper half: samples taken while the GPU was busy
        busy = sm[len(sm) // 2:] if sm else []
        med = busy[len(busy) // 2] if busy else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


def make_pool(n, seed, device):
    """n distinct device-resident 24 MP BGR frames (gradients + texture + noise; 1 in 8 pure noise, 1 in 8 flat)."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    out = torch.empty((n, H, W, 3), dtype=torch.uint8, device=device)
    yy = torch.linspace(0, 1, H, device=device)[:, None, None]
    xx = torch.linspace(0, 1, W, device=device)[None, :, None]
    for i in range(n):
        kind = i % 8
        if kind == 6:
            out[i] = torch.randint(0, 256, (H, W, 3), dtype=torch.uint8, device=device, generator=g)
            continue
        ph = torch.rand((1, 1, 3), device=device, generator=g) * 6.28
        fr = 1 + 5 * torch.rand((1, 1, 3), device=device, generator=g)
        base = 0.5 + 0.25 * torch.sin(fr * 6.28 * xx + ph) + 0.22 * torch.cos(fr * 4.1 * yy + 2 * ph)
        if kind == 3:
            base = torch.round(base * 5) / 5
        gain = 0.6 + 0.5 * float(torch.rand(1, device=device, generator=g))
        sigma = (0.0, 2.0, 4.0, 1.0, 6.0, 3.0, 0.0, 1.5)[kind]
        noise = torch.randn((H, W, 3), device=device, generator=g) * sigma
        out[i] = (base * 255 * gain + noise).clamp(0, 255).to(torch.uint8)
        if kind == 0:
            out[i, :, :, 0] = out[i, :, :, 1]
            out[i, :, :, 2] = out[i, :, :, 1]
    return out


def run_reference(args, rank, world):
    """CPU arm: the reference's own CPU path (oracle port: cv2/NumPy/SciPy/PIL/torch-CPU calls in the
    order of processing/batch_processor.py:198-233) timed on the host cores.  One 24 MP frame per step."""
    if rank != 0:
        return
    import numpy as np
    import torch
    from facet_b200.models.clip_vit import random_state_dict
    from facet_b200.synth import synth_embeddings, synth_image_bgr
    from oracle import cpu_port, grouping
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = random_state_dict(0)
    tags = torch.from_numpy(synth_embeddings(240, seed=7, cluster_fraction=0.0))
    frames = [synth_image_bgr(2000 + i, H, W) for i in range(2)]
    embs = []

    def step(i):
        _, vit = cpu_port.score_images_cpu([frames[i % len(frames)]], sd, tags)
        embs.append(vit["embedding"].numpy())
        e = np.concatenate(embs[-64:], axis=0)
        grouping.cosine_pairs(e, 0.9)

    for i in range(args.warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(i)
    dt = time.perf_counter() - t0
    val = args.steps / dt
    line = {
        "impl": "reference", "metric": "images/sec", "value": val, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(1, note="CPU arm: one 24 MP frame per step (bounded sample of the same workload)"),
        "cpu_baseline": {"value": val, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} steps x 1 synthetic 24 MP frame, oracle/cpu_port.py (cv2+NumPy+PIL+torch fp32)"},
        "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(batch, note=None):
    cfg = {"workload": "full legacy scoring pass on synthetic 6000x4000 (24 MP) BGR frames: technical metrics + pHash + CLIP "
                       "preprocess + ViT-L/14 224px + MLP aesthetic + tag similarities + Hamming and cosine duplicate pairs "
                       "(BASELINE.json configs[4], per-GPU share)",
           "image": [H, W, 3], "batch_per_gpu": batch, "parallelism": "data-parallel, one rank per GPU",
           "l2": "inputs larger than L2 (pool of batch x 72 MB frames per step)", "weights": "random-init ViT-L/14 (seed 0)",
           "precision": "ViT GEMM operands fp16 (the reference's CUDA precision, scorer.py:515), fp32 accumulate and residual "
                        "stream; technical metrics / preprocess exact integers"}
    if note:
        cfg["note"] = note
    return cfg


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from facet_b200 import _lib, ops
    from facet_b200.models.clip_vit import random_state_dict
    from facet_b200.processing.pipeline import ScoringPipeline
    from facet_b200.processing.scorer import Facet
    from facet_b200.synth import synth_embeddings
    from facet_b200.utils.duplicate import all_gather_embeddings

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    from facet_b200.processing.pipeline import bind_to_gpu_numa_node
    orig_affinity = os.sched_getaffinity(0)
    numa_node = bind_to_gpu_numa_node(local_rank)      # pinned e2e buffers local to the GPU's PCIe root
    if world > 1:
        # NCCL prints its version banner on stdout at NCCL_DEBUG=VERSION; stdout must carry one JSON line only
        # stdout must carry exactly one JSON line: send NCCL's start-up banner ("NCCL version ...", printed on
        # stdout when the communicator is created) to stderr by redirecting fd 1 during initialisation
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=device)
            warm = torch.zeros(1, device=device)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    lib = _lib.load()

    B = args.batch
    tags = synth_embeddings(240, seed=7, cluster_fraction=0.0)
    scorer = Facet(random_state_dict(0), text_embeddings=tags, tag_names=[f"tag{i // 4}" for i in range(240)], device=device)
    pool = make_pool(B, seed=2000 + 17 * rank, device=device)
    torch.cuda.synchronize()

    def step():
        out = scorer.score_images_device(pool)                      # technical + pHash + preprocess + ViT/heads/tags
        emb = all_gather_embeddings(out["embedding"]) if world > 1 else out["embedding"]
        pairs, _ = ops.cosine_pairs(emb, 0.90, part=rank, nparts=world)
        hashes = all_gather_embeddings(out["phash"].view(-1, 1)).view(-1) if world > 1 else out["phash"]
        ops.hamming_pairs(hashes, 6, part=rank, nparts=world)      # duplicate rule of utils/duplicate.py on the step's hashes
        return out, pairs

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    # ---- timed region: K steps, CUDA events on the launching stream -------------------------------------
    launches0 = int(lib.fb_launch_count())
    # one sampler per node (rank 0's GPU): concurrent nvidia-smi pollers slow every rank's launches
    sampler = ClockSampler(local_rank, period_s=max(args.clock_period_ms, 1.0) * 1e-3)
    if rank == 0 and args.clock_period_ms > 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.reset()
    e0.record()
    for _ in range(args.steps):
        out, pairs = step()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop()
    launches = int(lib.fb_launch_count()) - launches0
    if world > 1:
        t = torch.tensor([ms_total], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * B / (ms_step * 1e-3)

    # ---- per-kernel device times of the SAME steps: the library records a CUDA-event pair around each of
    # its launches on the launching stream (fb_profile_*); K more steps, identical inputs -------------------
    _lib.profile_enable(True)
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pe0.record()
    for _ in range(args.steps):
        step()
    pe1.record()
    prof = _lib.profile_read()
    _lib.profile_enable(False)
    ms_step_profiled = pe0.elapsed_time(pe1) / args.steps
    per_step = {k: (v[0] / args.steps, v[1] // args.steps) for k, v in prof.items() if v[1]}
    ms_tech, ms_pre = per_step["technical"][0], per_step["preprocess"][0]
    ms_phash = per_step.get("other", (0.0, 0))[0]
    ms_vit = sum(per_step[k][0] for k in ("im2col", "gemm", "layernorm", "attention", "vit_tail") if k in per_step)
    gemm_ms, gemm_launches = per_step["gemm"]
    stages = {"technical_ms": ms_tech, "phash_ms": ms_phash, "preprocess_ms": ms_pre, "vit_ms": ms_vit,
              "kernel_ms_per_step": {k: round(v[0], 4) for k, v in per_step.items()},
              "launches_per_step": {k: v[1] for k, v in per_step.items()},
              "step_ms_with_event_pairs": ms_step_profiled,
              "technical_gbs": B * TECH_BYTES_PER_IMAGE / (ms_tech * 1e-3) / 1e9,
              "vit_tflops": B * GFLOP_PER_IMAGE * 1e9 / (ms_vit * 1e-3) / 1e12}
    peaks = measured_peaks()
    gemm_tflops = B * GEMM_GFLOP_PER_IMAGE * 1e9 / (gemm_ms * 1e-3) / 1e12
    # DRAM traffic of the GEMM launches of one layer at batch 128, from profiles/r1_gemm_ncu_v2.txt (ncu --set full)
    traffic_per_layer_b128 = (73.8 + 153.7 + 204.3 + 82.5 + 75.9 + 222.8 + 436.6 + 106.3) * 1e6
    roofline = {"bound": "tensor", "kernel": f"gemm_bf16_kernel (tcgen05), {gemm_launches} launches per step",
                "achieved": gemm_tflops, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                "frac": gemm_tflops / peaks["bf16_tflops"],
                "traffic": (24 * traffic_per_layer_b128 * B / 128) if world >= 1 else None,
                "traffic_note": "bytes per step, scaled from the ncu capture at batch 128 (24 layers x 4 launches)",
                "algorithmic_flops_per_step": B * GEMM_GFLOP_PER_IMAGE * 1e9, "kernel_ms_per_step": gemm_ms,
                "timing": "CUDA-event pairs around every launch of the kernel, on the launching stream, over K steps",
                "peak_source": peaks["source"], "share_of_step": gemm_ms / ms_step_profiled,
                "technical_kernel": {"bound": "hbm", "achieved": stages["technical_gbs"], "peak": peaks["hbm_gbs"],
                                     "unit": "GB/s", "frac": stages["technical_gbs"] / peaks["hbm_gbs"],
                                     "algorithmic_bytes_per_step": B * TECH_BYTES_PER_IMAGE,
                                     "traffic": (610.1e6 + 144.9e6) / 8 * B,
                                     "traffic_note": "ncu dram bytes read + written (frames + the luma plane the pass also emits "
                                                     "for the pHash), 8-frame capture of profiles/r1_tech_stats_ncu_v2.txt scaled"}}

    # ---- e2e: pinned host frames -> pipeline (H2D + kernels + D2H inside the timed region) ---------------
    e2e = None
    if not args.no_e2e:
        pipe = ScoringPipeline(scorer, chunk=8)
        # pinned host frames: a 32-frame buffer (2.3 GB per rank) sent B/32 times per step keeps the host
        # footprint bounded at N=8 while every step still moves B x 72 MB over PCIe
        eb = min(B, 32)
        reps = max(1, B // eb)
        host = torch.empty((eb, H, W, 3), dtype=torch.uint8, pin_memory=True)
        host.copy_(pool[:eb])
        torch.cuda.synchronize()
        pipe.run_host(host)          # warm-up
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_e2e = max(2, min(args.steps, 5))
        d2h = 0
        a.record()
        for _ in range(n_e2e):
            # the step's B frames reach the pipeline as B/32 host batches, back to back like a loader delivers them
            res = pipe.run_host_stream(host for _r in range(reps))
            embs, hashes_h = [res["embedding"]], [res["phash"]]
            emb = torch.from_numpy(np.concatenate(embs)).to(device)
            emb = all_gather_embeddings(emb) if world > 1 else emb
            p_, _ = ops.cosine_pairs(emb, 0.90, part=rank, nparts=world)
            hh = torch.from_numpy(np.concatenate(hashes_h).view(np.int64)).to(device)
            hh = all_gather_embeddings(hh.view(-1, 1)).view(-1) if world > 1 else hh
            q_ = ops.hamming_pairs(hh, 6, part=rank, nparts=world)
            d2h = int(p_.cpu().numel() + q_.cpu().numel()) * 4
        b.record()
        barrier()
        ms_e = a.elapsed_time(b) / n_e2e
        if world > 1:
            t = torch.tensor([ms_e], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_e = float(t.item())
        frames_per_step = eb * reps
        e2e = {"value": world * frames_per_step / (ms_e * 1e-3), "unit": "images/s", "steps": n_e2e,
               "frames_per_step_per_gpu": frames_per_step,
               "numa_node_rank0": numa_node,
               "h2d_bytes_per_step": pipe.h2d_bytes(frames_per_step, H, W) + frames_per_step * (768 * 4 + 8),
               "d2h_bytes_per_step": pipe.d2h_bytes(frames_per_step, 240) + d2h}
        del host

    # ---- CPU baseline beside it (rank 0, N=1 only): bounded sample of the same workload -----------------
    cpu_baseline = None
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        os.sched_setaffinity(0, orig_affinity)      # the CPU arm may use every host core again
        from oracle import cpu_port
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        sd = random_state_dict(0)
        frames = [pool[i].cpu().numpy() for i in (1, 2, 4)]
        cpu_port.score_images_cpu(frames[:1], sd, torch.from_numpy(tags))
        t0 = time.perf_counter()
        tech_cpu, vit_cpu = cpu_port.score_images_cpu(frames, sd, torch.from_numpy(tags))
        dt = time.perf_counter() - t0
        cpu_baseline = {"value": len(frames) / dt, "unit": "images/s", "cores": cores, "kind": "port",
                        "sample": f"{len(frames)} of the step's 24 MP frames through oracle/cpu_port.py "
                                  "(cv2+NumPy+SciPy technical metrics, PIL preprocess, fp32 torch-CPU ViT-L/14)"}
        # parity spot check on those frames (checker only)
        got = scorer.score_images(np.stack(frames))
        cosv = [float(np.dot(np.frombuffer(g["clip_embedding"], np.float32), vit_cpu["embedding"][i].numpy())) for i, g in enumerate(got)]
        cpu_baseline["parity_spot_check"] = {
            "hist_exact": all(g["histogram_data"] == t["histogram"]["histogram_bytes"] for g, t in zip(got, tech_cpu)),
            "min_embedding_cosine": min(cosv)}

    if rank == 0:
        line = {
            "metric": "images/sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp16", "data": "synthetic", "config": workload_config(B),
            "clocks": clocks, "gpu_launches": launches, "e2e": e2e, "roofline": roofline, "cpu_baseline": cpu_baseline,
            "stages": stages, "pairs_last_step": int(pairs.shape[0]),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
