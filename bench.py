#!/usr/bin/env python
"""bench.py -- SMoE hot-path benchmark (contract in the task statement / DESIGN.md section "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c1|c2|c3|c4s|c4]

One "step" = one full training iteration of the hot path on the workload: pi-mask compaction,
fused forward (both sweeps), fused backward, statistics -> gradients, Adam, kernel-list upkeep.
Workload (default): BASELINE config 3 -- 1920x1080 RGB, 128x256 kernel grid (32,768 kernels), the
configuration the north-star target is quoted on; it fits one GPU and is the one that shards over
ranks (rows), so every N runs the same total work ("strong" scaling).
metric = pixel.kernel evaluations/s through forward+backward = N_pixels * K_active * steps / time.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (shape, kernels_per_dim, seed, description)
    "c1": ((128, 128, 1), [16, 16], 1001, "128x128x1, 16x16 kernels (config 1)"),
    "c2": ((512, 512, 1), [64, 64], 1002, "512x512x1, 64x64 kernels (config 2)"),
    "c3": ((1080, 1920, 3), [128, 256], 1003, "1920x1080 RGB, 128x256=32768 kernels (config 3)"),
    "c4s": ((360, 640, 16, 3), [16, 32, 16], 1004, "640x360x16 RGB video, 16x32x16=8192 kernels, 3x3 A (config 4 at 1/8 scale)"),
    "c4": ((720, 1280, 32, 3), [32, 64, 32], 1004, "1280x720x32 RGB video, 32x64x32=65536 kernels, 3x3 A (config 4; meant for 2/4/8 GPUs)"),
}
SMOE_KW = dict(use_determinant=True, train_inverse_cov=False, use_yuv=False, normalize_pis=True)


def synth_image(shape, seed):
    """SURVEY.md 8d synthetic pattern, uint8-quantised then /255 (utils.py:126-128)."""
    rs = np.random.RandomState(seed)
    H, W, C = shape[0], shape[1], shape[-1]
    v, u = np.meshgrid(np.linspace(0, 1, H, dtype=np.float32), np.linspace(0, 1, W, dtype=np.float32), indexing="ij")

    def base(uu, vv):
        return (0.5 + 0.25 * np.sin(2 * np.pi * (3 * uu + 2 * vv)) + 0.2 * (uu > vv)
                + 0.15 * np.exp(-((uu - .3) ** 2 + (vv - .6) ** 2) / 0.02)).astype(np.float32)
    shifts = [0.0, 0.11, 0.23]
    if len(shape) == 3:
        img = np.stack([base(u + shifts[c], v) for c in range(C)], axis=-1)
    else:
        ts = np.linspace(0, 1, shape[2])
        img = np.stack([np.stack([base(u + shifts[c] + 0.1 * t, v) for c in range(C)], axis=-1) for t in ts], axis=2)
    img = np.clip(img + 0.03 * rs.standard_normal(img.shape).astype(np.float32), 0, 1)
    return (np.round(img * 255).astype(np.uint8).astype(np.float32) / 255.).astype(np.float32)


def algorithmic_lane_instr(d, C):
    """FP32 lane-instructions per evaluation (FMA counted once)."""
    T = d * (d + 1) // 2
    f_fwd = 2 * d + T + 2 + C * (d + 1)
    f_bwd = 4 * d + 2 * T + 6 + 2 * d * C + 3 * C
    return f_fwd, f_bwd


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fd:
            j = json.load(fd)
        return j, "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# CPU arm: the oracle's restatement of the reference graph on the host cores (TensorFlow is not
# installable offline).  Same materialisation strategy as the reference: broadcast
# (K,N,d,d,d) einsum product, explicit K x N gate matrix, autograd backward, TF-style Adam.
# ----------------------------------------------------------------------------------------------
def cpu_reference_evals_per_s(workload, steps, warmup, budget_s=25.0):
    import torch
    from oracle.model import OracleAdam, OracleSmoe
    shape, kgrid, seed, _ = WORKLOADS[workload]
    d = len(shape) - 1
    # bounded sample of the same workload: a crop of the image with the kernels of the matching
    # crop of the grid (same pixels-per-kernel density), sized for ~1-2 s per step on 8 cores
    frac = {"c1": 1, "c2": 4, "c3": 8, "c4s": 8, "c4": 16}[workload]
    crop = tuple(max(s // frac, 8) for s in shape[:d]) + (shape[-1],)
    kcrop = [max(k // frac, 2) for k in kgrid]
    img = synth_image(shape, seed)[tuple(slice(0, c) for c in crop[:d])]
    torch.set_num_threads(os.cpu_count() or 1)
    m = OracleSmoe(img, kernels_per_dim=kcrop, dtype=torch.float32, einsum_mode="broadcast", **SMOE_KW)
    m.set_optimizer(OracleAdam(1e-3), OracleAdam(1e-5), OracleAdam(1.0))
    evals = int(np.prod(crop[:d])) * int(np.prod(kcrop))
    for _ in range(warmup):
        m.run_batched(train=True)
    times = []
    t_start = time.perf_counter()
    for _ in range(steps):
        t0 = time.perf_counter()
        m.run_batched(train=True)
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > budget_s:
            break
    t = float(np.median(times))
    return {"value": evals / t, "unit": "evals/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"crop {crop[:d]} px x {int(np.prod(kcrop))} kernels of {workload} ({evals:.3g} evals/step), "
                      f"median of {len(times)} steps, PyTorch-CPU float32, reference graph restated "
                      f"(TensorFlow unavailable)", "ms_per_step": t * 1e3, "steps": len(times)}, evals


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    shape, kgrid, _, desc = WORKLOADS[args.workload]
    base, evals = cpu_reference_evals_per_s(args.workload, max(args.steps, 3), max(args.warmup, 1), budget_s=120.0)
    line = {"impl": "reference", "metric": "pixel_kernel_evals_per_s_fwd_bwd", "value": base["value"], "unit": "evals/s",
            "n_gpus": args.gpus, "steps": base["steps"], "warmup": max(args.warmup, 1), "ms_per_step": base["ms_per_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "sampled": base["sample"]},
            "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": base["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def run_ours(args):
    import torch
    import __graft_entry__ as ge
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        ge.build()
    if world > 1:
        torch.distributed.barrier()
    from smoe_b200 import Smoe, AdamOptimizer

    shape, kgrid, seed, desc = WORKLOADS[args.workload]
    d, C = len(shape) - 1, shape[-1]
    img = synth_image(shape, seed)
    m = Smoe(img, kernels_per_dim=kgrid, dense_exec=int(args.dense_exec), **SMOE_KW)
    m.set_optimizer(AdamOptimizer(1e-3), AdamOptimizer(1e-5), AdamOptimizer(1.0))
    N, K = m.num_pixel, m.start_pis
    evals_per_step = float(N) * float(K)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed_steps(n, host_image=None):
        """n steps, each bracketed by its own CUDA events on the launching stream, L2 flushed
        (256 MB write) between steps outside the events; returns per-step ms (max over ranks)."""
        ms = []
        for _ in range(n):
            flush.fill_(1.0)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            m.run_batched(train=True, _host_image=host_image)
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
            if world > 1:
                torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms.append(float(t.item()))
        return ms

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    timed_steps(max(args.warmup, 3))
    if rank == 0:
        time.sleep(0.3)
        sampler.rows.clear()          # keep only the samples taken from here on
    l0 = m.gpu_launches
    ms = timed_steps(args.steps)
    launches = (m.gpu_launches - l0) // max(args.steps, 1)
    n_timed_samples = len(sampler.rows) if rank == 0 else 0
    total_s = sum(ms) / 1e3
    value = evals_per_step * args.steps / total_s

    # e2e: same steps through the public API with this step's pixels coming from pinned host memory -- the 8-bit
    # pixels as an image file holds them; the /255 conversion (utils.py:126-128) runs on the device
    b0, b1 = m._band
    u8 = np.round(img[b0:b1] * 255).astype(np.uint8)
    assert np.array_equal(u8.astype(np.float32) / 255., img[b0:b1])
    host_img = torch.from_numpy(np.ascontiguousarray(u8)).pin_memory()
    timed_steps(2, host_img)
    ms_e2e = timed_steps(args.steps, host_img)
    e2e_value = evals_per_step * args.steps / (sum(ms_e2e) / 1e3)
    h2d = host_img.numel() * host_img.element_size()
    d2h = m._host_stats.numel() * 4
    # a timed region of a few milliseconds is shorter than nvidia-smi's sampling period: keep the same step
    # running (untimed) until the sampler has seen it, and say so
    clocks = None
    if rank == 0 or world > 1:
        t_probe = time.time()
        probe = 0
        need_probe = torch.tensor([1.0 if (rank == 0 and len(sampler.rows) < 5) else 0.0], device="cuda")
        if world > 1:
            torch.distributed.all_reduce(need_probe, op=torch.distributed.ReduceOp.MAX)
        while need_probe.item() > 0 and time.time() - t_probe < 1.5:
            for _ in range(20):
                m.run_batched(train=True)
            probe += 20
            need_probe = torch.tensor([1.0 if (rank == 0 and len(sampler.rows) < 5 and time.time() - t_probe < 1.5) else 0.0],
                                      device="cuda")
            if world > 1:
                torch.distributed.all_reduce(need_probe, op=torch.distributed.ReduceOp.MAX)
        if rank == 0:
            clocks = sampler.stop()
            clocks["samples_in_timed_region"] = n_timed_samples
            clocks["window"] = "timed steps + e2e steps" + (f" + {probe} untimed steps of the same workload" if probe else "")

    # per-kernel durations of the two sweep kernels (CUDA events on the launching stream)
    ker = kernel_times(m, steps=max(3, min(args.steps, 10)))
    f_fwd, f_bwd = algorithmic_lane_instr(d, C)
    peaks, peak_src = measured_peaks()
    sms = torch.cuda.get_device_properties(local).multi_processor_count
    peak_tflops = sms * 128 * 2 * peaks["sm_max_mhz"] * 1e6 / 1e12
    local_evals = evals_per_step / world
    dom = "backward" if ker["backward_ms"] >= ker["forward_ms"] else "forward"
    f_dom = f_bwd if dom == "backward" else f_fwd
    achieved = local_evals * f_dom * 2 / (ker[dom + "_ms"] / 1e3) / 1e12
    roofline = {"bound": "fp32", "kernel": f"smoe::{dom}_kernel<{d},{C}>", "achieved": achieved, "peak": peak_tflops,
                "unit": "TFLOP/s", "frac": achieved / peak_tflops, "traffic": None,
                "peak_source": f"{sms} SMs x 128 FP32 lanes x 2 x sm_max_mhz={peaks['sm_max_mhz']} ({peak_src} MEASURED_PEAKS.json)",
                "algorithmic_lane_instr_per_eval": {"fwd": f_fwd, "bwd": f_bwd},
                "forward_ms": ker["forward_ms"], "backward_ms": ker["backward_ms"], "other_ms": ker["other_ms"],
                "step_frac_fwd_bwd": (local_evals * (f_fwd + f_bwd) * 2 / (np.mean(ms) / 1e3) / 1e12) / peak_tflops,
                "note": "achieved = ALGORITHMIC lane-instructions of SURVEY 8d (N*K*F) / launch time.  In the default mode "
                        "the kernels cull and skip (pixel,kernel) pairs whose contribution is exactly 0 (bit-identical "
                        "results, asserted by tests), so frac >> 1 means work provably not needed, not work not done; "
                        "'dense_exec' holds the same kernels executing every pair"}
    if clocks and clocks.get("sm_mhz"):
        roofline["frac_at_observed_clock"] = roofline["frac"] * peaks["sm_max_mhz"] / clocks["sm_mhz"]
    roofline["mode"] = {0: "exact culling + exact-zero skipping", 1: "every (pixel,kernel) pair executed",
                        2: "exact-zero skipping only"}[int(args.dense_exec)]
    tr = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if os.path.exists(tr):
        with open(tr) as fd:
            roofline["traffic"] = json.load(fd).get(f"{dom}_kernel<{d},{C}>@{args.workload}@mode{int(args.dense_exec)}")
    # the same two kernels with every (pixel, kernel) pair executed (dense_exec=1): the figure that compares
    # like with like against the algorithmic instruction count
    if world == 1 and int(args.dense_exec) == 0 and not args.no_dense:
        md = Smoe(img, kernels_per_dim=kgrid, dense_exec=1, **SMOE_KW)
        md.set_optimizer(AdamOptimizer(1e-3), AdamOptimizer(1e-5), AdamOptimizer(1.0))
        md.run_batched(train=True)
        kd = kernel_times(md, steps=2, with_step=False)
        domd = "backward" if kd["backward_ms"] >= kd["forward_ms"] else "forward"
        ach = evals_per_step * (f_bwd if domd == "backward" else f_fwd) * 2 / (kd[domd + "_ms"] / 1e3) / 1e12
        roofline["dense_exec"] = {"kernel": f"smoe::{domd}_kernel<{d},{C}>", "forward_ms": kd["forward_ms"],
                                  "backward_ms": kd["backward_ms"], "achieved": ach, "frac": ach / peak_tflops,
                                  "fwd_bwd_frac": (evals_per_step * (f_fwd + f_bwd) * 2 /
                                                   ((kd["forward_ms"] + kd["backward_ms"]) / 1e3) / 1e12) / peak_tflops}
        del md

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu, _ = cpu_reference_evals_per_s(args.workload, 5, 1)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
    if rank == 0:
        line = {"metric": "pixel_kernel_evals_per_s_fwd_bwd", "value": value, "unit": "evals/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": float(np.mean(ms)),
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": desc, "pixels": N, "kernels": K, "d": d, "C": C, "batches": 1,
                           "parallelism": f"rows sharded over {world} rank(s), 1 all-reduce/step" if world > 1 else "1 GPU",
                           "l2": "256 MB flush write between timed steps (outside the events)",
                           "dense_exec": int(args.dense_exec), "init": "regular grid, pi=1/K, use_determinant"},
                "iters_per_s": args.steps / total_s,
                "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": float(np.mean(ms_e2e))},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu}
        print(json.dumps(line))
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


def kernel_times(m, steps, with_step=True):
    """Average launch duration of the forward and backward sweep kernels, CUDA events on the
    launching stream around the C-ABI calls (parameters frozen: no Adam between launches)."""
    import ctypes as C
    import torch
    from smoe_b200._ffi import check, lib, ptr, stream_ptr
    L, st = lib(), stream_ptr()
    b = m._batches[0]
    counts, regs, scal = m._counts[0], m._regsums[0], m._scalars[0]
    ax2 = ptr(m._d_axes[2]) if m.dim_domain == 3 else ptr(None)
    check(L.smoe_pack(C.byref(m._cfg), ptr(m._theta), ptr(m._mus_grid), ptr(None), ptr(m._klist[0]), m.start_pis, ptr(m._packed), ptr(m._indices),
                      ptr(m._pos), ptr(counts), ptr(regs), ptr(m._chunk_bounds), ptr(m._pack_ws), st), "pack")
    fw, bw = [], []
    for _ in range(steps + 1):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        m._infl.zero_()
        scal.zero_()
        e[0].record()
        check(L.smoe_forward(C.byref(m._cfg), C.byref(b), ptr(m._packed), ptr(m._indices), ptr(counts),
                             ptr(m._chunk_bounds), m.start_pis, ptr(m._d_image), ptr(None),
                             ptr(m._d_axes[0]), ptr(m._d_axes[1]), ax2, ptr(m._d_res), ptr(None), ptr(None), ptr(m._infl),
                             ptr(m._pix), ptr(m._tile_qmin), ptr(scal), ptr(m._partials), ptr(m._ticket), st), "forward")
        e[1].record()
        check(L.smoe_backward(C.byref(m._cfg), C.byref(b), ptr(m._packed), ptr(counts), m.start_pis,
                              ptr(m._perm), ptr(m._pos), ptr(m._pix), ptr(m._tile_qmin), ptr(m._d_axes[0]), ptr(m._d_axes[1]), ax2, m._splits, ptr(m._raw_part), st),
              "backward")
        e[2].record()
        torch.cuda.synchronize()
        fw.append(e[0].elapsed_time(e[1]))
        bw.append(e[1].elapsed_time(e[2]))
    fw, bw = fw[1:], bw[1:]
    if not with_step:
        return {"forward_ms": float(np.mean(fw)), "backward_ms": float(np.mean(bw)), "other_ms": 0.0}
    # everything else in a step: pack, finalize, list upkeep, Adam, memsets
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        m.run_batched(train=True)
    e1.record()
    torch.cuda.synchronize()
    step = e0.elapsed_time(e1) / 3
    return {"forward_ms": float(np.mean(fw)), "backward_ms": float(np.mean(bw)),
            "other_ms": max(0.0, step - float(np.mean(fw)) - float(np.mean(bw)))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--dense-exec", type=int, default=0,
                    help="0: exact culling + exact-zero skipping (default); 1: execute every pair; 2: skipping only")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-dense", action="store_true", help="skip the dense_exec=1 roofline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
