#!/usr/bin/env python
"""Headline benchmark (BASELINE.json): Act3D keyframes/s on the C2 workload
(4 views 256x256 RGB-D, 16384 ghost points/level x 3 levels, batch 16 per GPU, E=60, H=4),
plus ChainedDiffuser denoise-steps/s (C3) as a secondary figure in the same JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One process per GPU (torchrun for N>1; inference shards keyframes across ranks with no data-path
collective => "weak" scaling, aggregate = sum over ranks / max time).  Timing: CUDA events per step
on the launching stream, L2 flushed between steps, max over ranks; `e2e` repeats the measurement
with pinned HOST inputs (H2D inside the timed region) and a D2H read of the predicted action.
`--impl reference` times the CPU oracle port of the reference (oracle/, same workload, bounded
sample) on the host cores of rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOAD = dict(name="act3d_C2", batch=16, ncam=4, hw=256, ghost_per_level=16384, levels=3, embed=60, heads=4,
                use_instruction=True)
TRAIN_WORKLOAD = dict(name="act3d_C4_train", batch=16, ncam=4, ghost_total=1000, embed=60, heads=4)
PLANNER_WORKLOAD = dict(name="planner_C3", batch=32, ncam=4, hw=256, length=50, steps=100, embed=120, heads=8)


CONFIG_WORKLOAD = ("Act3D forward C2: 4 views 256x256 RGB-D, 16384 ghost pts/level x 3 levels, batch 16/GPU, "
                   "E=60 H=4, use_instruction=1, backbone=resnet50 (random init, cuDNN)")


def env_int(name, default):
    return int(os.environ.get(name, default))


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm_gbs=d["hbm_gbs"], tflops_burst=d["bf16_tflops"], tflops_sustained=d["bf16_tflops_sustained"],
                    source="measured")
    return dict(hbm_gbs=6650.0, tflops_burst=1590.0, tflops_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        sm, reasons, mx = [], set(), 0.0
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = max(mx, float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ inputs
def act3d_inputs(batch, ncam, seed):
    from tests.golden import synth
    g = torch.Generator().manual_seed(seed)
    lo, hi = torch.tensor(synth.WORKSPACE_LO), torch.tensor(synth.WORKSPACE_HI)
    rgb = torch.rand(batch, ncam, 3, 256, 256, generator=g)
    pcd = (lo + torch.rand(batch, ncam, 256, 256, 3, generator=g) * (hi - lo)).permute(0, 1, 4, 2, 3).contiguous()
    instr = torch.randn(batch, 53, 512, generator=g)
    quat = torch.randn(batch, 4, generator=g)
    grip = torch.cat([lo + torch.rand(batch, 3, generator=g) * (hi - lo), quat / quat.norm(dim=-1, keepdim=True),
                      (torch.rand(batch, 1, generator=g) > 0.5).float()], -1)
    return rgb, pcd, instr, grip


def build_act3d():
    from tests.golden import synth
    from model import Act3D
    torch.manual_seed(0)
    w = WORKLOAD
    m = Act3D(backbone="resnet", image_size=(256, 256), embedding_dim=w["embed"], num_attn_heads=w["heads"],
              gripper_loc_bounds=synth.BOUNDS, num_ghost_points_val=w["ghost_per_level"] * w["levels"],
              num_sampling_level=w["levels"], use_instruction=w["use_instruction"])
    return m.eval()


def planner_inputs(batch, ncam, length, seed):
    from tests.golden import synth
    g = torch.Generator().manual_seed(seed)
    lo, hi = torch.tensor(synth.WORKSPACE_LO), torch.tensor(synth.WORKSPACE_HI)
    rgb = torch.rand(batch, ncam, 3, 256, 256, generator=g)
    pcd = (lo + torch.rand(batch, ncam, 256, 256, 3, generator=g) * (hi - lo)).permute(0, 1, 4, 2, 3).contiguous()
    instr = torch.randn(batch, 53, 512, generator=g)

    def pose():
        q = torch.randn(batch, 4, generator=g)
        return torch.cat([lo + torch.rand(batch, 3, generator=g) * (hi - lo), q / q.norm(dim=-1, keepdim=True)], -1)
    return torch.zeros(batch, length, dtype=torch.bool), rgb, pcd, instr, pose(), pose()


def build_planner():
    from tests.golden import synth
    from model import DiffusionPlanner
    torch.manual_seed(0)
    w = PLANNER_WORKLOAD
    m = DiffusionPlanner(backbone="resnet", image_size=(256, 256), embedding_dim=w["embed"], output_dim=7,
                         num_vis_ins_attn_layers=2, num_query_cross_attn_layers=6, use_instruction=True, use_goal=True,
                         use_goal_at_test=False, weight_tying=True, gripper_loc_bounds=synth.BOUNDS,
                         rotation_parametrization="6D", diffusion_timesteps=w["steps"])
    return m.eval()


def planner_step_fn(device):
    w = PLANNER_WORKLOAD
    m = build_planner().to(device)
    ins = [t.to(device) for t in planner_inputs(w["batch"], w["ncam"], w["length"], 5)]
    return lambda: m.compute_trajectory(*ins)


def run_planner(rank, world, device, iters=5):
    """Secondary figure: denoise-steps/s = B * 100 / time of one compute_trajectory (incl. the one-off context encode).
    Median of `iters` individually timed calls after two warm-up calls (cuDNN autotuning, buffer allocation)."""
    w = PLANNER_WORKLOAD
    fn = planner_step_fn(device)
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    times = []
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        times.append(s.elapsed_time(e))
    from act3d_chained_diffuser_b200.sharding import max_over_ranks
    ms = max_over_ranks(sorted(times)[len(times) // 2], device)
    return {"metric": "denoise-steps/s", "value": round(w["batch"] * w["steps"] * world / (ms * 1e-3), 1),
            "ms_per_trajectory_batch": round(ms, 3), "ms_all_calls": [round(t, 2) for t in times],
            "workload": "ChainedDiffuser compute_trajectory C3: batch 32/GPU, 50 waypoints, 100 DDPM steps, 4 views, E=120 H=8"}


def _ddp(model, local, world):
    """Stock DistributedDataParallel exactly as the reference wraps its models (engine.py:121-124:
    find_unused_parameters=True because 6 FPN tensors get no gradient; broadcast_buffers=False), plus
    gradient_as_bucket_view (gradients live in the all-reduce buckets: no copy in / out)."""
    if world == 1:
        return model
    return torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], broadcast_buffers=False,
                                                     find_unused_parameters=True, gradient_as_bucket_view=True)


def _local_mean_gradients(model, step_loss, reseed, world):
    """Gradients of this rank's own shard, computed on the bare module BEFORE it is wrapped in DDP, averaged over the
    ranks with a plain all-reduce: what DDP's bucketed all-reduce must reproduce."""
    params = [p for p in model.parameters() if p.requires_grad]
    reseed()
    model.zero_grad(set_to_none=True)
    step_loss(model).backward()
    mean = [torch.zeros_like(p) if p.grad is None else p.grad.detach().clone() for p in params]
    for g in mean:
        torch.distributed.all_reduce(g)
        g /= world
    model.zero_grad(set_to_none=True)
    return mean


def _ddp_gradient_check(model, net, mean, step_loss, reseed):
    """Once per run under N > 1: the gradients DDP leaves on every rank equal the all-reduced mean of the per-rank
    gradients (NCCL all-reduce of gradients only) and are identical on all ranks."""
    params = [p for p in model.parameters() if p.requires_grad]
    reseed()
    step_loss(net).backward()                                    # through DDP: bucketed all-reduce
    worst = 0.0
    for p, want in zip(params, mean):
        got = torch.zeros_like(p) if p.grad is None else p.grad.detach()
        lo, hi = got.clone(), got.clone()
        torch.distributed.all_reduce(lo, op=torch.distributed.ReduceOp.MIN)
        torch.distributed.all_reduce(hi, op=torch.distributed.ReduceOp.MAX)
        assert torch.equal(lo, hi), "DDP left different gradients on different ranks"
        denom = want.abs().max().item() + 1e-12
        worst = max(worst, (got - want).abs().max().item() / denom)
    assert worst <= 1e-4, f"DDP gradients differ from the all-reduced mean of the per-rank gradients: {worst:.2e}"
    return worst


def _time_train(step, world, device, iters, warmup):
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        loss = step()
    e.record()
    torch.cuda.synchronize()
    from act3d_chained_diffuser_b200.sharding import max_over_ranks
    return max_over_ranks(s.elapsed_time(e) / iters, device), float(loss.detach())


def run_train(rank, world, device, local, iters=6, warmup=3):
    """Secondary figure: Act3D training step (forward + keypose loss + backward + AdamW) under stock DDP
    (gradient all-reduce over NCCL only, engine.py:121-124), synthetic batch of TRAIN_WORKLOAD per GPU."""
    from tests.golden import synth
    from model import Act3D
    from act3d_chained_diffuser_b200.losses import keypose_loss
    w = TRAIN_WORKLOAD
    torch.manual_seed(0)
    model = Act3D(backbone="resnet", image_size=(256, 256), embedding_dim=w["embed"], num_attn_heads=w["heads"],
                  gripper_loc_bounds=synth.BOUNDS, num_ghost_points=w["ghost_total"], num_sampling_level=3,
                  use_instruction=True).to(device).train()
    model.seed_ghost_sampler(99 + rank)
    rgb, pcd, instr, grip = [t.to(device) for t in act3d_inputs(w["batch"], w["ncam"], seed=300 + rank)]
    gt = grip.clone()
    gt[:, :3] += 0.02

    def batch_loss(m, rgb, pcd, instr, grip, gt):
        out = m(rgb, pcd, instr, grip, gt_action=gt)
        return sum(keypose_loss(out, gt).values())

    step_loss = lambda m: batch_loss(m, rgb, pcd, instr, grip, gt)
    from act3d_chained_diffuser_b200.train_graph import GraphedTrainStep, build_ddp, freeze_parameters_without_gradient
    reseed = lambda: model.seed_ghost_sampler(99 + rank)
    # the six FPN output blocks that never reach the loss stop requiring gradients, so DDP runs without
    # find_unused_parameters (no per-step graph traversal / used-parameter all-reduce) and the step can be captured
    frozen = freeze_parameters_without_gradient(model, step_loss)
    mean = _local_mean_gradients(model, step_loss, reseed, world) if world > 1 else None
    net = model if world == 1 else build_ddp(model, device_ids=[local], broadcast_buffers=False,
                                             gradient_as_bucket_view=True)
    opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-4)

    def step():
        loss = step_loss(net)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    check = _ddp_gradient_check(model, net, mean, step_loss, reseed) if world > 1 else None
    ms_eager, _ = _time_train(step, world, device, iters, warmup)
    # the whole step (forward, loss, backward, bucketed all-reduce, AdamW) as one CUDA graph
    graph_note = "whole step replayed from a CUDA graph"
    try:
        graphed = GraphedTrainStep(model, batch_loss, lambda ps: torch.optim.AdamW(ps, lr=1e-4, capturable=True),
                                   (rgb, pcd, instr, grip, gt), net=net)
        ms, loss = _time_train(lambda: graphed(rgb, pcd, instr, grip, gt), world, device, max(iters, 20), warmup)
    except Exception as exc:       # a capture that fails on some NCCL / driver combination must not take the bench line with it
        graph_note = f"eager launches (whole-step capture failed: {type(exc).__name__}: {str(exc)[:120]})"
        torch.cuda.synchronize()
        ms, loss = _time_train(step, world, device, iters, warmup)
    out = {"metric": "train keyframes/s", "value": round(w["batch"] * world / (ms * 1e-3), 1), "ms_per_step": round(ms, 3),
           "eager_launch_ms_per_step": round(ms_eager, 3), "final_loss": round(loss, 4),
           "parallelism": f"DDP x{world} (NCCL gradient all-reduce only, inside the captured step)",
           "frozen_unreachable_parameters": len(frozen),
           "workload": f"Act3D training step: {w['batch']} keyframes/GPU, 4 views 256x256, {w['ghost_total']} ghost points "
                       f"(333/level), use_instruction=1, frozen ResNet-50, fp32, AdamW; {graph_note}"}
    if check is not None:
        out["ddp_gradient_check_max_rel"] = float(f"{check:.2e}")
    return out


def run_train_planner(rank, world, device, local, iters=6, warmup=3):
    """Secondary figure: ChainedDiffuser training step (main_trajectory.py:177-199: noise one random timestep, one
    denoiser call, L1 losses, backward, AdamW) under stock DDP: the 17 MB gradient all-reduce of SURVEY 8(e)."""
    w = PLANNER_WORKLOAD
    torch.manual_seed(0)
    model = build_planner().to(device).train()
    mask, rgb, pcd, instr, cur, goal = [t.to(device) for t in planner_inputs(w["batch"], w["ncam"], w["length"], 500 + rank)]
    g = torch.Generator().manual_seed(700 + rank)
    alpha = torch.linspace(0, 1, w["length"]).view(1, -1, 1)
    gt_traj = (cur.cpu().unsqueeze(1) * (1 - alpha) + goal.cpu().unsqueeze(1) * alpha + 0.01 * torch.randn(w["batch"], w["length"], 7, generator=g))
    gt_traj[..., 3:] = gt_traj[..., 3:] / gt_traj[..., 3:].norm(dim=-1, keepdim=True)
    gt_traj = gt_traj.to(device)

    def step_loss(m):
        return m(gt_traj, mask, rgb, pcd, instr, cur, goal)

    reseed = lambda: torch.manual_seed(1234 + rank)
    mean = _local_mean_gradients(model, step_loss, reseed, world) if world > 1 else None
    net = _ddp(model, local, world)
    opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-4)

    def step():
        loss = step_loss(net)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    check = _ddp_gradient_check(model, net, mean, step_loss, reseed) if world > 1 else None
    ms, loss = _time_train(step, world, device, iters, warmup)
    nbytes = sum(p.numel() * 4 for p in model.parameters() if p.requires_grad)
    out = {"metric": "train trajectories/s", "value": round(w["batch"] * world / (ms * 1e-3), 1), "ms_per_step": round(ms, 3),
           "final_loss": round(loss, 4), "parallelism": f"DDP x{world} (NCCL gradient all-reduce only)",
           "gradient_bytes": int(nbytes),
           "workload": f"ChainedDiffuser training step: {w['batch']} trajectories/GPU, {w['length']} waypoints, 4 views 256x256, "
                       "E=120 H=8, frozen ResNet-50, fp32, AdamW"}
    if check is not None:
        out["ddp_gradient_check_max_rel"] = float(f"{check:.2e}")
    return out


# ------------------------------------------------------------------------------------------ our arm
def run_ours(args, rank, world, device, local=0):
    from act3d_chained_diffuser_b200 import lib
    lib.load()
    w = WORKLOAD
    model = build_act3d().to(device)
    model.seed_ghost_sampler(1234 + rank)
    host = [t.pin_memory() for t in act3d_inputs(w["batch"], w["ncam"], seed=100 + rank)]
    dev_in = [t.to(device) for t in host]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)          # > 126 MB L2

    def step_resident(inputs=None):
        with torch.no_grad():
            return model(*(inputs or dev_in))

    def step_e2e():
        with torch.no_grad():
            out = model(*host)          # pinned host tensors: the module uploads them (images first, rest overlapped)
            return torch.cat([out["position"], out["rotation"].reshape(len(host[0]), -1), out["gripper"]], -1).cpu()

    def timed(fn, steps, warmup, profile=False):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        lib.reset_launch_count()
        model._profile_events = [] if profile else None
        evs = []
        for _ in range(steps):
            flush.fill_(1)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            evs.append((s, e))
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        total_ms = sum(s.elapsed_time(e) for s, e in evs)
        launches = lib.launch_count()                  # kernel-launching C-ABI calls (ours) in the timed region
        prof = model._profile_events
        model._profile_events = None
        return total_ms, launches, prof

    clocks = ClockSampler(torch.cuda.current_device() if device.index is None else device.index)
    clocks.start()
    # resident inputs: the forward is replayed from a CUDA graph (model.use_cuda_graph: no host launch cost)
    model.use_cuda_graph = True
    ms_res, _, _ = timed(step_resident, args.steps, args.warmup)
    clk = clocks.stop()
    # the same forward launched eagerly for a few steps: counts our launches per step and times the ghost kernel with events
    model.use_cuda_graph = False
    k_prof = min(args.steps, 10)
    _, launches_prof, prof = timed(step_resident, k_prof, 2, profile=True)
    launches = launches_prof // k_prof * args.steps
    # end to end: pinned host inputs handed to model(...), uploaded inside the timed region (images first, then the trunk
    # graph starts while the point clouds / instruction / gripper cross PCIe on the copy stream), result read back
    model.use_cuda_graph = True
    ms_e2e, _, _ = timed(step_e2e, args.steps, max(2, args.warmup // 2))

    # ---- strong scaling: the SAME 16 keyframes split over the N ranks (16 / N per GPU), no data-path collective
    strong = None
    if world > 1 and w["batch"] % world == 0:
        from act3d_chained_diffuser_b200.sharding import shard_bounds
        lo, hi = shard_bounds(w["batch"], rank, world)
        per = hi - lo
        shard = [t[lo:hi].contiguous() for t in dev_in]
        k_strong = max(5, args.steps // 2)
        from act3d_chained_diffuser_b200.sharding import max_over_ranks
        ms_strong = max_over_ranks(timed(lambda: step_resident(shard), k_strong, 3)[0], device)
        strong = {"metric": "keyframes/s", "scaling": "strong", "value": round(w["batch"] * k_strong / (ms_strong * 1e-3), 1),
                  "ms_per_step": round(ms_strong / k_strong, 3), "keyframes_total": w["batch"], "keyframes_per_gpu": per,
                  "note": "16 keyframes in total, sharded over the ranks; compare with the N=1 value of the same line format"}

    from act3d_chained_diffuser_b200.sharding import max_over_ranks
    ms_res, ms_e2e = max_over_ranks(ms_res, device), max_over_ranks(ms_e2e, device)

    # ---- roofline of the dominant kernel: fused ghost-point cross-attention stack
    peaks = measured_peaks()
    nk = 32 * 32 * w["ncam"] + 1 + (53 if w["use_instruction"] else 0)
    flops_launch = 4.0 * w["ghost_per_level"] * nk * w["embed"] * 2 * w["batch"]      # 2 layers, B samples / launch
    kern_ms = [s.elapsed_time(e) for tag, s, e in (prof or []) if tag == "ghost_xattn"]
    roof = None
    if kern_ms:
        avg = sum(kern_ms) / len(kern_ms)
        ach = flops_launch / (avg * 1e-3) / 1e12
        # traffic = dram__bytes_read + write of one launch from the committed ncu capture (37.6 MB + 19.9 MB; it cannot be
        # measured outside a profiler): the K/V tile images stay in L2, DRAM sees the first touch of K/V plus the part of
        # the ghost features that spills past L2 -- nowhere near the HBM roofline.
        roof = {"kernel": "xattn6_kernel (fused ghost-point cross-attention stack: tcgen05/TMEM attention + linear layers, "
                          "FMA-pipe polynomial for 6 of 16 exponentials)",
                "bound": "tensor", "achieved": round(ach, 2),
                "peak": peaks["tflops_sustained"], "unit": "TFLOP/s", "frac": round(ach / peaks["tflops_sustained"], 4),
                "traffic": 53.8e6, "traffic_source": "profiles/r2_xattn6_ncu.txt (ncu --set full, one C2 launch)", "avg_launch_ms": round(avg, 4), "launches_timed": len(kern_ms),
                "share_of_step": round(avg * w["levels"] / (ms_res / args.steps), 4), "peak_source": peaks["source"] + " bf16 sustained",
                "algorithmic_flops_per_launch": flops_launch,
                "binding_unit": {"unit": "MUFU ex2 (16 / clk / SM at 1.965 GHz, measured 15.9 with tools/micro/mufu_bench.cu)",
                                 "floor_ms": round(w["batch"] * w["ghost_per_level"] * nk * w["heads"] * 2 / (148 * 16 * 1.965e9) * 1e3, 3),
                                 "frac": round(w["batch"] * w["ghost_per_level"] * nk * w["heads"] * 2 / (148 * 16 * 1.965e9) * 1e3 / avg, 4)},
                "note": "true-E attention FLOPs 4*Nq*Nk*E per layer-sample.  head_dim 15 means one exp per 30 useful FLOPs: "
                        "the binding unit is MUFU (16 ex2/clk/SM, measured with tools/micro/mufu_bench.cu), not the tensor pipe; "
                        "binding_unit.floor_ms is the launch time with every exponential on the XU at 100 %; the kernel moves "
                        "6 of 16 to the FMA pipe (ncu: XU 59 %, ALU 34 %, FMA 21 %, tensor 9 %; profiles/r2_xattn6_ncu.txt)"}

    kf = w["batch"] * world * args.steps
    line = {
        "metric": "keyframes/s", "value": round(kf / (ms_res * 1e-3), 3), "unit": "keyframes/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_res / args.steps, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32 (fp16 tensor-core operands, fp32 accumulate)",
        "data": "synthetic", "impl": "ours",
        "config": {"workload": CONFIG_WORKLOAD, "batch_per_gpu": w["batch"], "l2": "flushed between steps (256 MiB write)",
                   "parallelism": f"replicas x{world} (no data-path collective)",
                   "launch": "CUDA-graph replay (trunk graph + attention graph); e2e: images uploaded first, the other inputs on a copy stream under the trunk graph"},
        "e2e": {"value": round(kf / (ms_e2e * 1e-3), 3), "unit": "keyframes/s",
                "h2d_bytes_per_step": int(sum(t.numel() * t.element_size() for t in host)),
                "d2h_bytes_per_step": int(w["batch"] * 8 * 4), "ms_per_step": round(ms_e2e / args.steps, 3)},
        "gpu_launches": int(launches), "gpu_launches_note": "kernels of libact3d_b200.so per timed region (counted on eager launches of the "
                                                             "same forward; the resident measurement replays them from a CUDA graph)",
        "clocks": clk,
        "roofline": roof,
    }
    also = {}
    if strong is not None:
        also["strong_scaling"] = strong
    if not args.no_planner:
        line["secondary"] = run_planner(rank, world, device)
        also["planner_denoise_steps_per_s"] = line["secondary"]["value"]
    if not args.no_train:
        line["secondary_train"] = run_train(rank, world, device, local)
        line["secondary_train_planner"] = run_train_planner(rank, world, device, local)
        also["train_keyframes_per_s_ddp"] = line["secondary_train"]["value"]
        also["train_trajectories_per_s_ddp"] = line["secondary_train_planner"]["value"]
        if "ddp_gradient_check_max_rel" in line["secondary_train"]:
            also["ddp_gradient_check_max_rel"] = max(line["secondary_train"]["ddp_gradient_check_max_rel"],
                                                     line["secondary_train_planner"]["ddp_gradient_check_max_rel"])
    line["config"]["also_measured"] = also          # the driver keeps `config`: secondary figures travel with it
    return line


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_act3d_keyframes_per_s(steps, warmup, threads, budget_s=150.0):
    """The reference's CPU implementation restated (oracle port), B=1 keyframe of the C2 workload per step
    (the reference materialises a 1 GB score tensor per sample-layer; it is batch-linear, so B=1 is the bounded sample).
    Warm-up steps beyond the first are dropped when one step takes so long that the run would not end in minutes."""
    from oracle import act3d_ref
    from tests.golden import synth
    torch.set_num_threads(threads)
    w = WORKLOAD
    model = build_act3d()
    sd = model.state_dict()
    cfg = act3d_ref.Act3DConfig(gripper_loc_bounds=synth.BOUNDS, use_instruction=w["use_instruction"],
                                ghost_points_per_level=w["ghost_per_level"])
    rgb, pcd, instr, grip = act3d_inputs(1, w["ncam"], seed=7)
    trunk = act3d_ref.trunk_from_module(model)
    import numpy as np
    np.random.seed(0)
    times, warm_done, t_begin = [], 0, time.perf_counter()
    with torch.no_grad():
        while len(times) < steps:
            t0 = time.perf_counter()
            act3d_ref.act3d_forward(sd, cfg, trunk, rgb, pcd, instr, grip)
            dt = time.perf_counter() - t0
            if warm_done < warmup and (warm_done == 0 or (time.perf_counter() - t_begin) + (steps + 1) * dt < budget_s):
                warm_done += 1
                continue
            times.append(dt)
    return len(times) / sum(times), sum(times) / len(times), warm_done


def gpu_eager_reference_keyframes_per_s(device, steps=3):
    """The same oracle port (plain eager PyTorch, fp32, materialised score tensors, host ghost sampler: the reference's
    own execution model, act3d.py:176-357) on THIS GPU at the benchmark's batch: the stand-in for "the reference
    single-GPU PyTorch keyframes/sec at batch 16" of north_star (the reference itself cannot travel to the GPU box).
    Falls back to smaller batches if the 17 GB-per-layer score tensors do not fit."""
    from oracle import act3d_ref
    from tests.golden import synth
    import numpy as np
    w = WORKLOAD
    model = build_act3d().to(device)
    sd = {k: v for k, v in model.state_dict().items()}
    cfg = act3d_ref.Act3DConfig(gripper_loc_bounds=synth.BOUNDS, use_instruction=w["use_instruction"],
                                ghost_points_per_level=w["ghost_per_level"])
    trunk = act3d_ref.trunk_from_module(model)
    batch = w["batch"]
    while batch >= 1:
        try:
            ins = [t.to(device) for t in act3d_inputs(batch, w["ncam"], seed=7)]
            np.random.seed(0)
            with torch.no_grad():
                act3d_ref.act3d_forward(sd, cfg, trunk, *ins)              # warm-up (cuDNN autotune, allocator)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(steps):
                    out = act3d_ref.act3d_forward(sd, cfg, trunk, *ins)
                    out["position"].cpu()
                torch.cuda.synchronize()
                dt = (time.perf_counter() - t0) / steps
            peak = torch.cuda.max_memory_allocated(device) / 2 ** 30
            del ins, out
            torch.cuda.empty_cache()
            return {"value": round(batch / dt, 2), "unit": "keyframes/s", "batch": batch, "ms_per_step": round(dt * 1e3, 2),
                    "steps": steps, "peak_memory_gib": round(peak, 1),
                    "what": "oracle port of the reference forward, eager PyTorch fp32 on this GPU (kind: port)"}
        except torch.cuda.OutOfMemoryError:
            torch.cuda.empty_cache()
            batch //= 2
    return {"value": None, "what": "eager PyTorch port ran out of memory even at batch 1"}


def run_reference(args, rank, world):
    """`--impl reference`: the reference's own execution model on the host cores (oracle port; a Python reference cannot
    travel to the GPU box), same metric / unit / config / steps as our arm; every step is a bounded sample of the
    16-keyframe batch: one keyframe (the model is batch-linear: no cross-sample operation)."""
    if rank != 0:
        return None
    threads = os.cpu_count() or 1
    kfs, sec, warm = cpu_act3d_keyframes_per_s(args.steps, args.warmup, threads)
    w = WORKLOAD
    return {
        "metric": "keyframes/s", "value": round(kfs, 4), "unit": "keyframes/s", "impl": "reference",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(sec * 1e3, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": CONFIG_WORKLOAD, "batch_per_gpu": w["batch"], "l2": "flushed between steps (256 MiB write)",
                   "parallelism": f"replicas x{world} (no data-path collective)"},
        "cpu_baseline": {"value": round(kfs, 4), "unit": "keyframes/s", "cores": threads, "kind": "port",
                         "sample": f"{args.steps} steps of 1 keyframe each (of the 16-keyframe batch; batch-linear model) after {warm} "
                                   f"warm-up step(s), {sec:.2f} s per keyframe; arithmetic in plain fp32"},
        "e2e": {"value": round(kfs, 4), "unit": "keyframes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-planner", action="store_true")
    ap.add_argument("--no-train", action="store_true")
    args = ap.parse_args()
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)

    if args.impl == "reference":
        line = run_reference(args, rank, world)
        if line is not None:
            print(json.dumps(line), flush=True)
        return

    assert torch.cuda.is_available(), "bench.py --impl ours needs a GPU (there is no CPU fallback)"
    torch.backends.cudnn.benchmark = True          # as the reference's entry points do (main_keypose.py:518-520)
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")     # whole-step CUDA-graph capture of DDP (train_graph.py)
        torch.distributed.init_process_group("nccl", device_id=device)
    args.warmup = max(args.warmup, 3)
    line = run_ours(args, rank, world, device, local)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            ref_gpu = gpu_eager_reference_keyframes_per_s(device)
            kfs, sec, warm = cpu_act3d_keyframes_per_s(6, 1, threads, budget_s=60.0)
            line["cpu_baseline"] = {"value": round(kfs, 4), "unit": "keyframes/s", "cores": threads, "kind": "port",
                                    "sample": f"6 keyframes (B=1 per step) of the same workload after {warm} warm-up, {sec:.2f} s each",
                                    "ref_gpu": ref_gpu}
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
