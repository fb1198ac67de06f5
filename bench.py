#!/usr/bin/env python
"""bench.py -- NeFeS render-path throughput on B200 (rays/s, forward + backward, 64 + 64 samples).

Workload = BASELINE.json configs[1]: the stage-1 colour-only training step of
run_nefes.py on a 7-Scenes-stairs-shaped synthetic camera (640x480 intrinsics -> 60x80 ray grid,
focal 65.688, near 0, far 4), 4 images x 1536 random rays = 6144 rays per GPU per step:
get_rays_batch -> gather -> render() [stratified sampling, coarse field, compositing, sample_pdf,
fine field with transient heads, compositing] -> NeRF-W colour loss -> backward to all weights ->
(N>1: one all-reduce of the flat gradients) -> Adam.  Weights are the reference constructor's
random init; data is synthetic.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl engine|reference] [--precision fp32|bf16]
                    [--workload train|c3|refine|sweep] [--scaling weak|strong]

One JSON line on stdout (rank 0).  `value` times K steps with inputs resident in HBM; `e2e` times the
same steps through the public API with the step's host inputs copied from pinned memory and the loss
read back every step.  `--impl reference` times the UNMODIFIED reference (staged under oracle/_ref by
oracle/build_ref.py; the oracle port when that copy is absent) on the host cores, same workload.

Workloads: train = C2 (the default and the headline); c3 = the stage-2 step (C2 + the 128-channel
feature L1 term, so the feature cotangent path runs); c3s3 = the stage-3 step (7 patches of 16x16 per image, affine colour
transform, FusionNet, colour + 0.02 feature + 0.02 fusion losses; the fields AND the FusionNet / exposure network train); refine = C4 (50 pose-gradient iterations per query, queries sharded
over ranks, final gather; metric refine iters/s); sweep = C5-shaped inference render of 2^12..2^20 rays (ray ranges sharded).
--scaling strong keeps the GLOBAL batch at 6144 rays (6144 / N per GPU) for train / c3.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, FOCAL, NEAR, FAR = 60, 80, 525.505 / 2 / 4, 0., 4.
N_IMAGES, N_RAND = 4, 1536
RAYS = N_IMAGES * N_RAND
# algorithmic MLP work per ray, SURVEY.md 8d: fwd 68.32 MFLOP, fwd+bwd 204.96 MFLOP (train mode)
MLP_FLOP_PER_RAY_FWD_BWD = 204.96e6
METRIC = "NeFeS rays/sec (fwd+bwd, 64+64 samples)"
MLP_FLOP_FINE_FWD_PER_POINT = 2 * 184064       # MACs per point of the fine field


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch per kernel tag at the bench shape, from this round's
    `ncu --set full` capture (tools/ncu_traffic.py writes profiles/r2_ncu_traffic.json next to the .txt summary)."""
    p = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")
    try:
        return json.load(open(p))
    except Exception:
        return {}


def nerfw_loss(ret, target, lambda_u=0.01):
    """Caller-side loss of the stage-1 step: script/models/losses.py:112-132 (coef 1)."""
    c_l = 0.5 * ((ret["rgb0"] - target) ** 2).mean()
    f_l = ((ret["rgb_map"] - target) ** 2 / (2 * ret["beta"].unsqueeze(1) ** 2)).mean()
    return c_l + f_l + 3 + torch.log(ret["beta"]).mean() + lambda_u * ret["transient_sigmas"].mean()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1387.4), d.get("hbm_gbs", 6553.3), "measured (MEASURED_PEAKS.json: copy bandwidth, sustained bf16 GEMM)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clocks / throttle reasons sampled DURING the timed region: NVML in-process every 5 ms (what nvidia-smi reads; a
    timed region of 20 steps is ~60 ms, shorter than one nvidia-smi start-up on an 8-GPU box), plus the recipe's
    `nvidia-smi -lms` line as a second source when it manages to print inside the region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index
        self.nv, self.nv_max, self.nv_reasons, self.stop, self.nth = [], None, set(), False, None

    def _nvml(self):
        try:
            import pynvml as N
            N.nvmlInit()
            h = N.nvmlDeviceGetHandleByIndex(self.index)
            self.nv_max = float(N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM))
            bits = (("hw_slowdown", N.nvmlClocksThrottleReasonHwSlowdown), ("hw_thermal_slowdown", N.nvmlClocksThrottleReasonHwThermalSlowdown),
                    ("sw_thermal_slowdown", N.nvmlClocksThrottleReasonSwThermalSlowdown), ("sw_power_cap", N.nvmlClocksThrottleReasonSwPowerCap))
            while not self.stop:
                self.nv.append(float(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)))
                r = N.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for name, bit in bits:
                    if r & bit:
                        self.nv_reasons.add(name)
                time.sleep(0.005)
        except Exception:
            pass

    def __enter__(self):
        self.nth = threading.Thread(target=self._nvml, daemon=True)
        self.nth.start()
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        self.stop = True
        if self.nth is not None:
            self.nth.join(timeout=1)
        if self.proc is not None:
            if not self.nv:
                time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])), mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if self.nv:
            out = {"sm_mhz": float(np.median(self.nv)), "sm_max_mhz": max([self.nv_max or 0.0] + mx), "reasons": sorted(reasons | self.nv_reasons),
                   "samples": len(self.nv), "source": "nvml (in-process, 5 ms)" + (f" + nvidia-smi ({len(sm)} lines)" if sm else "")}
            return out
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 100"}


# --------------------------------------------------------------------------------------------------
def host_batches(n_steps, seed, pinned=True, n_rand=N_RAND, feat=False, patches=False):
    """Per-step host inputs, as the reference's DataLoader + np.random.choice produce them
    (run_nefes.py:47-71): poses [4,3,4], pixel indices [4,n_rand], target rgb [4*n_rand,3], hist [4,10]
    (+ target features [4*n_rand,128] for the stage-2 step)."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "poses_stairs.npz"))
    poses = torch.tensor(g["train_gt"][:N_IMAGES].reshape(N_IMAGES, 3, 4), dtype=torch.float32)
    rng = np.random.RandomState(seed)
    pin = (lambda t: t.pin_memory()) if (pinned and torch.cuda.is_available()) else (lambda t: t)
    out = []
    for _ in range(n_steps):
        if patches:                                  # run_nefes.py:86-98: the same 7 random 16x16 crops for every image of the batch
            h0, w0 = rng.randint(0, H - 16, 7), rng.randint(0, W - 16, 7)
            one = np.concatenate([((h0[i] + np.arange(16))[:, None] * W + (w0[i] + np.arange(16))[None, :]).reshape(-1) for i in range(7)])
            idx = np.stack([one] * N_IMAGES)
        else:
            idx = np.stack([rng.choice(H * W, n_rand, replace=False) for _ in range(N_IMAGES)])
        b = dict(pose=pin(poses.clone()), idx=pin(torch.from_numpy(idx).long()),
                 target=pin(torch.from_numpy(rng.rand(N_IMAGES * n_rand, 3).astype(np.float32))),
                 hist=pin(torch.zeros(N_IMAGES, 10) if not patches else torch.from_numpy(rng.randint(0, 30, (N_IMAGES, 10)).astype(np.float32))))
        if feat:
            b["target_f"] = pin(torch.from_numpy(rng.randn(N_IMAGES * n_rand, 128).astype(np.float32)))
        out.append(b)
    return out


class Args:
    nerfh_nff, use_fine_only, NeRFW, transient_at_test, netchunk = True, False, True, True, 1 << 21


def train_workload(workload, n_rand, world=1):
    """The part of `config` both arms share for the training workloads: same string, same numbers."""
    wl = {"train": "C2 stage-1 colour-only NeRF-W training step",
          "c3": "C3 stage-2 training step (C2 + 0.04 x L1 of the rendered 128-channel feature map against target features)",
          "c3s3": "C3 stage-3 training step (7 patches of 16x16 per image; affine colour transform, FusionNet (BatchNorm batch statistics per "
                  "rank), colour + 0.02 x feature L1 + 0.02 x fusion L1; fields, FusionNet and exposure network all train)"}[workload]
    rays = N_IMAGES * n_rand
    return {"workload": f"{wl}, 7-Scenes-stairs camera 640x480 -> 60x80, 4 images x {n_rand} rays = {rays} rays/GPU/step, "
                        "64 coarse + 64 fine samples, Adam",
            "rays_per_gpu_per_step": rays, "global_rays_per_step": rays * world}


def engine_setup(a):
    """Process group, device, the two fields and the render kwargs create_nerf builds."""
    import torch.distributed as dist
    import nefes_b200 as nb
    from nefes_b200 import _lib
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL prints its version banner on stdout at levels VERSION and WARN unless a debug file is named (honoured only
        # above VERSION): stdout carries ONE JSON line, so send NCCL's output to stderr
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    assert world == a.gpus, f"--gpus {a.gpus} but WORLD_SIZE={world}"
    _lib.lib()
    coarse = nb.NeRFH_NFF("coarse", W=128, precision=a.precision).to(dev)
    fine = nb.NeRFH_NFF("fine", W=128, encode_appearance=True, encode_transient=True, precision=a.precision).to(dev)
    q = nb.StandardQuery(Args.netchunk)          # create_nerf's query function: render_rays is one engine call
    kw = dict(network_query_fn=q, N_importance=64, N_samples=64, network_fn=coarse, network_fine=fine,
              use_viewdirs=True, white_bkgd=False, args=Args(), ndc=False, lindisp=False, near=NEAR, far=FAR,
              perturb=1., raw_noise_std=0., test_time=False, retraw=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    return dict(nb=nb, dist=dist, rank=rank, world=world, local=local, dev=dev, coarse=coarse, fine=fine, kw=kw,
                barrier=barrier, max_over_ranks=max_over_ranks)


def finish(c, out):
    if c["rank"] == 0:
        print(json.dumps(out), flush=True)
    if c["world"] > 1:
        # a process group whose collectives were captured into a CUDA graph does not always tear down cleanly
        # (observed: destroy_process_group / interpreter exit hanging after the result was printed): drain the device,
        # meet once more, and leave without running the teardown
        torch.cuda.synchronize()
        c["dist"].barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def run_train(a, c):
    """Workloads train (C2) and c3 (stage 2): the optimiser step of run_nefes.py:113-270 through the public API."""
    from nefes_b200 import _lib, ops, parallel
    nb, dev, world, rank, coarse, fine, kw = c["nb"], c["dev"], c["world"], c["rank"], c["coarse"], c["fine"], c["kw"]
    barrier, max_over_ranks = c["barrier"], c["max_over_ranks"]
    stage2 = a.workload in ("c3", "c3s3")
    stage3 = a.workload == "c3s3"
    n_rand = N_RAND // world if a.scaling == "strong" else N_RAND
    if a.as_world > 1:                         # tuning aid: one GPU doing the per-GPU share of a strong-scaling job of that size
        n_rand = N_RAND // a.as_world
    if stage3:
        n_rand = 7 * 16 * 16                                                     # run_nefes.py:86-94: 7 crops of 16x16 per image
    rays = N_IMAGES * n_rand
    params = [coarse.flat, fine.flat]          # stages 1-2 train the two fields (FusionNet joins at stage 3)
    opt = nb.FlatAdam(params, lr=5e-4)
    extra = []
    if stage3:                                 # nerfh_nff.py:661-682: grad_vars = every parameter of both models
        extra = list(coarse.fusion_net.parameters()) + list(coarse.exposure_embedding.parameters())
        coarse.fusion_net.gemm_tf32 = a.precision != "fp32"          # the convolutions' GEMMs on tcgen05 kind::tf32
        opt_extra = torch.optim.Adam(extra, lr=5e-4, betas=(0.9, 0.999), capturable=True)

    class EncArgs:
        encode_hist = True
    loss_fn = nb.NerfWLoss(coef=1, lambda_u=0.01)                                # losses.py:96-132 on two kernels
    loss_fn2 = nb.ColorFeatureFusionNerfWLoss(coef=1, L1_loss=True)              # run_nefes.py:360

    def step(b):
        """b: dict of DEVICE tensors.  One optimiser step, returns the loss tensor."""
        ro, rd = nb.get_rays_batch(H, W, FOCAL, b["pose"])                      # [4,60,80,3]
        ro = torch.gather(ro.reshape(N_IMAGES, -1, 3), 1, b["idx"][..., None].expand(-1, -1, 3)).reshape(-1, 3)
        rd = torch.gather(rd.reshape(N_IMAGES, -1, 3), 1, b["idx"][..., None].expand(-1, -1, 3)).reshape(-1, 3)
        hist = b["hist"][:, None, :].expand(-1, n_rand, -1).reshape(-1, 10)
        rgb, disp, acc, ex = nb.render(H, W, FOCAL, chunk=32768, rays=(ro, rd), img_idx=hist, **kw)
        res = {"rgb_coarse": ex["rgb0"], "rgb_fine": rgb, "beta": ex["beta"], "transient_sigmas": ex["transient_sigmas"]}
        if stage3:                                                               # run_nefes.py:150-161, 238-243
            rgb_t = coarse.affine_color_transform(EncArgs(), rgb, b["hist"], N_IMAGES)
            _, _, fus = coarse.run_fusion_net(rgb_t, ex["feat_map"], 16, 16, N_IMAGES * 7)
            res.update(rgb_fine=rgb_t, feat_fine=ex["feat_map"], feat_fusion=fus.permute(0, 2, 3, 1).reshape(-1, 128))
            l_rgb, l_f, l_fu = loss_fn2(res, {"rgb": b["target"], "feat": b["target_f"]}, switch_on=True, color_only_switch=False)
            loss = l_rgb + 0.02 * l_f + 0.02 * l_fu
        elif stage2:                                                             # run_nefes.py:244-248
            res["feat_fine"] = ex["feat_map"]
            l_rgb, l_f = loss_fn2(res, {"rgb": b["target"], "feat": b["target_f"]}, switch_on=False, color_only_switch=False)
            loss = l_rgb + 0.04 * l_f
        else:
            loss = loss_fn(res, b["target"])
        opt.zero_grad(set_to_none=True)
        if stage3:
            opt_extra.zero_grad(set_to_none=True)
        loss.backward()
        if world > 1:
            parallel.allreduce_grads(params)
            for p_ in extra:                   # BatchNorm statistics stay per rank (the reference is single-GPU): stated deviation
                dist_.all_reduce(p_.grad)
                p_.grad.div_(world)
        opt.step(grad_scale=1.0 / world)
        if stage3:
            opt_extra.step()
        return loss

    dist_ = c["dist"]
    n_total = a.warmup + a.steps
    host = host_batches(n_total, seed=1000 + rank, n_rand=n_rand, feat=stage2, patches=stage3)
    resident = [{k: v.to(dev) for k, v in b.items()} for b in host]
    torch.cuda.synchronize()

    # ---- warm-up (eager), then the step is captured ONCE into a CUDA graph: ~100 launches, most of them small, replay
    #      without per-launch CPU cost.  Inputs enter through static device buffers; Adam's step count / lr live on the
    #      device (FlatAdam), torch's generator is graph-aware, NCCL all-reduce is capturable. --no-graph: eager steps.
    for b in resident[:a.warmup]:
        step(b)
    barrier()
    static = {k: v.clone() for k, v in resident[0].items()}
    graph, static_loss = None, None
    if not a.no_graph:
        try:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_loss = step(static)
            torch.cuda.synchronize()
        except Exception as e:                       # capture is an optimisation, not a requirement
            sys.stderr.write(f"[bench] CUDA-graph capture failed ({type(e).__name__}: {e}); timing eager steps\n")
            graph = None
            torch.cuda.synchronize()

    def run_step(b):
        if graph is None:
            return step(b)
        for k, v in b.items():
            static[k].copy_(v, non_blocking=True)
        graph.replay()
        return static_loss

    # ---- per-kernel table: a few EAGER steps with the library's CUDA events around every hot launch -----------------
    ops.PROFILE = []
    _lib.lib().nefes_prof_enable(1)
    l0 = _lib.lib().nefes_launch_count()
    n_prof = min(3, a.steps)
    for b in resident[a.warmup:a.warmup + n_prof]:
        step(b)
    torch.cuda.synchronize()
    launches_per_step = int(_lib.lib().nefes_launch_count() - l0) // n_prof
    prof, ops.PROFILE = ops.PROFILE, None
    import ctypes
    buf = ctypes.create_string_buffer(1 << 16)
    _lib.check(_lib.lib().nefes_prof_report(buf, len(buf)), "nefes_prof_report")
    _lib.lib().nefes_prof_enable(0)
    kern = json.loads(buf.value.decode())
    for k in kern.values():                          # per-step figures below divide by a.steps
        k["launches"] *= a.steps / n_prof; k["ms"] *= a.steps / n_prof; k["alg_bytes"] *= a.steps / n_prof; k["alg_flops"] *= a.steps / n_prof
    mlp_ms = sum(s.elapsed_time(e) for _, s, e in prof) / n_prof

    # ---- device-resident timing ---------------------------------------------------------------
    run_step(resident[0])
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(c["local"]) as clk:
        ev0.record()
        for b in resident[a.warmup:]:
            run_step(b)
        ev1.record()
        barrier()
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    launches = launches_per_step * a.steps
    value = world * rays * a.steps / (ms / 1e3)

    # ---- end to end: host inputs from pinned memory, loss read back, every step ----------------
    losses = []
    barrier()
    ev0.record()
    for b in host[a.warmup:]:
        if graph is None:
            d = {k: v.to(dev, non_blocking=True) for k, v in b.items()}
            losses.append(float(step(d).item()))
        else:
            losses.append(float(run_step(b).item()))          # H2D straight into the graph's input buffers
    ev1.record()
    barrier()
    ms_e2e = max_over_ranks(ev0.elapsed_time(ev1))
    e2e_value = world * rays * a.steps / (ms_e2e / 1e3)
    h2d = sum(v.numel() * v.element_size() for v in host[0].values())

    tf_peak, hbm_peak, which = measured_peaks()
    achieved = MLP_FLOP_PER_RAY_FWD_BWD * rays / (mlp_ms / 1e3) / 1e12
    # per-kernel table: achieved = algorithmic bytes (flops) of the launches / their measured duration
    table = {}
    for tag, k in kern.items():
        if k["ms"] <= 0:
            continue
        table[tag] = {"launches_per_step": k["launches"] / a.steps, "ms_per_step": k["ms"] / a.steps,
                      "GB_per_s": k["alg_bytes"] / k["ms"] / 1e6, "TFLOP_per_s": k["alg_flops"] / k["ms"] / 1e9,
                      "hbm_frac": k["alg_bytes"] / k["ms"] / 1e6 / hbm_peak, "tensor_frac": k["alg_flops"] / k["ms"] / 1e9 / tf_peak}
    # dominant kernel = the single launch that takes longest (a tag that launches 8 small kernels does not outrank it)
    top = max(table, key=lambda t: table[t]["ms_per_step"] / table[t]["launches_per_step"]) if table else None
    roof = None
    if top:
        t = table[top]
        per_launch = t["ms_per_step"] / t["launches_per_step"]
        # K5 (the field MLP) is TENSOR-bound -- algorithmic FLOP of the launch (2 x MACs/point x points) over
        # its measured duration against the sustained bf16 peak is the headline; the same launch against the HBM peak
        # (algorithmic bytes INCLUDING the saved activation copies this design writes) is the second figure.
        roof = {"bound": "tensor", "kernel": top, "achieved": t["TFLOP_per_s"], "peak": tf_peak, "unit": "TFLOP/s",
                "frac": t["tensor_frac"], "traffic": ncu_traffic().get(top), "peak_source": which,
                "ms_per_launch": per_launch, "ms_per_step_in_kernel": t["ms_per_step"],
                "share_of_step": t["ms_per_step"] / (ms / a.steps),
                "hbm": {"bound": "hbm", "achieved": t["GB_per_s"], "peak": hbm_peak, "unit": "GB/s", "frac": t["hbm_frac"],
                        "note": "algorithmic bytes include the bf16 activation copies saved for backward (design choice, DESIGN.md 4)"}}
    out = {
        "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None,
        "dtype": {"fp32": "f32", "tf32": "tf32", "bf16": "bf16"}[a.precision], "data": "synthetic",
        "config": {**train_workload(a.workload, n_rand, world), "parallelism": f"dp{world} (rays sharded, 1 gradient all-reduce)",
                   "mlp_precision": a.precision, "cuda_graph": graph is not None,
                   "l2": "no explicit flush: each step streams >20 GB of activation / gradient tiles through HBM (>> 126 MB L2); "
                         "only the 1.4 MB of weights is legitimately L2-resident"},
        "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / a.steps},
        "gpu_launches": launches,
        "roofline": roof,
        "kernels": table,
        "mlp_tensor": {"what": "field MLP (K5) forward+backward, all launches of one step, against the tensor roofline",
                       "achieved": achieved, "peak": tf_peak, "unit": "TFLOP/s", "frac": achieved / tf_peak,
                       "ms_per_step": mlp_ms, "share_of_step": mlp_ms / (ms / a.steps)},
        "clocks": clk.summary(),
        "final_loss": losses[-1] if losses else None,
    }
    graph = None
    # BASELINE's second metric at every N: refinement iterations/s (queries sharded over ranks, no collective)
    if not a.no_extras and not stage2:
        out["refine_iters_per_s"] = refine_rate(a, c, n_queries=1)["iters_per_s"]
    if world == 1 and not a.no_extras and not stage2:
        out["extras"] = side_measurements(a, nb, step, coarse, fine, resident, kw, dev)
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(a)
    finish(c, out)


def refine_rate(a, c, n_queries, n_iters=50, e2e=False):
    """C4: every rank refines `n_queries` queries of `n_iters` iterations (bf16 fields, engine-resident iteration replayed
    from a CUDA graph); whole-job iterations/s = world x queries x iters / max-over-ranks device time."""
    from nefes_b200 import refine
    dev, world, rank, coarse, fine = c["dev"], c["world"], c["rank"], c["coarse"], c["fine"]
    g = np.load(os.path.join(ROOT, "tests", "golden", "poses_stairs.npz"))
    kwt = dict(c["kw"])
    kwt.update(perturb=0., test_time=True)
    kwt.pop("retraw", None)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    rng = np.random.RandomState(77 + rank)
    n_all = g["dfnet_init"].shape[0]
    ids = [(rank + i * world) % n_all for i in range(n_queries + 1)]          # rank r: queries r, r+W, ... (8e)
    host_q = [(torch.tensor(g["dfnet_init"][i].reshape(3, 4), dtype=torch.float32).pin_memory(),
               torch.from_numpy(rng.randn(128, H * W).astype(np.float32)).pin_memory()) for i in ids]
    for p in (coarse.flat, fine.flat):
        p.requires_grad_(False)
    try:
        init, tgt = host_q[0][0].to(dev), host_q[0][1].to(dev)
        refine.refine_pose(init, tgt, H, W, FOCAL, kwt, n_iters=10)             # warm-up query: captures the iteration graph
        dev_q = [(i_.to(dev), t_.to(dev)) for i_, t_ in host_q[1:]]
        c["barrier"]()
        ev0.record()
        poses = []
        for qi in range(n_queries):
            if e2e:
                i_, t_ = host_q[1 + qi][0].to(dev, non_blocking=True), host_q[1 + qi][1].to(dev, non_blocking=True)
                poses.append(refine.refine_pose(i_, t_, H, W, FOCAL, kwt, n_iters=n_iters)[0].cpu())
            else:
                poses.append(refine.refine_pose(dev_q[qi][0], dev_q[qi][1], H, W, FOCAL, kwt, n_iters=n_iters)[0])
        ev1.record()
        c["barrier"]()
        ms = c["max_over_ranks"](ev0.elapsed_time(ev1))
    finally:
        for p in (coarse.flat, fine.flat):
            p.requires_grad_(True)
    its = world * n_queries * n_iters
    return {"iters_per_s": its / (ms / 1e3), "ms_per_iter": ms / (n_queries * n_iters), "queries_per_s": world * n_queries / (ms / 1e3),
            "ms": ms, "h2d_bytes_per_query": 48 + 128 * H * W * 4, "d2h_bytes_per_query": 48}


def run_refine(a, c):
    """Workload refine (C4): a step = one query of 50 pose-gradient iterations per GPU."""
    from nefes_b200 import _lib
    world = c["world"]
    l0 = _lib.lib().nefes_launch_count()
    with ClockSampler(c["local"]) as clk:
        r = refine_rate(a, c, n_queries=a.steps)
    launches = int(_lib.lib().nefes_launch_count() - l0)
    r2 = refine_rate(a, c, n_queries=a.steps, e2e=True)
    tf_peak, hbm_peak, which = measured_peaks()
    flop_iter = 111.0e6 * H * W                       # refine fwd 63.88 + bwd 47.12 MFLOP per ray
    ach = flop_iter / (r["ms_per_iter"] / 1e3) / 1e12
    out = {"metric": "NeFeS refine iters/sec (50 pose-gradient iterations per query, 60x80 render, 64+64 samples)",
           "value": r["iters_per_s"], "unit": "iters/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
           "ms_per_step": r["ms"] / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": {"fp32": "f32", "tf32": "tf32", "bf16": "bf16"}[a.precision], "data": "synthetic",
           "config": {"workload": "C4 DFNet+NeFeS50 test-time refinement: per query 50 Adam iterations on the 6 pose parameters, each a full "
                                  "60x80 test_time render (4800 rays, 64+64 samples) + cosine feature loss + backward to the pose; "
                                  "initial poses = DFNet_stairs_results rows, random-normal target features, frozen random-init fields; "
                                  "FusionNet / exposure MLP / DFNet excluded (SURVEY.md 8d C4)",
                      "queries_per_gpu": a.steps, "parallelism": f"dp{world} (queries r, r+N, ...; no collective)", "mlp_precision": a.precision,
                      "cuda_graph": True, "l2": "each iteration streams ~3 GB of activation / gradient tiles (>> L2)"},
           "queries_per_s": r["queries_per_s"], "ms_per_iter": r["ms_per_iter"],
           "e2e": {"value": r2["iters_per_s"], "unit": "iters/s", "h2d_bytes_per_step": r2["h2d_bytes_per_query"],
                   "d2h_bytes_per_step": r2["d2h_bytes_per_query"], "ms_per_step": r2["ms"] / a.steps},
           "gpu_launches": launches,
           "roofline": {"bound": "tensor", "kernel": "refinement iteration (all launches)", "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s",
                        "frac": ach / tf_peak, "traffic": None, "peak_source": which},
           "clocks": clk.summary()}
    if c["rank"] == 0 and world == 1 and not a.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(a)
    finish(c, out)


def run_sweep(a, c):
    """Workload sweep (C5-shaped): test_time render (no gradients) of 2^12 .. 2^20 rays in the reference's chunks of 32768,
    contiguous ray ranges per rank (8e), front-end A.  `value` = rays/s at 2^20 rays."""
    from nefes_b200 import _lib, parallel
    nb, dev, world, rank = c["nb"], c["dev"], c["world"], c["rank"]
    kwt = dict(c["kw"])
    kwt.update(perturb=0., test_time=True, near=0., far=10.)
    kwt.pop("retraw", None)
    hash_fe = a.frontend == "hash"
    if hash_fe:
        # front-end B: NeRFH_TCNN fields (HashGrid 16 levels x 2 features, T = 2^log2T, bound 25; SH degree 4;
        # heads 64 wide), the tcnn query function, heads on the tf32 tensor-core GEMM
        from nefes_b200.hashgrid import NeRFH_TCNN, TcnnQuery
        hc = NeRFH_TCNN("coarse", bound=25, log2_hashmap_size=a.log2T).to(dev)
        hf = NeRFH_TCNN("fine", encode_appearance=True, encode_transient=True, in_channels_a=50, in_channels_t=20, bound=25,
                        log2_hashmap_size=a.log2T).to(dev)

        class HArgs:
            nerfh_nff, use_fine_only, NeRFW, transient_at_test = False, False, True, True
        kwt.update(network_fn=hc, network_fine=hf, network_query_fn=TcnnQuery(1 << 21), args=HArgs())
        if a.precision != "fp32":
            _lib.lib().nefes_gemm_mode(1)
    Hc, Wc, fc = 60, 106, 93.0                        # Cambridge-shaped camera (cambridge_scenes.py:149)
    g = np.load(os.path.join(ROOT, "tests", "golden", "poses_stairs.npz"))
    pose = torch.tensor(g["test_gt"][0].reshape(3, 4), dtype=torch.float32, device=dev)
    ro, rd = nb.get_rays(Hc, Wc, fc, pose)
    ro, rd = ro.reshape(-1, 3), rd.reshape(-1, 3)
    gen = torch.Generator(device=dev).manual_seed(5)
    hist = torch.zeros(1, 10, device=dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    table, launches = {}, 0
    sizes = [1 << k for k in range(12, 19 if hash_fe else 21)]
    with torch.no_grad(), ClockSampler(c["local"]) as clk:
        for n in sizes:
            lo, hi = parallel.shard_range(n, rank, world)
            pix = torch.randint(0, Hc * Wc, (hi - lo,), device=dev, generator=gen)
            rays = (ro[pix].contiguous(), rd[pix].contiguous())
            reps = max(a.steps if n == sizes[-1] else 3, 1)
            for _ in range(max(a.warmup if n == sizes[-1] else 1, 1)):
                nb.render(Hc, Wc, fc, chunk=32768, rays=rays, img_idx=hist, **kwt)
            c["barrier"]()
            l0 = _lib.lib().nefes_launch_count()
            ev0.record()
            for _ in range(reps):
                rgb, _, _, ex = nb.render(Hc, Wc, fc, chunk=32768, rays=rays, img_idx=hist, **kwt)
            ev1.record()
            c["barrier"]()
            ms = c["max_over_ranks"](ev0.elapsed_time(ev1)) / reps
            launches = int(_lib.lib().nefes_launch_count() - l0)
            table[str(n)] = {"rays_per_s": n / (ms / 1e3), "ms": ms, "rays_per_gpu": hi - lo}
        # end to end at the largest size: rays from pinned host memory, rgb + feature map read back
        n = sizes[-1]
        lo, hi = parallel.shard_range(n, rank, world)
        h_o, h_d = rays[0].cpu().pin_memory(), rays[1].cpu().pin_memory()
        h_rgb = torch.empty(hi - lo, 3).pin_memory()             # results land in pinned host memory (a pageable .cpu() of the
        h_feat = torch.empty(hi - lo, 128).pin_memory()          # 128-channel feature map runs at a tenth of the link rate)
        c["barrier"]()
        ev0.record()
        for _ in range(a.steps):
            r_ = (h_o.to(dev, non_blocking=True), h_d.to(dev, non_blocking=True))
            rgb, _, _, ex = nb.render(Hc, Wc, fc, chunk=32768, rays=r_, img_idx=hist, **kwt)
            h_rgb.copy_(rgb, non_blocking=True)
            if "feat_map" in ex and ex["feat_map"].shape[-1] == 128:
                h_feat.copy_(ex["feat_map"], non_blocking=True)
        ev1.record()
        c["barrier"]()
        ms_e2e = c["max_over_ranks"](ev0.elapsed_time(ev1)) / a.steps
    tf_peak, hbm_peak, which = measured_peaks()
    top = table[str(n)]
    if hash_fe:
        _lib.lib().nefes_gemm_mode(0)
        table_mb = float(hc.encoder.params.numel() * 4 / 1e6)
        gather = 192 * 16 * 8 * 2 * 4                    # bytes gathered per ray (fp32 table): 192 points x 1024 B
        out = {"metric": "NeFeS rays/sec (inference render, 64+64 samples)", "value": top["rays_per_s"], "unit": "rays/s", "n_gpus": world,
               "steps": a.steps, "warmup": a.warmup, "ms_per_step": top["ms"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
               "dtype": "f32 tables, tf32 heads" if a.precision != "fp32" else "f32", "data": "synthetic",
               "config": {"workload": f"C5-shaped ray sweep, front-end B (HashGrid L=16 F=2 T=2^{a.log2T}, bound 25, SH degree 4, heads 64 wide; "
                                      "PARITY UNPINNED: tiny-cuda-nn is not vendored), Cambridge camera, test_time render without gradients, "
                                      f"2^12..2^{len(sizes) + 11} rays in chunks of 32768; value at the largest size",
                          "global_rays": n, "parallelism": f"dp{world}", "table_MB_per_field": table_mb, "cuda_graph": False,
                          "l2": f"hash tables {table_mb:.0f} MB per field x 2 fields against 126 MB of L2"},
               "sweep": table,
               "e2e": {"value": n / (ms_e2e / 1e3), "unit": "rays/s", "h2d_bytes_per_step": (hi - lo) * 24, "d2h_bytes_per_step": (hi - lo) * 12,
                       "ms_per_step": ms_e2e},
               "gpu_launches": launches * a.steps,
               "roofline": {"bound": "hbm", "kernel": "hash gather (all levels, both fields)", "achieved": gather * n / (top["ms"] / 1e3) / 1e9,
                            "peak": hbm_peak, "unit": "GB/s", "frac": gather * n / (top["ms"] / 1e3) / 1e9 / hbm_peak, "traffic": None,
                            "peak_source": which,
                            "note": "algorithmic gather bytes of the WHOLE render over its duration (the staged route also spends time in the "
                                    "heads and torch glue); a table of 2^19 entries per level is L2-resident, 2^21 / 2^22 spill to HBM"},
               "clocks": clk.summary()}
        finish(c, out)
        return
    ach = 63.88e6 * n / (top["ms"] / 1e3) / 1e12      # refine/inference forward 63.88 MFLOP per ray
    out = {"metric": "NeFeS rays/sec (inference render, 64+64 samples)", "value": top["rays_per_s"], "unit": "rays/s", "n_gpus": world,
           "steps": a.steps, "warmup": a.warmup, "ms_per_step": top["ms"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": {"fp32": "f32", "tf32": "tf32", "bf16": "bf16"}[a.precision], "data": "synthetic",
           "config": {"workload": "C5-shaped ray sweep, front-end A: Cambridge camera 60x106 f=93, near 0 far 10, test_time render without "
                                  "gradients, 2^12..2^20 rays in chunks of 32768; value at 2^20 rays", "global_rays": n,
                      "parallelism": f"dp{world} (contiguous ray ranges, no collective)", "mlp_precision": a.precision, "cuda_graph": False,
                      "l2": "2^20 rays stream 30 GB of raw tiles through HBM (>> L2)"},
           "sweep": table,
           "e2e": {"value": n / (ms_e2e / 1e3), "unit": "rays/s", "h2d_bytes_per_step": (hi - lo) * 24, "d2h_bytes_per_step": (hi - lo) * 131 * 4,
                   "ms_per_step": ms_e2e},
           "gpu_launches": launches * a.steps,
           "roofline": {"bound": "tensor", "kernel": "inference render (all launches)", "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s",
                        "frac": ach / tf_peak, "traffic": None, "peak_source": which},
           "clocks": clk.summary()}
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(a)          # the reference's render() on the host cores, 4096-ray inference renders
    finish(c, out)


def run_engine(a):
    c = engine_setup(a)
    if a.workload in ("train", "c3", "c3s3"):
        run_train(a, c)
    elif a.workload == "refine":
        run_refine(a, c)
    else:
        run_sweep(a, c)


def side_measurements(a, nb, step, coarse, fine, resident, kw, dev):
    """Reported next to the headline, not part of it: the same training step with the fp32 (parity) field path, and
    BASELINE's second metric -- refinement iterations/s (C4: full 60x80 render from a pose, test_time, cosine
    feature loss, backward to the 6 pose parameters, Adam), in both arithmetics."""
    from nefes_b200 import refine
    ex = {}
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for other in ("fp32", "tf32"):
        if a.precision == other:
            continue
        coarse.precision = fine.precision = other
        for b in resident[:2]:
            step(b)
        torch.cuda.synchronize()
        ev0.record()
        for b in resident[:5]:
            step(b)
        ev1.record()
        torch.cuda.synchronize()
        ex[f"train_step_{other}_field"] = {"rays_per_s": RAYS * 5 / (ev0.elapsed_time(ev1) / 1e3), "ms_per_step": ev0.elapsed_time(ev1) / 5}
        coarse.precision = fine.precision = a.precision
    # encoder front-end B (K4b): HashGrid gather, forward and backward, at the bench step's point count
    from nefes_b200.hashgrid import HashGridEncoding
    enc = HashGridEncoding().to(dev)
    npts = RAYS * 192
    xs = torch.rand(npts, 3, device=dev)
    for tag, need_grad in (("fwd", False), ("fwd_bwd", True)):
        for rep in range(2):                          # first pass = warm-up
            torch.cuda.synchronize()
            ev0.record()
            y = enc(xs)
            if need_grad:
                y.backward(torch.ones_like(y))
                enc.params.grad = None
            ev1.record()
            torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
        # algorithmic bytes per point: 16 levels x 8 corners x 2 features x 4 B gathered (+ the same again scattered
        # in backward) + 128 B of output row (+ 128 B of cotangent)
        bytes_pt = 16 * 8 * 8 + 128 if not need_grad else 2 * (16 * 8 * 8 + 128) + 16 * 8 * 8
        ex[f"hashgrid_{tag}"] = {"points": npts, "ms": ms, "gather_GB_per_s": npts * bytes_pt / ms / 1e6,
                                 "table_MB": float(enc.params.numel() * 4 / 1e6),
                                 "note": "fp32 table, T=2^19: 49 MB, resident in the 126 MB L2 -- the gather rate is an L2 / L1 figure and has no "
                                         "HBM roofline; at T=2^21 / 2^22 (169 / 316 MB) it spills (bench.py --workload sweep --frontend hash)"}
    del enc, xs
    g = np.load(os.path.join(ROOT, "tests", "golden", "poses_stairs.npz"))
    init = torch.tensor(g["dfnet_init"][0].reshape(3, 4), dtype=torch.float32, device=dev)
    target = torch.randn(128, H * W, device=dev)
    kwt = dict(kw)
    kwt.update(perturb=0., test_time=True)
    kwt.pop("retraw", None)
    for p in (coarse.flat, fine.flat):
        p.requires_grad_(False)
    try:
        for prec in ("fp32", "bf16"):
            coarse.precision = fine.precision = prec
            refine.refine_pose(init, target, H, W, FOCAL, kwt, n_iters=10)      # warm-up query: captures the iteration graph
            torch.cuda.synchronize()
            ev0.record()
            refine.refine_pose(init, target, H, W, FOCAL, kwt, n_iters=20)
            ev1.record()
            torch.cuda.synchronize()
            ms = ev0.elapsed_time(ev1) / 20
            ex[f"refine_{prec}"] = {"iters_per_s": 1e3 / ms, "ms_per_iter": ms, "rays_per_iter": H * W,
                                    "queries_per_s_at_50_iters": 1e3 / ms / 50}
    finally:
        coarse.precision = fine.precision = a.precision
        for p in (coarse.flat, fine.flat):
            p.requires_grad_(True)
    return ex


# --------------------------------------------------------------------------------------------------
def reference_arm(workload, n_rays):
    """-> (step, kind, what): `step()` runs ONE step of `workload` on the host cores and returns the number of work units
    (rays, or refinement iterations) it processed.  kind "reference": the UNMODIFIED reference modules staged under
    oracle/_ref (oracle/build_ref.py) -- its render(), its NeRFH_NFF, its losses, torch.optim.Adam as create_nerf builds it;
    kind "port": the oracle restatement, when no reference tree travels with the repo."""
    from oracle import ref_loader
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    hb = host_batches(1, seed=7, pinned=False, feat=True, patches=(workload == "c3s3"), n_rand=(1792 if workload == "c3s3" else N_RAND))[0]
    g = np.load(os.path.join(ROOT, "tests", "golden", "poses_stairs.npz"))
    root = ref_loader.reference_root()
    if root is not None:
        R, M, U = ref_loader.import_reference(root)
        import models.losses as RLoss
        coarse, fine, base = ref_loader.reference_render_kwargs(M)
        params = list(coarse.parameters()) + list(fine.parameters())           # nerfh_nff.py:661-680 grad_vars
        if workload in ("train", "c3", "c3s3"):
            n_img_rays = 1792 if workload == "c3s3" else N_RAND
            opt = torch.optim.Adam(params=params, lr=5e-4, betas=(0.9, 0.999))  # nerfh_nff.py:682
            import contextlib
            import io
            with contextlib.redirect_stdout(io.StringIO()):                      # the class prints its mode on stdout
                loss_func = RLoss.ColorFeatureFusionNerfWLoss(coef=1, L1_loss=True)
            ro, rd = U.get_rays_batch(H, W, FOCAL, hb["pose"])
            idx = hb["idx"][..., None].expand(-1, -1, 3)
            ro = torch.gather(ro.reshape(N_IMAGES, -1, 3), 1, idx).reshape(-1, 3)[:n_rays]
            rd = torch.gather(rd.reshape(N_IMAGES, -1, 3), 1, idx).reshape(-1, 3)[:n_rays]
            hist = hb["hist"][:, None, :].expand(-1, n_img_rays, -1).reshape(-1, 10)[:n_rays]

            def step():
                rgb, disp, acc, ex = R.render(H, W, FOCAL, chunk=32768, rays=torch.stack([ro, rd], 0), retraw=True, img_idx=hist,
                                              perturb=1., raw_noise_std=0., test_time=False, near=NEAR, far=FAR, **base)
                res = {"rgb_fine": rgb, "rgb_coarse": ex["rgb0"], "feat_fine": ex["feat_map"], "beta": ex["beta"],
                       "transient_sigmas": ex["transient_sigmas"]}                # run_nefes.py:217-231
                if workload == "c3s3":
                    # the exposure network is tiny-cuda-nn (absent here): the reference arm skips affine_color_transform
                    _, _, fus = coarse.run_fusion_net(rgb, ex["feat_map"], 16, 16, N_IMAGES * 7)
                    res["feat_fusion"] = fus.permute(0, 2, 3, 1).reshape(-1, 128)
                    l_rgb, l_f, l_fu = loss_func(res, {"rgb": hb["target"][:n_rays], "feat": hb["target_f"][:n_rays]}, switch_on=True,
                                                 color_only_switch=False)
                    loss = l_rgb + 0.02 * l_f + 0.02 * l_fu
                elif workload == "c3":
                    l_rgb, l_f = loss_func(res, {"rgb": hb["target"][:n_rays], "feat": hb["target_f"][:n_rays]}, switch_on=False,
                                           color_only_switch=False)
                    loss = l_rgb + 0.04 * l_f
                else:
                    loss = loss_func(res, {"rgb": hb["target"][:n_rays]}, switch_on=False, color_only_switch=True)
                opt.zero_grad()
                loss.backward()
                opt.step()
                return n_rays
        elif workload == "refine":
            import models.poses as RPose
            for p_ in params:
                p_.requires_grad_(False)
            init = torch.eye(4)
            init[:3] = torch.tensor(g["dfnet_init"][0].reshape(3, 4), dtype=torch.float32)
            pose_net = RPose.LearnPose(1, True, True, init_c2w=init[None].clone(), lietorch=False)
            opt = torch.optim.Adam([{"params": [pose_net.r], "lr": 0.0087}, {"params": [pose_net.t], "lr": 0.01}])
            target = torch.randn(128, H * W, generator=torch.Generator().manual_seed(3))
            hist = torch.zeros(1, 10)

            def step():
                c2w = pose_net(0)
                rgb, disp, acc, ex = R.render(H, W, FOCAL, chunk=32768, c2w=c2w[:3, :4], img_idx=hist, perturb=False, raw_noise_std=0.,
                                              test_time=True, near=NEAR, far=FAR, **base)
                loss = 1 - torch.nn.functional.cosine_similarity(ex["feat_map"].t(), target, dim=1, eps=1e-6).mean()
                opt.zero_grad()
                loss.backward()
                opt.step()
                return 1
        else:                                                                      # sweep: inference render
            pose = torch.tensor(g["test_gt"][0].reshape(3, 4), dtype=torch.float32)
            ro, rd = U.get_rays(60, 106, 93.0, pose)
            pix = torch.randint(0, 60 * 106, (n_rays,), generator=torch.Generator().manual_seed(5))
            rays = torch.stack([ro.reshape(-1, 3)[pix], rd.reshape(-1, 3)[pix]], 0)
            hist = torch.zeros(1, 10)

            def step():
                with torch.no_grad():
                    R.render(60, 106, 93.0, chunk=32768, rays=rays, img_idx=hist, perturb=False, raw_noise_std=0., test_time=True,
                             near=0., far=10., **base)
                return n_rays
        return step, "reference", (f"the unmodified reference ({os.path.relpath(root, ROOT) if root.startswith(ROOT) else root}: models/rendering.py render(), "
                                   "nerfh_nff.py NeRFH_NFF, losses.py, torch.optim.Adam), torch CPU fp32, anomaly mode off")
    # ---- no reference tree: the oracle restatement --------------------------------------------------------------------------
    from oracle import nefes_oracle as O
    Pc, Pf = O.clone_params(O.init_field("coarse"), requires_grad=True), O.clone_params(O.init_field("fine"), requires_grad=True)
    if workload in ("train", "c3", "c3s3"):
        opt = torch.optim.Adam(list(Pc.values()) + list(Pf.values()), lr=5e-4)
        ro, rd = O.camera_rays_batch(H, W, FOCAL, hb["pose"])
        idx = hb["idx"][..., None].expand(-1, -1, 3)
        ro = torch.gather(ro.reshape(N_IMAGES, -1, 3), 1, idx).reshape(-1, 3)[:n_rays]
        rd = torch.gather(rd.reshape(N_IMAGES, -1, 3), 1, idx).reshape(-1, 3)[:n_rays]

        def step():
            t_rand, u = torch.rand(n_rays, 64), torch.rand(n_rays, 64)
            ret = O.render(H, W, FOCAL, Pc, Pf, rays=(ro, rd), near=NEAR, far=FAR, test_time=False, t_rand=t_rand, u=u)
            loss = nerfw_loss(ret, hb["target"][:n_rays])
            if workload in ("c3", "c3s3"):       # the port has no FusionNet: stage 3 falls back to the stage-2 loss
                loss = loss + 0.04 * (ret["feat_map"] - hb["target_f"][:n_rays]).abs().mean()
            opt.zero_grad()
            loss.backward()
            opt.step()
            return n_rays
    elif workload == "refine":
        Pc, Pf = O.init_field("coarse"), O.init_field("fine")
        init = torch.tensor(g["dfnet_init"][0].reshape(3, 4), dtype=torch.float32)
        r6 = torch.zeros(6, requires_grad=True)
        opt = torch.optim.Adam([r6], lr=0.01)
        target = torch.randn(128, H * W, generator=torch.Generator().manual_seed(3))

        def step():
            Rm = O.so3_exp(r6[:3]) @ init[:, :3]
            c2w = torch.cat([Rm, (r6[3:] + init[:, 3])[:, None]], 1)
            ret = O.render(H, W, FOCAL, Pc, Pf, c2w=c2w, near=NEAR, far=FAR, test_time=True)
            loss = O.cosine_feature_loss(ret["feat_map"].t(), target)
            opt.zero_grad()
            loss.backward()
            opt.step()
            return 1
    else:
        Pc, Pf = O.init_field("coarse"), O.init_field("fine")
        pose = torch.tensor(g["test_gt"][0].reshape(3, 4), dtype=torch.float32)
        ro, rd = O.camera_rays(60, 106, 93.0, pose)
        pix = torch.randint(0, 60 * 106, (n_rays,), generator=torch.Generator().manual_seed(5))
        rays = (ro.reshape(-1, 3)[pix], rd.reshape(-1, 3)[pix])

        def step():
            with torch.no_grad():
                O.render(60, 106, 93.0, Pc, Pf, rays=rays, near=0., far=10., test_time=True)
            return n_rays
    return step, "port", "oracle port of the reference path (no reference tree on this box), torch CPU fp32, anomaly mode off"


WORK_UNIT = {"train": ("rays/s", RAYS), "c3": ("rays/s", RAYS), "c3s3": ("rays/s", 7168), "refine": ("iters/s", 1), "sweep": ("rays/s", 4096)}


def cpu_baseline(a, budget_s=25.0):
    """The reference's own CPU implementation of the workload on all host cores, on a bounded sample (about 10-30 s of CPU
    work): full-size steps (6144 rays for train / c3, one 4800-ray refinement iteration, a 4096-ray inference render),
    1 warm-up + as many timed steps as fit the budget (at least 2)."""
    unit, n = WORK_UNIT[a.workload]
    step, kind, what = reference_arm(a.workload, n)
    step()
    times, units = [], 0
    while len(times) < 2 or (sum(times) < budget_s and len(times) < 8):
        t0 = time.perf_counter()
        units += step()
        times.append(time.perf_counter() - t0)
    return {"value": units / sum(times), "unit": unit, "cores": os.cpu_count() or 1, "kind": kind,
            "sample": f"{len(times)} full-size steps ({n} {'rays' if unit == 'rays/s' else 'iteration(s)'} each) after 1 warm-up; {what}; torch {torch.__version__}",
            "seconds_per_step": sum(times) / len(times)}


def run_reference(a):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    steps, warm = max(1, a.steps), max(1, a.warmup)
    unit, n = WORK_UNIT[a.workload]
    step, kind, what = reference_arm(a.workload, n)
    # keep the whole run within a few minutes: cap the time spent in timed steps
    budget_s, times, units = 150.0, [], 0
    for i in range(warm + steps):
        t0 = time.perf_counter()
        u = step()
        dt = time.perf_counter() - t0
        if i >= warm:
            times.append(dt)
            units += u
        if sum(times) > budget_s:
            break
    k, tot = len(times), sum(times)
    v = units / tot
    metric = {"train": METRIC, "c3": METRIC, "c3s3": METRIC, "refine": "NeFeS refine iters/sec (50 pose-gradient iterations per query, 60x80 render, 64+64 samples)",
              "sweep": "NeFeS rays/sec (inference render, 64+64 samples)"}[a.workload]
    wl = {"train": "C2 stage-1 colour-only NeRF-W training step (same as the engine arm), full 6144 rays per step",
          "c3": "C3 stage-2 training step (same as the engine arm), full 6144 rays per step",
          "c3s3": "C3 stage-3 training step (patches + FusionNet; the tiny-cuda-nn exposure network is absent, so no affine colour transform), 7168 rays per step",
          "refine": "C4 refinement: one step = ONE pose-gradient iteration (full 60x80 render + cosine loss + backward to the pose + Adam)",
          "sweep": "C5-shaped inference render, one step = 4096 rays"}[a.workload]
    out = {"impl": "reference", "metric": metric, "value": v, "unit": unit, "n_gpus": a.gpus, "steps": k,
           "warmup": warm, "ms_per_step": 1e3 * tot / k, "higher_is_better": True, "scaling": a.scaling if a.workload in ("train", "c3", "c3s3") else "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": ({**train_workload(a.workload, 1792 if a.workload == "c3s3" else N_RAND), "parallelism": "host cores (no GPU)",
                       "note": wl} if a.workload in ("train", "c3", "c3s3") else {"workload": wl, "units_per_step": n}),
           "cpu_baseline": {"value": v, "unit": unit, "cores": os.cpu_count() or 1, "kind": kind,
                            "sample": f"{k} steps x {n} {'rays' if unit == 'rays/s' else 'iteration'}; {what}; all host threads"},
           "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="engine", choices=["engine", "reference"])
    p.add_argument("--precision", default=os.environ.get("NEFES_PRECISION", "bf16"), choices=["fp32", "tf32", "bf16"])
    p.add_argument("--no-extras", action="store_true", help="skip the fp32-path and refinement side measurements")
    p.add_argument("--no-graph", action="store_true", help="time eager steps instead of a captured CUDA graph of the step")
    p.add_argument("--workload", default="train", choices=["train", "c3", "c3s3", "refine", "sweep"])
    p.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    p.add_argument("--frontend", default="pe", choices=["pe", "hash"], help="sweep only: front-end A (PE + NeFeS MLP) or B (HashGrid + SH, nerfh_tcnn)")
    p.add_argument("--log2T", type=int, default=19, help="sweep --frontend hash: log2 of the hash-table size per level (reference: 19)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--as-world", type=int, default=1, help="train workloads, tuning aid: run the per-GPU share (N_RAND / k rays per image) "
                   "of a k-GPU strong-scaling job on the GPUs present; the line's config states the ray count")
    a = p.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "engine" else a.warmup
    if a.impl == "reference":
        run_reference(a)
    else:
        run_engine(a)


if __name__ == "__main__":
    main()
