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

One JSON line on stdout (rank 0).  `value` times K steps with inputs resident in HBM; `e2e` times the
same steps through the public API with the step's host inputs copied from pinned memory and the loss
read back every step.  `--impl reference` times the CPU oracle port of the reference path on the
host cores (the reference itself is Python and cannot travel to the GPU box).
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
# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch at the bench shape, from the `ncu --set full` captures
# summarised in profiles/ (fine net, 6144 rays x 128 samples); per-launch like roofline.achieved
NCU_TRAFFIC = {"chain_fwd_fine": 2.738e9}


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
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
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
        if self.proc is not None:
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
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
def host_batches(n_steps, seed, pinned=True):
    """Per-step host inputs, as the reference's DataLoader + np.random.choice produce them
    (run_nefes.py:47-71): poses [4,3,4], pixel indices [4,1536], target rgb [6144,3], hist [4,10]."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "poses_stairs.npz"))
    poses = torch.tensor(g["train_gt"][:N_IMAGES].reshape(N_IMAGES, 3, 4), dtype=torch.float32)
    rng = np.random.RandomState(seed)
    pin = (lambda t: t.pin_memory()) if (pinned and torch.cuda.is_available()) else (lambda t: t)
    out = []
    for _ in range(n_steps):
        idx = np.stack([rng.choice(H * W, N_RAND, replace=False) for _ in range(N_IMAGES)])
        out.append(dict(pose=pin(poses.clone()), idx=pin(torch.from_numpy(idx).long()),
                        target=pin(torch.from_numpy(rng.rand(RAYS, 3).astype(np.float32))),
                        hist=pin(torch.zeros(N_IMAGES, 10))))
    return out


def run_engine(a):
    import torch.distributed as dist
    import nefes_b200 as nb
    from nefes_b200 import _lib, ops, parallel
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
    params = [coarse.flat, fine.flat]          # stage 1 trains the two fields (FusionNet joins at stage 3)
    opt = nb.FlatAdam(params, lr=5e-4)

    class Args:
        nerfh_nff, use_fine_only, NeRFW, transient_at_test, netchunk = True, False, True, True, 1 << 21
    q = nb.StandardQuery(Args.netchunk)          # create_nerf's query function: render_rays is one engine call
    kw = dict(network_query_fn=q, N_importance=64, N_samples=64, network_fn=coarse, network_fine=fine,
              use_viewdirs=True, white_bkgd=False, args=Args(), ndc=False, lindisp=False, near=NEAR, far=FAR,
              perturb=1., raw_noise_std=0., test_time=False, retraw=True)

    loss_fn = nb.NerfWLoss(coef=1, lambda_u=0.01)                                # losses.py:96-132 on two kernels

    def step(b):
        """b: dict of DEVICE tensors.  One optimiser step, returns the loss tensor."""
        ro, rd = nb.get_rays_batch(H, W, FOCAL, b["pose"])                      # [4,60,80,3]
        ro = torch.gather(ro.reshape(N_IMAGES, -1, 3), 1, b["idx"][..., None].expand(-1, -1, 3)).reshape(-1, 3)
        rd = torch.gather(rd.reshape(N_IMAGES, -1, 3), 1, b["idx"][..., None].expand(-1, -1, 3)).reshape(-1, 3)
        hist = b["hist"][:, None, :].expand(-1, N_RAND, -1).reshape(-1, 10)
        rgb, disp, acc, ex = nb.render(H, W, FOCAL, chunk=32768, rays=(ro, rd), img_idx=hist, **kw)
        loss = loss_fn({"rgb_coarse": ex["rgb0"], "rgb_fine": rgb, "beta": ex["beta"],
                        "transient_sigmas": ex["transient_sigmas"]}, b["target"])
        opt.zero_grad(set_to_none=True)
        loss.backward()
        if world > 1:
            parallel.allreduce_grads(params)
        opt.step(grad_scale=1.0 / world)
        return loss

    n_total = a.warmup + a.steps
    host = host_batches(n_total, seed=1000 + rank)
    resident = [{k: v.to(dev) for k, v in b.items()} for b in host]
    torch.cuda.synchronize()

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
    with ClockSampler(local) as clk:
        ev0.record()
        for b in resident[a.warmup:]:
            run_step(b)
        ev1.record()
        barrier()
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    launches = launches_per_step * a.steps
    value = world * RAYS * a.steps / (ms / 1e3)

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
    e2e_value = world * RAYS * a.steps / (ms_e2e / 1e3)
    h2d = sum(v.numel() * v.element_size() for v in host[0].values())

    tf_peak, hbm_peak, which = measured_peaks()
    achieved = MLP_FLOP_PER_RAY_FWD_BWD * RAYS / (mlp_ms / 1e3) / 1e12
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
    out = {
        "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if a.precision == "fp32" else "bf16", "data": "synthetic",
        "config": {"workload": "C2 stage-1 colour-only NeRF-W training step, 7-Scenes-stairs camera 640x480 -> 60x80, "
                               "4 images x 1536 rays = 6144 rays/GPU/step, 64 coarse + 64 fine samples, Adam",
                   "rays_per_gpu_per_step": RAYS, "global_rays_per_step": RAYS * world, "parallelism": f"dp{world} (rays sharded, 1 gradient all-reduce)",
                   "mlp_precision": a.precision, "cuda_graph": graph is not None,
                   "l2": "no explicit flush: each step streams >20 GB of activation / gradient tiles through HBM (>> 126 MB L2); "
                         "only the 1.4 MB of weights is legitimately L2-resident"},
        "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / a.steps},
        "gpu_launches": launches,
        # dominant kernel of the step (largest share of device time).  Every MLP kernel of this design streams saved
        # activation / gradient images through HBM and is bounded by it, hence bound = hbm;
        # `traffic` is dram read+write of one launch from the committed ncu capture (profiles/), null if not captured.
        "roofline": ({"bound": "hbm", "kernel": top, "achieved": table[top]["GB_per_s"], "peak": hbm_peak, "unit": "GB/s",
                      "frac": table[top]["hbm_frac"], "traffic": NCU_TRAFFIC.get(top), "peak_source": which,
                      "ms_per_launch": table[top]["ms_per_step"] / table[top]["launches_per_step"],
                      "ms_per_step_in_kernel": table[top]["ms_per_step"],
                      "share_of_step": table[top]["ms_per_step"] / (ms / a.steps)} if top else None),
        "kernels": table,
        "mlp_tensor": {"what": "field MLP (K5) forward+backward, all launches of one step, against the tensor roofline",
                       "achieved": achieved, "peak": tf_peak, "unit": "TFLOP/s", "frac": achieved / tf_peak,
                       "ms_per_step": mlp_ms, "share_of_step": mlp_ms / (ms / a.steps)},
        "clocks": clk.summary(),
        "final_loss": losses[-1] if losses else None,
    }
    if world == 1 and not a.no_extras:
        out["extras"] = side_measurements(a, nb, step, coarse, fine, resident, kw, dev)
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(sample_rays=a.cpu_rays, reps=a.cpu_reps)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        # a process group whose collectives were captured into a CUDA graph does not always tear down cleanly
        # (observed: destroy_process_group / interpreter exit hanging after the result was printed): release the graph,
        # drain the device, meet once more, and leave without running the teardown
        graph = None
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def side_measurements(a, nb, step, coarse, fine, resident, kw, dev):
    """Reported next to the headline, not part of it: the same training step with the fp32 (parity) field path, and
    BASELINE's second metric -- refinement iterations/s (C4: full 60x80 render from a pose, test_time, cosine
    feature loss, backward to the 6 pose parameters, Adam), in both arithmetics."""
    from nefes_b200 import refine
    ex = {}
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if a.precision != "fp32":
        coarse.precision = fine.precision = "fp32"
        for b in resident[:2]:
            step(b)
        torch.cuda.synchronize()
        ev0.record()
        for b in resident[:5]:
            step(b)
        ev1.record()
        torch.cuda.synchronize()
        ex["train_step_fp32_field"] = {"rays_per_s": RAYS * 5 / (ev0.elapsed_time(ev1) / 1e3), "ms_per_step": ev0.elapsed_time(ev1) / 5}
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
                                 "table_MB": float(enc.params.numel() * 4 / 1e6), "note": "fp32 table, T=2^19, L2-resident"}
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
def cpu_baseline(sample_rays=1024, reps=3):
    """The reference path's CPU arithmetic (oracle port, torch CPU fp32, all host threads) on a bounded
    sample of the same step: sample_rays of the 6144 rays, forward + NeRF-W loss + backward + Adam."""
    from oracle import nefes_oracle as O
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    Pc, Pf = O.clone_params(O.init_field("coarse"), requires_grad=True), O.clone_params(O.init_field("fine"), requires_grad=True)
    opt = torch.optim.Adam(list(Pc.values()) + list(Pf.values()), lr=5e-4)
    b = host_batches(1, seed=7, pinned=False)[0]
    ro, rd = O.camera_rays_batch(H, W, FOCAL, b["pose"])
    idx = b["idx"][..., None].expand(-1, -1, 3)
    ro = torch.gather(ro.reshape(N_IMAGES, -1, 3), 1, idx).reshape(-1, 3)[:sample_rays]
    rd = torch.gather(rd.reshape(N_IMAGES, -1, 3), 1, idx).reshape(-1, 3)[:sample_rays]
    times = []
    for i in range(reps + 1):
        t_rand, u = torch.rand(sample_rays, 64), torch.rand(sample_rays, 64)
        t0 = time.perf_counter()
        ret = O.render(H, W, FOCAL, Pc, Pf, rays=(ro, rd), near=NEAR, far=FAR, test_time=False, t_rand=t_rand, u=u)
        loss = nerfw_loss(ret, b["target"][:sample_rays])
        opt.zero_grad()
        loss.backward()
        opt.step()
        times.append(time.perf_counter() - t0)
    med = float(np.median(times[1:]))
    return {"value": sample_rays / med, "unit": "rays/s", "cores": threads, "kind": "port",
            "sample": f"{sample_rays} of the step's 6144 rays, 64+64 samples, fwd+loss+bwd+Adam, median of {reps} after 1 warm-up, "
                      f"torch {torch.__version__} CPU fp32, autograd anomaly mode off (the stock scripts turn it on)",
            "seconds_per_sample": med}


def run_reference(a):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    steps, warm = max(1, a.steps), max(1, a.warmup)
    from oracle import nefes_oracle as O
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    n = a.cpu_rays
    Pc, Pf = O.clone_params(O.init_field("coarse"), requires_grad=True), O.clone_params(O.init_field("fine"), requires_grad=True)
    opt = torch.optim.Adam(list(Pc.values()) + list(Pf.values()), lr=5e-4)
    hb = host_batches(1, seed=7, pinned=False)[0]
    ro, rd = O.camera_rays_batch(H, W, FOCAL, hb["pose"])
    idx = hb["idx"][..., None].expand(-1, -1, 3)
    ro = torch.gather(ro.reshape(N_IMAGES, -1, 3), 1, idx).reshape(-1, 3)[:n]
    rd = torch.gather(rd.reshape(N_IMAGES, -1, 3), 1, idx).reshape(-1, 3)[:n]
    # keep the whole run within a few minutes: cap the number of timed steps
    budget_s, t_first = 150.0, None
    times = []
    for i in range(warm + steps):
        t_rand, u = torch.rand(n, 64), torch.rand(n, 64)
        t0 = time.perf_counter()
        ret = O.render(H, W, FOCAL, Pc, Pf, rays=(ro, rd), near=NEAR, far=FAR, test_time=False, t_rand=t_rand, u=u)
        loss = nerfw_loss(ret, hb["target"][:n])
        opt.zero_grad()
        loss.backward()
        opt.step()
        dt = time.perf_counter() - t0
        if i >= warm:
            times.append(dt)
        if sum(times) > budget_s:
            break
    k = len(times)
    tot = sum(times)
    v = n * k / tot
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": "rays/s", "n_gpus": a.gpus, "steps": k,
           "warmup": warm, "ms_per_step": 1e3 * tot / k, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": {"workload": "C2 stage-1 colour-only NeRF-W training step (same as the engine arm); each step is a bounded "
                                  f"sample of {n} of the 6144 rays", "rays_per_step_sample": n},
           "cpu_baseline": {"value": v, "unit": "rays/s", "cores": threads, "kind": "port",
                            "sample": f"{n} rays/step x {k} steps, oracle port of the reference path (the reference is Python "
                                      "and absent on the GPU box), torch CPU fp32, all host threads"},
           "e2e": {"value": v, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="engine", choices=["engine", "reference"])
    p.add_argument("--precision", default=os.environ.get("NEFES_PRECISION", "bf16"), choices=["fp32", "bf16"])
    p.add_argument("--no-extras", action="store_true", help="skip the fp32-path and refinement side measurements")
    p.add_argument("--no-graph", action="store_true", help="time eager steps instead of a captured CUDA graph of the step")
    p.add_argument("--cpu-rays", type=int, default=1024)
    p.add_argument("--cpu-reps", type=int, default=3)
    p.add_argument("--no-cpu-baseline", action="store_true")
    a = p.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "engine" else a.warmup
    if a.impl == "reference":
        run_reference(a)
    else:
        run_engine(a)


if __name__ == "__main__":
    main()
