"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from
/root/reference) and the oracle restatement on identical seeded inputs.

Run in the build container only (the GPU box has no /root/reference):
    python oracle/make_golden.py
Every vector written here is first asserted bit-equal (CPU fp32) between the reference
and oracle/nefes_oracle.py, so the committed fixtures pin both.  Test infrastructure.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)


def import_reference():
    """The unmodified reference modules, third-party imports stubbed (oracle/ref_loader.py)."""
    from oracle import ref_loader
    return ref_loader.import_reference(REF)


def eq(name, a, b):
    a, b = a.detach(), b.detach()
    if not torch.equal(a, b):
        raise SystemExit(f"[make_golden] oracle != reference for {name}: "
                         f"max abs diff {(a - b).abs().max().item():.3e}")


def close(name, a, b, rtol=2e-5, atol=1e-7):
    """Gradients: autograd accumulates in graph order, which a restatement cannot pin to
    the last ulp; forward values are bit-equal, gradients are checked to ~1e-5 relative."""
    a, b = a.detach(), b.detach()
    if not torch.allclose(a, b, rtol=rtol, atol=atol * float(b.abs().max())):
        raise SystemExit(f"[make_golden] oracle !~ reference for {name}: "
                         f"max abs diff {(a - b).abs().max().item():.3e} (scale {b.abs().max().item():.3e})")


def npy(d):
    return {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v))
            for k, v in d.items() if v is not None}


def main():
    from oracle import nefes_oracle as O
    R, M, U = import_reference()
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)

    # ---- pose fixtures (data rows, 7-Scenes stairs) ----------------------
    pr = os.path.join(REF, "paper_result/DFNet_NeFeS50_7Scenes_colmap/stairs")
    gt = np.loadtxt(os.path.join(pr, "stairs_test_gt.txt"))[:64].astype(np.float64)
    init = np.loadtxt(os.path.join(pr, "DFNet_stairs_results.txt"))[:64].astype(np.float64)
    train = np.loadtxt(os.path.join(pr, "stairs_train_gt.txt"))[:16].astype(np.float64)
    np.savez_compressed(os.path.join(OUT, "poses_stairs.npz"), test_gt=gt, dfnet_init=init, train_gt=train)
    H, W, focal, near, far = 60, 80, 525.505 / 2 / 4, 0., 4.
    c2w = torch.tensor(gt[0].reshape(3, 4), dtype=torch.float32)

    # ---- models: reference constructor vs oracle init -----------------------------------
    kw = dict(D=8, W=128, skips=[4], in_channels_xyz=63, in_channels_dir=27)
    ref_c = M.NeRFH_NFF("coarse", **kw)
    ref_f = M.NeRFH_NFF("fine", encode_appearance=True, encode_transient=True, **kw)
    Pc, Pf = O.init_field("coarse"), O.init_field("fine")
    sd_c, sd_f = ref_c.state_dict(), ref_f.state_dict()
    for k in Pc:
        eq("coarse." + k, Pc[k], sd_c[k])
    for k in Pf:
        eq("fine." + k, Pf[k], sd_f[k])
    np.savez_compressed(os.path.join(OUT, "weights.npz"),
                        **{"coarse/" + k: v.numpy() for k, v in Pc.items()},
                        **{"fine/" + k: v.numpy() for k, v in Pf.items()})

    # ---- G1: get_rays / get_rays_batch ---------------------------------------------------
    ro, rd = U.get_rays(H, W, focal, c2w)
    oo, od = O.camera_rays(H, W, focal, c2w)
    eq("rays_o", oo, ro), eq("rays_d", od, rd)
    c2w_b = torch.tensor(train[:4].reshape(4, 3, 4), dtype=torch.float32)
    rob, rdb = U.get_rays_batch(H, W, focal, c2w_b)
    oob, odb = O.camera_rays_batch(H, W, focal, c2w_b)
    eq("rays_o_b", oob, rob), eq("rays_d_b", odb, rdb)
    np.savez_compressed(os.path.join(OUT, "g1_rays.npz"), H=H, W=W, focal=focal, c2w=c2w.numpy(),
                        rays_o=ro.numpy(), rays_d=rd.numpy(), c2w_b=c2w_b.numpy(),
                        rays_d_b=rdb.numpy(), rays_o_b=rob.numpy())

    # ---- G2: sample_pdf (stage-level known-answer: bins, weights, u -> samples, inds) ----
    g = torch.Generator().manual_seed(11)
    n = 96
    zc = O.coarse_depths(torch.zeros(n, 1), 4 * torch.ones(n, 1), 64, torch.rand(n, 64, generator=g))
    bins = .5 * (zc[:, 1:] + zc[:, :-1])
    wts = torch.rand(n, 62, generator=g) ** 4
    wts[:8] = 0                                  # all-epsilon rows
    wts[8:16, 5:] = 0                            # mass in the first bins only
    wts[16:24, :50] = 0                          # mass at the tail
    u = torch.rand(n, 64, generator=g)
    torch.manual_seed(5)
    ref_rand = R.sample_pdf(bins, wts, 64, det=False)          # draws torch.rand itself
    torch.manual_seed(5)
    u_ref = torch.rand(n, 64)
    s_rand, i_rand, cdf = O.importance_depths(bins, wts, 64, u_ref)
    eq("sample_pdf(rand)", s_rand, ref_rand)
    ref_det = R.sample_pdf(bins, wts, 64, det=True)
    s_det, i_det, _ = O.importance_depths(bins, wts, 64, None)
    eq("sample_pdf(det)", s_det, ref_det)
    ref_py = R.sample_pdf(bins, wts, 64, det=False, pytest=True)   # the reference's own determinism hook
    np.random.seed(0)
    u_py = torch.Tensor(np.random.rand(n, 64))
    s_py, i_py, _ = O.importance_depths(bins, wts, 64, u_py)
    eq("sample_pdf(pytest)", s_py, ref_py)
    np.savez_compressed(os.path.join(OUT, "g2_sample_pdf.npz"), bins=bins.numpy(), weights=wts.numpy(),
                        cdf=cdf.numpy(), u_rand=u_ref.numpy(), samples_rand=s_rand.numpy(),
                        inds_rand=i_rand.numpy().astype(np.int32), samples_det=s_det.numpy(),
                        inds_det=i_det.numpy().astype(np.int32), u_pytest=u_py.numpy(),
                        samples_pytest=s_py.numpy(), inds_pytest=i_py.numpy().astype(np.int32),
                        u_det=torch.linspace(0., 1., 64).numpy(), t_vals=torch.linspace(0., 1., 64).numpy())

    # ---- G3: raw2outputs, every mode, forward + gradient w.r.t. raw ----------------------
    g = torch.Generator().manual_seed(3)
    nr = 4
    z64 = torch.sort(torch.rand(nr, 64, generator=g) * 4, -1)[0]
    z128 = torch.sort(torch.rand(nr, 128, generator=g) * 4, -1)[0]

    def mk_raw(s, c, trans):
        raw = torch.randn(nr, s, c, generator=g)
        raw[..., 131] = torch.nn.functional.softplus(raw[..., 131] * 3)        # sigma >= 0
        if trans:
            raw[..., 132:135] = torch.sigmoid(raw[..., 132:135])
            raw[..., 135:137] = torch.nn.functional.softplus(raw[..., 135:137])
        return raw
    g3 = {"z64": z64, "z128": z128}
    cases = {
        "coarse_train": dict(s=64, c=132, kw=dict(typ="coarse", test_time=False)),
        "fine_train": dict(s=128, c=137, kw=dict(typ="fine", test_time=False, output_transient=True,
                                                 beta_min=0.1, transient_at_test=True)),
        "fine_test_tat": dict(s=128, c=137, kw=dict(typ="fine", test_time=True, output_transient=True,
                                                    beta_min=0.1, transient_at_test=True)),
        "fine_test_static": dict(s=128, c=137, kw=dict(typ="fine", test_time=True, output_transient=True,
                                                       beta_min=0.1, transient_at_test=False)),
        "fine_notransient": dict(s=128, c=132, kw=dict(typ="fine", test_time=False)),
    }
    names = ("rgb", "feat", "disp", "acc", "weights", "depth", "transient_sigmas", "beta")
    for cname, cs in cases.items():
        trans = cs["c"] == 137
        raw = mk_raw(cs["s"], cs["c"], trans)
        z = z64 if cs["s"] == 64 else z128
        outs_ref, outs_orc, grads = [], [], []
        for impl in ("ref", "orc"):
            r = raw.clone().requires_grad_(True)
            if impl == "ref":
                torch.manual_seed(0)
                tup = M.raw2outputs_NeRFH_NFF(r, z, raw_noise_std=0, **cs["kw"])
            else:
                tup = O.composite(r, z, **cs["kw"]).astuple()
            # scalar probe touching every differentiable output with fixed pseudo-random cotangents
            gg = torch.Generator().manual_seed(17)
            loss = 0
            for t in tup:
                if t is not None and t.requires_grad:
                    loss = loss + (t * torch.randn(t.shape, generator=gg)).sum()
            loss.backward()
            (outs_ref if impl == "ref" else outs_orc).extend(tup)
            grads.append(r.grad.clone())
        for nme, a, b in zip(names, outs_orc, outs_ref):
            if b is None:
                assert a is None, (cname, nme)
                continue
            eq(f"composite[{cname}].{nme}", a, b)
            g3[f"{cname}/{nme}"] = b
        close(f"composite[{cname}].d_raw", grads[1], grads[0])
        g3[f"{cname}/raw"] = raw
        g3[f"{cname}/d_raw"] = grads[0]
    # coarse test: sigma-only raw [N,64,1]
    raw1 = torch.nn.functional.softplus(torch.randn(nr, 64, 1, generator=g) * 3)
    tup = M.raw2outputs_NeRFH_NFF(raw1, z64, typ="coarse", test_time=True)
    oc = O.composite(raw1, z64, typ="coarse", test_time=True)
    eq("composite[coarse_test].acc", oc.acc, tup[3]), eq("composite[coarse_test].weights", oc.weights, tup[4])
    g3.update({"coarse_test/raw": raw1, "coarse_test/acc": tup[3], "coarse_test/weights": tup[4]})
    np.savez_compressed(os.path.join(OUT, "g3_composite.npz"), **npy(g3))

    # ---- G4: the MLP (NeRFH_NFF.forward) on random embedded inputs -----------------------
    g = torch.Generator().manual_seed(4)
    xin = torch.cat([torch.rand(200, 3, generator=g) * 4 - 2, torch.nn.functional.normalize(
        torch.randn(200, 3, generator=g), dim=-1)], -1)
    emb = torch.cat([O.freq_encode(xin[:, :3], 10), O.freq_encode(xin[:, 3:], 4)], -1)
    embed_fn, in_ch, _ = M.get_embedder(10, 0, -1)
    embeddirs_fn, in_ch_d, _ = M.get_embedder(4, 0, -1)
    eq("embed xyz", emb[:, :63], embed_fn(xin[:, :3])), eq("embed dir", emb[:, 63:], embeddirs_fn(xin[:, 3:]))
    with torch.no_grad():
        r_sig = ref_c(emb[:, :63], sigma_only=True)
        r_sta = ref_c(emb, output_transient=False)
        r_ful = ref_f(emb, output_transient=True)
        eq("mlp sigma", O.field_forward(Pc, emb[:, :63], mode="sigma"), r_sig)
        eq("mlp static", O.field_forward(Pc, emb[:, :63], emb[:, 63:], "static"), r_sta)
        eq("mlp full", O.field_forward(Pf, emb[:, :63], emb[:, 63:], "full"), r_ful)
    np.savez_compressed(os.path.join(OUT, "g4_mlp.npz"), xyz=xin[:, :3].numpy(), dirs=xin[:, 3:].numpy(),
                        emb=emb.numpy(), sigma=r_sig.numpy(), static=r_sta.numpy(), full=r_ful.numpy())

    # ---- G5: end-to-end render(), train + test mode, forward + gradients -----------------
    class Args:
        nerfh_nff = True
        use_fine_only = False
        NeRFW = True
        transient_at_test = True
    query = lambda inputs, viewdirs, ts, fn, typ, output_transient, test_time, store_rgb: \
        M.run_network_NeRFH_NFF(inputs, viewdirs, ts, fn, embed_fn=embed_fn, embeddirs_fn=embeddirs_fn,
                                typ=typ, output_transient=output_transient, netchunk=1 << 21,
                                test_time=test_time, store_rgb=store_rgb)
    base = dict(network_query_fn=query, N_importance=64, N_samples=64, network_fn=ref_c,
                network_fine=ref_f, use_viewdirs=True, white_bkgd=False, args=Args(), ndc=False,
                lindisp=False, near=near, far=far)
    pix = torch.randperm(H * W, generator=torch.Generator().manual_seed(9))[:48]
    hist = torch.zeros(1, 10)
    rays = (ro.reshape(-1, 3)[pix], rd.reshape(-1, 3)[pix])
    g5 = {"pix": pix, "rays_o": rays[0], "rays_d": rays[1]}

    # train mode: weight gradients
    probe_keys = ["xyz_encoding_1.0.weight", "xyz_encoding_5.0.weight", "xyz_encoding_8.0.bias",
                  "static_sigma.0.weight", "static_rgb.0.weight", "dir_encoding.0.weight"]
    fine_only = ["transient_encoding.0.weight", "transient_beta.0.bias", "transient_rgb.0.weight"]
    tgt = torch.rand(48, 3, generator=torch.Generator().manual_seed(21))
    ref_c.zero_grad(), ref_f.zero_grad()
    torch.manual_seed(123)
    rgb, disp, acc, ex = R.render(H, W, focal, chunk=32768, rays=rays, img_idx=hist, perturb=1.0,
                                  raw_noise_std=0., test_time=False, retraw=True, **base)
    out_ref = dict(rgb_map=rgb, disp_map=disp, acc_map=acc, **ex)
    loss_ref = O.nerfw_loss(out_ref, tgt) + 0.04 * (out_ref["feat_map"].abs().mean() + out_ref["feat0"].abs().mean())
    loss_ref.backward()
    t_rand, u_tr = O.draw_train_randoms(48, seed=123)
    Pc_g, Pf_g = O.clone_params(Pc, requires_grad=True), O.clone_params(Pf, requires_grad=True)
    out_orc = O.render(H, W, focal, Pc_g, Pf_g, rays=rays, near=near, far=far, hist=hist, test_time=False,
                       t_rand=t_rand, u=u_tr, return_aux=True)
    loss_orc = O.nerfw_loss(out_orc, tgt) + 0.04 * (out_orc["feat_map"].abs().mean() + out_orc["feat0"].abs().mean())
    loss_orc.backward()
    for k in out_ref:
        eq("render[train]." + k, out_orc[k], out_ref[k])
        g5["train/" + k] = out_ref[k]
    eq("render[train].loss", loss_orc, loss_ref)
    for k in probe_keys:
        close("grad coarse " + k, Pc_g[k].grad, dict(ref_c.named_parameters())[k].grad)
        close("grad fine " + k, Pf_g[k].grad, dict(ref_f.named_parameters())[k].grad)
        g5["train/grad_coarse/" + k] = Pc_g[k].grad
        g5["train/grad_fine/" + k] = Pf_g[k].grad
    for k in fine_only:
        close("grad fine " + k, Pf_g[k].grad, dict(ref_f.named_parameters())[k].grad)
        g5["train/grad_fine/" + k] = Pf_g[k].grad
    g5.update({"train/t_rand": t_rand, "train/u": u_tr, "train/target": tgt, "train/loss": loss_ref,
               "train/inds": out_orc["_aux"]["inds"].to(torch.int32), "train/z_fine": out_orc["_aux"]["z_fine"],
               "train/z_coarse": out_orc["_aux"]["z_coarse"], "train/z_samples": out_orc["_aux"]["z_samples"]})

    # test mode (refinement): gradient of a feature-cosine + colour probe w.r.t. the pose
    for p in list(ref_c.parameters()) + list(ref_f.parameters()):
        p.requires_grad_(False)
    feat_t = torch.randn(128, H * W // 50, generator=torch.Generator().manual_seed(31))
    sub = torch.arange(0, H * W, 50)
    c2w_r = c2w.clone().requires_grad_(True)
    rgb, disp, acc, ex = R.render(H, W, focal, chunk=32768, c2w=c2w_r, img_idx=hist, perturb=False,
                                  raw_noise_std=0., test_time=True, **base)
    l_ref = O.cosine_feature_loss(ex["feat_map"][sub].t(), feat_t) + rgb[sub].mean()
    l_ref.backward()
    c2w_o = c2w.clone().requires_grad_(True)
    out_o = O.render(H, W, focal, Pc, Pf, c2w=c2w_o, near=near, far=far, hist=hist, test_time=True,
                     return_aux=True)
    l_orc = O.cosine_feature_loss(out_o["feat_map"][sub].t(), feat_t) + out_o["rgb_map"][sub].mean()
    l_orc.backward()
    for k, v in dict(rgb_map=rgb, disp_map=disp, acc_map=acc, **ex).items():
        eq("render[test]." + k, out_o[k], v)
        g5["test/" + k] = v[sub]
    close("render[test].d_c2w", c2w_o.grad, c2w_r.grad)
    g5.update({"test/sub": sub, "test/feat_target": feat_t, "test/loss": l_ref, "test/d_c2w": c2w_r.grad,
               "test/inds": out_o["_aux"]["inds"][sub].to(torch.int32),
               "test/z_fine": out_o["_aux"]["z_fine"][sub], "test/c2w": c2w})
    np.savez_compressed(os.path.join(OUT, "g5_render.npz"), **npy(g5))

    # ---- G7: the two dormant options of render_rays, lindisp (rendering.py:97-100) and white_bkgd (nerfh_nff.py:126-127) ----
    for p_ in list(ref_c.parameters()) + list(ref_f.parameters()):
        p_.requires_grad_(True)
    g7 = {}
    rays7 = (ro.reshape(-1, 3)[pix[:32]], rd.reshape(-1, 3)[pix[:32]])
    base7 = dict(base)
    base7.update(near=0.5, far=4., lindisp=True, white_bkgd=True)
    torch.manual_seed(77)
    rgb, disp, acc, ex = R.render(H, W, focal, chunk=32768, rays=rays7, img_idx=hist, perturb=1.0, raw_noise_std=0., test_time=False,
                                  retraw=True, **base7)
    t_rand7, u7 = O.draw_train_randoms(32, seed=77)
    out7 = O.render(H, W, focal, Pc, Pf, rays=rays7, near=0.5, far=4., hist=hist, test_time=False, t_rand=t_rand7, u=u7,
                    lindisp=True, white_bkgd=True)
    for k, v in dict(rgb_map=rgb, disp_map=disp, acc_map=acc, **ex).items():
        eq("render[lindisp, white_bkgd]." + k, out7[k], v)
        g7["train/" + k] = v
    g7.update({"rays_o": rays7[0], "rays_d": rays7[1], "t_rand": t_rand7, "u": u7})
    np.savez_compressed(os.path.join(OUT, "g7_options.npz"), **npy(g7))

    # ---- G8: FusionNet (nerfh_nff.py:356-418) through run_fusion_net (:578-603), training and eval mode, forward + backward ----
    import nefes_b200.nerfh_nff as NB
    torch.manual_seed(5)
    ref_fn = M.FusionNet(128)
    torch.manual_seed(5)
    my_fn = NB.FusionNet(128)
    for (ka, va), (kb, vb) in zip(ref_fn.state_dict().items(), my_fn.state_dict().items()):
        assert ka == kb, (ka, kb)
        eq("FusionNet init " + ka, vb, va)
    g = torch.Generator().manual_seed(8)
    Bf, Hf, Wf = 2, 16, 16
    rgb8, feat8 = torch.rand(Bf * Hf * Wf, 3, generator=g), torch.randn(Bf * Hf * Wf, 128, generator=g)
    cot8 = torch.randn(Bf, 128, Hf, Wf, generator=g)
    with torch.no_grad():                                  # non-trivial BatchNorm affine and running statistics
        ref_fn.net[7].weight.copy_(torch.rand(128, generator=g) + 0.5)
        ref_fn.net[7].bias.copy_(torch.randn(128, generator=g) * 0.1)
        ref_fn.net[7].running_mean.copy_(torch.randn(128, generator=g) * 0.05)
        ref_fn.net[7].running_var.copy_(torch.rand(128, generator=g) * 0.5 + 0.1)
    bn_init = {k: v.clone() for k, v in ref_fn.net[7].state_dict().items()}
    holder = types.SimpleNamespace(fusion_net=ref_fn, W_features=128)
    g8 = {"rgb": rgb8, "feat": feat8, "cot": cot8, "B": Bf, "H": Hf, "W": Wf}
    g8.update({"bn/" + k: v for k, v in bn_init.items()})
    g8["w_checksum"] = torch.stack([v.double().abs().sum() for k, v in ref_fn.state_dict().items() if k.endswith("weight")])
    for mode in ("train", "eval"):
        ref_fn.net[7].load_state_dict(bn_init)
        ref_fn.train(mode == "train")
        ref_fn.zero_grad()
        a, b = rgb8.clone().requires_grad_(True), feat8.clone().requires_grad_(True)
        _, _, out_r = M.NeRFH_NFF.run_fusion_net(holder, a, b, Hf, Wf, Bf)
        (out_r * cot8).sum().backward()
        Pfn = {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point and "running" not in k) for k, v in ref_fn.state_dict().items()}
        Pfn["net.7.running_mean"], Pfn["net.7.running_var"] = bn_init["running_mean"].clone(), bn_init["running_var"].clone()
        a2, b2 = rgb8.clone().requires_grad_(True), feat8.clone().requires_grad_(True)
        out_o = O.fusion_net(Pfn, a2, b2, Bf, Hf, Wf, training=(mode == "train"))
        (out_o * cot8).sum().backward()
        eq(f"fusion_net[{mode}]", out_o, out_r)
        close(f"fusion_net[{mode}] d_rgb", a2.grad, a.grad)
        close(f"fusion_net[{mode}] d_feat", b2.grad, b.grad)
        g8[f"{mode}/out"], g8[f"{mode}/d_rgb"], g8[f"{mode}/d_feat"] = out_r, a.grad, b.grad
        for k, v in ref_fn.named_parameters():
            close(f"fusion_net[{mode}] grad {k}", Pfn[k].grad, v.grad, rtol=1e-4, atol=1e-6)
            g8[f"{mode}/grad/{k}"] = v.grad.reshape(-1)[::37] if v.grad.numel() > 4096 else v.grad
        if mode == "train":
            eq("fusion_net running_mean", Pfn["net.7.running_mean"], ref_fn.net[7].running_mean)
            eq("fusion_net running_var", Pfn["net.7.running_var"], ref_fn.net[7].running_var)
            g8["train/running_mean"], g8["train/running_var"] = ref_fn.net[7].running_mean.clone(), ref_fn.net[7].running_var.clone()
    np.savez_compressed(os.path.join(OUT, "g8_fusion.npz"), **npy(g8))

    # ---- G6: stage-2/3 loss, ColorFeatureFusionNerfWLoss (losses.py:134-173), L1 and MSE feature terms -------------
    import models.losses as RLoss
    g = torch.Generator().manual_seed(6)
    n = 96
    res = {"rgb_fine": torch.rand(n, 3, generator=g), "rgb_coarse": torch.rand(n, 3, generator=g),
           "beta": torch.rand(n, generator=g) + 0.1, "transient_sigmas": torch.rand(n, 128, generator=g),
           "feat_fine": torch.randn(n, 128, generator=g), "feat_coarse": torch.randn(n, 128, generator=g),
           "feat_fusion": torch.randn(n, 128, generator=g)}
    tg = {"rgb": torch.rand(n, 3, generator=g), "feat": torch.randn(n, 128, generator=g)}
    tg["feat"][:4] = res["feat_fine"][:4]                     # exact zeros of a - t: sign(0) = 0 in the L1 gradient
    g6 = {"in/" + k: v for k, v in res.items()}
    g6.update({"target/" + k: v for k, v in tg.items()})
    for l1 in (True, False):
        lf = RLoss.ColorFeatureFusionNerfWLoss(coef=1, L1_loss=l1)
        for tag, kw_ in (("color", dict(switch_on=False, color_only_switch=True)), ("stage2", dict(switch_on=False, color_only_switch=False)),
                         ("stage3", dict(switch_on=True, color_only_switch=False))):
            leaf = {k: v.clone().requires_grad_(True) for k, v in res.items()}
            out_r = lf(leaf, tg, **kw_)
            out_o = O.color_feature_fusion_nerfw_loss(res, tg, L1_loss=l1, **kw_)
            out_r = out_r if isinstance(out_r, tuple) else (out_r,)
            out_o = out_o if isinstance(out_o, tuple) else (out_o,)
            assert len(out_r) == len(out_o)
            for i, (a, b) in enumerate(zip(out_o, out_r)):
                eq(f"loss[{'l1' if l1 else 'mse'}/{tag}][{i}]", a, b)
                g6[f"{'l1' if l1 else 'mse'}/{tag}/{i}"] = b
            w = (1.0, 0.04, 0.02)
            sum(wi * li for wi, li in zip(w, out_r)).backward()   # the caller's weighting (run_nefes.py:238-251)
            for k in ("feat_fine", "feat_coarse", "feat_fusion"):
                if leaf[k].grad is not None:
                    g6[f"{'l1' if l1 else 'mse'}/{tag}/grad/{k}"] = leaf[k].grad
    np.savez_compressed(os.path.join(OUT, "g6_loss.npz"), **npy(g6))

    tot = sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT))
    print(f"[make_golden] all reference-vs-oracle checks bit-equal; wrote {tot / 1e6:.2f} MB to {OUT}")


if __name__ == "__main__":
    main()
