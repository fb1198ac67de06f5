"""CPU: the oracle restatement against the committed golden vectors (which were produced by
the unmodified reference, see oracle/make_golden.py).  Forward values bit-exact."""
import torch

from oracle import nefes_oracle as O


def test_init_matches_reference_constructor(weights):
    wc, wf = weights
    pc, pf = O.init_field("coarse"), O.init_field("fine")
    assert set(pc) == set(wc) and set(pf) == set(wf)
    for k in pc:
        assert torch.equal(pc[k], wc[k]), k
    for k in pf:
        assert torch.equal(pf[k], wf[k]), k
    assert sum(v.numel() for v in pf.values()) == 185609


def test_rays(golden):
    g = golden("g1_rays.npz")
    o, d = O.camera_rays(int(g["H"]), int(g["W"]), float(g["focal"]), g["c2w"])
    assert torch.equal(o, g["rays_o"]) and torch.equal(d, g["rays_d"])
    ob, db = O.camera_rays_batch(int(g["H"]), int(g["W"]), float(g["focal"]), g["c2w_b"])
    assert torch.equal(ob, g["rays_o_b"]) and torch.equal(db, g["rays_d_b"])


def test_sample_pdf_known_answers(golden):
    g = golden("g2_sample_pdf.npz")
    for tag, u in (("rand", g["u_rand"]), ("det", None), ("pytest", g["u_pytest"])):
        s, inds, cdf = O.importance_depths(g["bins"], g["weights"], 64, u)
        assert torch.equal(s, g["samples_" + tag]), tag
        assert torch.equal(inds.to(torch.int32), g["inds_" + tag]), tag
        assert torch.equal(cdf, g["cdf"])
    assert torch.equal(torch.linspace(0., 1., 64), g["t_vals"])


CASES = {
    "coarse_train": dict(typ="coarse", test_time=False),
    "fine_train": dict(typ="fine", test_time=False, output_transient=True, transient_at_test=True),
    "fine_test_tat": dict(typ="fine", test_time=True, output_transient=True, transient_at_test=True),
    "fine_test_static": dict(typ="fine", test_time=True, output_transient=True, transient_at_test=False),
    "fine_notransient": dict(typ="fine", test_time=False),
}
NAMES = ("rgb", "feat", "disp", "acc", "weights", "depth", "transient_sigmas", "beta")


def test_composite_all_modes(golden):
    g = golden("g3_composite.npz")
    for case, kw in CASES.items():
        raw = g[case + "/raw"].clone().requires_grad_(True)
        z = g["z64"] if raw.shape[1] == 64 else g["z128"]
        out = O.composite(raw, z, **kw).astuple()
        gg = torch.Generator().manual_seed(17)
        loss = 0
        for name, t in zip(NAMES, out):
            key = f"{case}/{name}"
            if key in g:
                assert torch.equal(t.detach(), g[key]), key
            if t is not None and t.requires_grad:
                loss = loss + (t * torch.randn(t.shape, generator=gg)).sum()
        loss.backward()
        ref = g[case + "/d_raw"]
        assert torch.allclose(raw.grad, ref, rtol=2e-5, atol=1e-7 * float(ref.abs().max())), case
    c = O.composite(g["coarse_test/raw"], g["z64"], typ="coarse", test_time=True)
    assert c.rgb is None and torch.equal(c.weights, g["coarse_test/weights"])


def test_mlp(golden, weights):
    g = golden("g4_mlp.npz")
    wc, wf = weights
    emb = torch.cat([O.freq_encode(g["xyz"], 10), O.freq_encode(g["dirs"], 4)], -1)
    assert torch.equal(emb, g["emb"])
    with torch.no_grad():
        assert torch.equal(O.field_forward(wc, emb[:, :63], mode="sigma"), g["sigma"])
        assert torch.equal(O.field_forward(wc, emb[:, :63], emb[:, 63:], "static"), g["static"])
        assert torch.equal(O.field_forward(wf, emb[:, :63], emb[:, 63:], "full"), g["full"])


def test_render_train_and_test_mode(golden, weights):
    g = golden("g5_render.npz")
    wc, wf = weights
    H, W, focal = 60, 80, 525.505 / 2 / 4
    pc, pf = O.clone_params(wc, requires_grad=True), O.clone_params(wf, requires_grad=True)
    out = O.render(H, W, focal, pc, pf, rays=(g["rays_o"], g["rays_d"]), near=0., far=4.,
                   test_time=False, t_rand=g["train/t_rand"], u=g["train/u"], return_aux=True)
    for k in ("rgb_map", "disp_map", "acc_map", "feat_map", "rgb0", "disp0", "acc0", "z_std",
              "transient_sigmas", "beta", "feat0"):
        assert torch.equal(out[k].detach(), g["train/" + k]), k
    assert torch.equal(out["_aux"]["inds"].to(torch.int32), g["train/inds"])
    assert torch.equal(out["_aux"]["z_fine"], g["train/z_fine"])
    loss = O.nerfw_loss(out, g["train/target"]) + 0.04 * (out["feat_map"].abs().mean() + out["feat0"].abs().mean())
    assert torch.equal(loss.detach(), g["train/loss"])
    loss.backward()
    for k, v in g.items():
        if k.startswith("train/grad_fine/"):
            got = pf[k[len("train/grad_fine/"):]].grad
            assert torch.allclose(got, v, rtol=2e-5, atol=1e-7 * float(v.abs().max())), k
    # refinement mode: gradient to the camera pose
    c2w = g["test/c2w"].clone().requires_grad_(True)
    out = O.render(H, W, focal, wc, wf, c2w=c2w, near=0., far=4., test_time=True, return_aux=True)
    sub = g["test/sub"]
    assert torch.equal(out["feat_map"][sub].detach(), g["test/feat_map"])
    assert torch.equal(out["_aux"]["inds"][sub].to(torch.int32), g["test/inds"])
    l = O.cosine_feature_loss(out["feat_map"][sub].t(), g["test/feat_target"]) + out["rgb_map"][sub].mean()
    l.backward()
    ref = g["test/d_c2w"]
    assert torch.allclose(c2w.grad, ref, rtol=2e-5, atol=1e-7 * float(ref.abs().max()))


def test_stage23_loss_golden(golden):
    """ColorFeatureFusionNerfWLoss (losses.py:134-173): values bit-equal to the reference class in all three call modes."""
    g = golden("g6_loss.npz")
    res = {k[3:]: v for k, v in g.items() if k.startswith("in/")}
    tg = {k[7:]: v for k, v in g.items() if k.startswith("target/")}
    for l1 in (True, False):
        for tag, kw in (("color", dict(switch_on=False, color_only_switch=True)), ("stage2", dict(switch_on=False, color_only_switch=False)),
                        ("stage3", dict(switch_on=True, color_only_switch=False))):
            out = O.color_feature_fusion_nerfw_loss(res, tg, L1_loss=l1, **kw)
            out = out if isinstance(out, tuple) else (out,)
            for i, v in enumerate(out):
                assert torch.equal(v, g[f"{'l1' if l1 else 'mse'}/{tag}/{i}"]), (l1, tag, i)


def test_lindisp_and_white_bkgd_golden(golden, weights):
    """The oracle's two dormant render options (rendering.py:97-100 lindisp, nerfh_nff.py:126-127 white_bkgd) against the
    fixture rendered by the unmodified reference (G7)."""
    g = golden("g7_options.npz")
    wc, wf = weights
    out = O.render(60, 80, 525.505 / 2 / 4, wc, wf, rays=(g["rays_o"], g["rays_d"]), near=0.5, far=4., test_time=False,
                   t_rand=g["t_rand"], u=g["u"], lindisp=True, white_bkgd=True)
    for k in ("rgb_map", "disp_map", "acc_map", "feat_map", "rgb0", "disp0", "acc0", "z_std", "transient_sigmas", "beta", "feat0"):
        assert torch.equal(out[k], g["train/" + k]), k


def test_fusion_net_golden(golden):
    """The oracle's FusionNet restatement (nerfh_nff.py:356-418, :578-603) against the reference module's outputs and input
    gradients (G8), training and eval mode.  The weights are the seeded default init of the constructor (checksum in the
    fixture); BatchNorm state from the fixture."""
    import nefes_b200.nerfh_nff as NB
    g = golden("g8_fusion.npz")
    B, H, W = int(g["B"]), int(g["H"]), int(g["W"])
    torch.manual_seed(5)
    m = NB.FusionNet(128)
    m.net[7].load_state_dict({k[3:]: v for k, v in g.items() if k.startswith("bn/")})
    chk = torch.stack([v.double().abs().sum() for k, v in m.state_dict().items() if k.endswith("weight")])
    assert float((chk - g["w_checksum"]).abs().max()) < 1e-9
    for mode in ("train", "eval"):
        P = {k: v.detach().clone() for k, v in m.state_dict().items()}
        rgb, feat = g["rgb"].clone().requires_grad_(True), g["feat"].clone().requires_grad_(True)
        out = O.fusion_net(P, rgb, feat, B, H, W, training=(mode == "train"))
        scale = float(g[f"{mode}/out"].abs().max())
        assert float((out.detach() - g[f"{mode}/out"]).abs().max()) < 1e-5 * scale, mode
        (out * g["cot"]).sum().backward()
        for got, want in ((rgb.grad, g[f"{mode}/d_rgb"]), (feat.grad, g[f"{mode}/d_feat"])):
            assert float((got - want).abs().max()) < 1e-4 * float(want.abs().max()), mode
        if mode == "train":
            assert torch.allclose(P["net.7.running_mean"], g["train/running_mean"], rtol=1e-5, atol=1e-7)
            assert torch.allclose(P["net.7.running_var"], g["train/running_var"], rtol=1e-5, atol=1e-7)
