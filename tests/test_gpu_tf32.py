"""GPU: the tf32 tensor-core path (NEFES_PREC_TF32): tcgen05 kind::tf32 GEMMs under the fp32 path's layer-at-a-time structure.
(1) the GEMM itself through nefes_linear_{fwd,dgrad,wgrad} with nefes_gemm_mode(1) against torch fp32 on awkward shapes
(unaligned leading dimensions, N = 1, N not a multiple of 16, K not a multiple of 32, split reductions); (2) the north-star
bar on that path: end-to-end render() against the UNMODIFIED reference's fixture within 1e-3."""
import numpy as np
import pytest
import torch

from oracle import nefes_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"
H, W, FOCAL, NEAR, FAR = 60, 80, 525.505 / 2 / 4, 0., 4.


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.fixture()
def tf32_gemm():
    from nefes_b200 import _lib as L
    prev = L.lib().nefes_gemm_mode(1)
    yield L
    L.lib().nefes_gemm_mode(prev)


def tf32_round(t):
    """Round-to-nearest to 10 explicit mantissa bits (what cvt.rna.tf32 does)."""
    i = t.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


@pytest.mark.parametrize("M,N,K,lda,act", [(1000, 128, 191, 191, 1), (777, 1, 128, 128, 2), (4096, 131, 64, 64, 0), (300, 64, 155, 160, 1),
                                           (129, 137, 63, 64, 0), (2048, 128, 1600, 1600, 0)])
def test_tf32_linear_fwd(tf32_gemm, M, N, K, lda, act):
    L = tf32_gemm
    g = torch.Generator(device=DEV).manual_seed(M + N + K)
    A = torch.randn(M, lda, device=DEV, generator=g)
    Wt = torch.randn(N, K, device=DEV, generator=g) / K ** 0.5
    b = torch.randn(N, device=DEV, generator=g)
    C = torch.full((M, N + 3), 7.0, device=DEV)                       # ldc > N: the padding columns must stay untouched
    L.check(L.lib().nefes_linear_fwd(L.ptr(A), lda, L.ptr(Wt), L.ptr(b), L.ptr(C), N + 3, M, N, K, act, L.stream_of(A)), "fwd")
    want = tf32_round(A[:, :K].contiguous()).double() @ tf32_round(Wt).double().t() + b.double()
    want = {0: want, 1: want.relu(), 2: torch.nn.functional.softplus(want)}[act]
    assert rel(C[:, :N], want) < 2e-5, rel(C[:, :N], want)             # same operand rounding: only accumulation order differs
    exact = A[:, :K].double() @ Wt.double().t() + b.double()
    exact = {0: exact, 1: exact.relu(), 2: torch.nn.functional.softplus(exact)}[act]
    assert rel(C[:, :N], exact) < 2e-3                                  # against unrounded fp32 operands: tf32's 2^-11 per operand
    assert float((C[:, N:] - 7.0).abs().max()) == 0.0


def test_tf32_linear_dgrad_and_wgrad(tf32_gemm):
    L = tf32_gemm
    g = torch.Generator(device=DEV).manual_seed(5)
    for M, N, K in ((3000, 64, 155), (70000, 131, 64), (513, 128, 191)):
        dC = torch.randn(M, N, device=DEV, generator=g)
        Wt = torch.randn(N, K, device=DEV, generator=g) / N ** 0.5
        A = torch.randn(M, K, device=DEV, generator=g)
        mask = torch.randn(M, K, device=DEV, generator=g)
        dA = torch.empty(M, K, device=DEV)
        L.check(L.lib().nefes_linear_dgrad(L.ptr(dC), N, L.ptr(Wt), L.ptr(dA), K, M, N, K, L.ptr(mask), K, L.stream_of(dC)), "dgrad")
        want = (tf32_round(dC).double() @ tf32_round(Wt).double()) * (mask > 0)
        assert rel(dA, want) < 2e-5, (M, N, K, rel(dA, want))
        dW = torch.zeros(N, K, device=DEV)
        L.check(L.lib().nefes_linear_wgrad(L.ptr(dC), N, L.ptr(A), K, L.ptr(dW), M, N, K, L.stream_of(dC)), "wgrad")
        want = tf32_round(dC).double().t() @ tf32_round(A).double()
        assert rel(dW, want) < 5e-5, (M, N, K, rel(dW, want))
        assert rel(dW, dC.double().t() @ A.double()) < 2e-3


class Args:
    nerfh_nff, use_fine_only, NeRFW, transient_at_test, netchunk = True, False, True, True, 1 << 21


@pytest.mark.parametrize("fused", [False, True])
def test_render_train_golden_on_the_tf32_path(golden, weights, fused):
    """The north-star bar (rgb / feature / depth within 1e-3 relative of the reference on its random-init weights) on a
    TENSOR-CORE path: the reference's fixture g5 (outputs, loss, weight gradients), as test_render_train_golden checks it
    for the fp32 SIMT path."""
    import nefes_b200 as nb
    g = golden("g5_render.npz")
    wc, wf = weights
    c = nb.NeRFH_NFF("coarse", W=128, precision="tf32")
    f = nb.NeRFH_NFF("fine", W=128, encode_appearance=True, encode_transient=True, precision="tf32")
    c.load_state_dict(wc, strict=False)
    f.load_state_dict(wf)
    c, f = c.to(DEV), f.to(DEV)
    q = nb.StandardQuery(Args.netchunk) if fused else (
        lambda inputs, viewdirs, ts, fn, typ, output_transient, test_time, store_rgb:
        nb.run_network_NeRFH_NFF(inputs, viewdirs, ts, fn, typ=typ, output_transient=output_transient, netchunk=Args.netchunk,
                                 test_time=test_time, store_rgb=store_rgb))
    kw = dict(network_query_fn=q, N_importance=64, N_samples=64, network_fn=c, network_fine=f, use_viewdirs=True, white_bkgd=False,
              args=Args(), ndc=False, lindisp=False, near=NEAR, far=FAR, perturb=1., raw_noise_std=0., test_time=False)
    rays = (g["rays_o"].to(DEV), g["rays_d"].to(DEV))
    rgb, disp, acc, ex = nb.render(H, W, FOCAL, chunk=32768, rays=rays, img_idx=torch.zeros(1, 10), t_rand=g["train/t_rand"].to(DEV),
                                   u=g["train/u"].to(DEV), return_aux=True, retraw=True, **kw)
    out = dict(rgb_map=rgb, disp_map=disp, acc_map=acc, **ex)
    assert torch.equal(out["aux_z_coarse"].cpu(), g["train/z_coarse"])
    assert float((out["aux_inds"].cpu() != g["train/inds"]).float().mean()) < 0.01
    worst = {}
    for k in ("rgb_map", "acc_map", "feat_map", "rgb0", "acc0", "feat0", "beta", "transient_sigmas", "disp_map", "disp0"):
        worst[k] = rel(out[k], g["train/" + k])
        assert worst[k] < 1e-3, (k, worst[k])                       # the north-star bar
    print("tf32 path, worst relative error per output vs the reference fixture:", {k: f"{v:.1e}" for k, v in worst.items()})
    loss = O.nerfw_loss(out, g["train/target"].to(DEV)) + 0.04 * (out["feat_map"].abs().mean() + out["feat0"].abs().mean())
    assert abs(float(loss) - float(g["train/loss"])) < 1e-3 * abs(float(g["train/loss"]))
    loss.backward()
    vf, vc = f.layer_views(f.flat.grad), c.layer_views(c.flat.grad)
    gw = {}
    for key, ref in g.items():
        if key.startswith("train/grad_fine/"):
            gw[key] = rel(vf[key[len("train/grad_fine/"):]], ref)
        if key.startswith("train/grad_coarse/"):
            gw[key] = rel(vc[key[len("train/grad_coarse/"):]], ref)
    print("tf32 path, weight gradients vs the reference fixture:", {k.split('/')[-1] + k.split('/')[1][-1]: f"{v:.1e}" for k, v in gw.items()})
    assert max(gw.values()) < 1e-2, gw                              # tf32 operands in dgrad / wgrad: the first layer (PE inputs up to 2^9 x) is the worst, 8e-3
