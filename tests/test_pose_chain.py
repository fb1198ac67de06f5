"""CPU: the pose parameterisation chain of the refinement -- host-side torch mirrors."""
import numpy as np
import torch

from nefes_b200 import refine


def twist_exp(t, r):
    """Independent statement of SE3.exp([t, r]): the matrix exponential of the 4x4 twist [[K(r), t], [0, 0]] in fp64."""
    X = torch.zeros(4, 4, dtype=torch.float64)
    X[:3, :3] = refine.vec2skew(r.double())
    X[:3, 3] = t.double()
    return torch.matrix_exp(X)


def test_se3_exp_matches_matrix_exponential():
    """poses.py:31-32, 44: SE3.exp([t, r]).matrix().  lietorch is not vendored by the reference (parity unpinned for the
    library itself); the closed form here is checked against the definition, including the small-angle series."""
    g = torch.Generator().manual_seed(0)
    for scale in (1.0, 0.3, 1e-3, 1e-6, 0.0):
        r = torch.randn(3, generator=g, dtype=torch.float64) * scale
        t = torch.randn(3, generator=g, dtype=torch.float64)
        R, vt = refine.se3_exp(t, r)
        E = twist_exp(t, r)
        assert float((R - E[:3, :3]).abs().max()) < 1e-12, scale
        assert float((vt - E[:3, 3]).abs().max()) < 1e-12, scale


def test_se3_exp_gradient_is_finite_and_right_at_zero():
    r = torch.zeros(3, dtype=torch.float64, requires_grad=True)
    t = torch.tensor([0.3, -0.2, 0.5], dtype=torch.float64, requires_grad=True)
    R, vt = refine.se3_exp(t, r)
    (vt * torch.tensor([1.0, 2.0, 3.0], dtype=torch.float64)).sum().backward()
    # d(V t)/dr at r = 0 is 0.5 * d(K t)/dr = -0.5 * skew(t)
    want = (-0.5 * refine.vec2skew(t.detach())).t() @ torch.tensor([1.0, 2.0, 3.0], dtype=torch.float64)
    assert torch.isfinite(r.grad).all() and float((r.grad - want).abs().max()) < 1e-12
    assert float((t.grad - torch.tensor([1.0, 2.0, 3.0], dtype=torch.float64)).abs().max()) < 1e-12


def test_learn_pose_lietorch_and_plain():
    init = torch.eye(4)[None].repeat(2, 1, 1)
    init[1, :3, 3] = torch.tensor([1.0, 2.0, 3.0])
    for lie in (False, True):
        p = refine.LearnPose(2, True, True, init, lietorch=lie)
        with torch.no_grad():
            p.r[1] = torch.tensor([0.2, -0.1, 0.3])
            p.t[1] = torch.tensor([0.5, 0.0, -0.5])
        c2w = p(1)
        assert c2w.shape == (4, 4) and torch.equal(c2w[3], torch.tensor([0., 0., 0., 1.]))
        E = twist_exp(p.t[1].detach(), p.r[1].detach())
        assert float((c2w[:3, :3].double() - E[:3, :3]).abs().max()) < 1e-6
        want_t = (E[:3, 3] if lie else p.t[1].detach().double()) + init[1, :3, 3].double()
        assert float((c2w[:3, 3].double() - want_t).abs().max()) < 1e-6
        assert torch.equal(p(0), torch.eye(4))                      # zero delta: the initial pose, exactly


def test_fix_coord_supp_and_svd_reg():
    """dm/direct_pose_model.py:210-232 and dm/DFM_pose_refine.py:119-129."""
    g = torch.Generator().manual_seed(1)
    pose = torch.randn(5, 3, 4, generator=g)
    ws = {"pose_scale": 0.5, "move_all_cam_vec": [1.0, -2.0, 0.5], "pose_scale2": 3.0}
    ref = pose.clone()
    ref[:, :3, 3] *= ws["pose_scale"]
    ref[:, :3, 3] += torch.tensor(ws["move_all_cam_vec"])
    ref[:, :3, 3] *= ws["pose_scale2"]
    out = refine.fix_coord_supp(None, pose, ws)
    assert torch.equal(out, ref) and not torch.equal(pose, ref)        # same numbers, argument untouched
    ident = refine.fix_coord_supp(None, pose, {"pose_scale": 1, "move_all_cam_vec": [0., 0., 0.], "pose_scale2": 1})
    assert torch.equal(ident, pose)                                    # the 7-Scenes stairs world setup
    u, s, v = torch.svd(pose[:, :3, :3])
    reg = refine.svd_reg(pose)
    assert float((reg[:, :3, :3] - u @ v.transpose(-2, -1)).abs().max()) == 0.0
    eye = reg[:, :3, :3] @ reg[:, :3, :3].transpose(-2, -1)
    assert float((eye - torch.eye(3)).abs().max()) < 1e-5 and torch.equal(reg[:, :, 3], pose[:, :, 3])
