"""GPU: the fused ScoreNet / RK45 / PC / energy kernels through the C ABI vs the golden fixtures
(outputs of the reference itself, tests/golden/make_golden.py) and vs the oracle on fresh inputs.

Tolerances (BASELINE.json north_star, fp32 mode): final poses within 1e-3 rad geodesic and 1e-4
translation of the reference; step counts (nfev / accepted) must match exactly."""
import numpy as np
import pytest
import torch

from genpose2_b200 import synthetic
from oracle import pose_oracle as po
from tests.util import load_golden, pose_errors, rep

pytestmark = pytest.mark.gpu

ROT_TOL = 1e-3   # rad
TRANS_TOL = 1e-4


def make_net(seed, agent_type="score", mlp_mode="fp32"):
    from genpose2_b200.config import get_config
    from genpose2_b200.posenet import GFObjectPose
    from genpose2_b200.sde import init_sde
    cfg = get_config()
    cfg.agent_type = agent_type
    cfg.mlp_mode = mlp_mode
    net = GFObjectPose(cfg, *init_sde("ve")).cuda().eval()
    net.load_state_dict(synthetic.random_gfobjectpose_state_dict(seed))
    return net


@pytest.mark.parametrize("mlp_mode", ["fp32", "fp32_ffma"])
def test_scorenet_eval_matches_oracle(mlp_mode):
    net = make_net(100, mlp_mode=mlp_mode)
    trunk = po.Trunk(synthetic.random_gfobjectpose_state_dict(100))
    g = torch.Generator().manual_seed(0)
    B, R = 5, 50
    feat = torch.relu(torch.randn(B, 1024, generator=g))
    x = torch.randn(B * R, 9, generator=g)
    for tval in (1.0, 0.55, 0.123, 1e-5):
        t = torch.full((B * R, 1), tval)
        want = trunk.score(rep(feat, R), x, t)
        got = net({"pts_feat": rep(feat, R).cuda(), "sampled_pose": x.cuda(), "t": t.cuda()}, mode="score").cpu()
        scale = want.abs().max()
        print(mlp_mode, "eval rel err", tval, float((got - want).abs().max() / scale))
        assert (got - want).abs().max() <= 2e-5 * scale, (tval, float((got - want).abs().max()), float(scale))
        # hoisted per-object path gives the same numbers as one-object-per-row
        got2 = net({"_gp_pts_feat_obj": feat.cuda(), "_gp_rows_per_object": R, "pts_feat": None,
                    "sampled_pose": x.cuda(), "t": t.cuda()}, mode="score").cpu()
        assert torch.equal(got, got2)
    # rows with different t inside one tile
    t = torch.rand(B * R, 1, generator=g)
    want = trunk.score(rep(feat, R), x, t)
    got = net({"pts_feat": rep(feat, R).cuda(), "sampled_pose": x.cuda(), "t": t.cuda()}, mode="score").cpu()
    assert ((got - want).abs() <= 2e-5 * want.abs().max()).all()


@pytest.mark.parametrize("mlp_mode", ["fp32", "fp32_ffma"])
@pytest.mark.parametrize("name", ["ode_c1_T1", "ode_b4_T055", "ode_track_T025"])
def test_ode_sampler_matches_reference_golden(name, mlp_mode):
    from genpose2_b200 import samplers
    g = load_golden(name)
    net = make_net(int(g["score_seed"]), mlp_mode=mlp_mode)
    R, B = int(g["R"]), int(g["B"])
    feat, center = torch.from_numpy(g["feat"]).cuda(), torch.from_numpy(g["center"]).cuda()
    noise = torch.from_numpy(g["noise"])
    init = rep(torch.from_numpy(g["init_x"]), R).cuda() if "init_x" in g else None
    data = {"pts": torch.zeros(B * R, 1, 3), "pts_center": rep(center, R), "_gp_pts_feat_obj": feat,
            "_gp_rows_per_object": R}
    xs, x = samplers.cond_ode_sampler(net, data, lambda shape, T: noise.clone(), net.sde_fn, atol=1e-5, rtol=1e-5,
                                      device="cuda", eps=1e-5, T=float(g["T0"]), pose_mode="rot_matrix",
                                      init_x=init)
    st = samplers.ode_stats()
    assert x.dtype == torch.float64 and xs.dtype == torch.float64
    assert st["status"] == 0
    chaotic = name == "ode_c1_T1"
    if not chaotic:
        # evaluation settings: the controller takes exactly the reference's steps
        assert st["nfev"] + 1 == int(g["nfev"]), (st, int(g["nfev"]))
        assert xs.shape == (B * R, int(g["S"]), 9)
    else:
        # T0 = 1.0: the embedded error estimate is a cancellation of scores that are only accurate to ~1e-6, so the
        # first attempt's error norm already differs by 2e-4 relative between two float32-class evaluators (FFMA vs
        # split-bf16 tensor cores: 2.6451 vs 2.6456) and the step-size sequences decorrelate within ~25 steps.
        # 57 or 58 accepted steps both occur; the count is held to +-2 steps of the reference's.
        assert abs(st["nfev"] + 1 - int(g["nfev"])) <= 12, (st, int(g["nfev"]))
        assert abs(xs.shape[1] - int(g["S"])) <= 2 and xs.shape[0] == B * R
    # The bound is the north-star tolerance (1e-3 rad / 1e-4) wherever the problem is determined that well: at the
    # evaluation settings T0 = 0.55 / 0.25 the reference's own 1-thread and 8-thread CPU results (x_1thread vs x)
    # agree to < 1e-5.  At T0 = 1.0 (sigma_max = 50, random weights) they do not: the ODE amplifies
    # float32-rounding-sized changes of the score by 1e3..1e4, heavy-tailed (a hypothesis now and then lands in
    # another basin).  tests/golden/make_sensitivity.py measures that envelope -- 32 runs of THE REFERENCE's own
    # cond_ode_sampler with the score perturbed by 1e-6 relative, the size of a changed sgemm summation order: median
    # 8e-4 rad / 9e-4, max 7.8e-3 rad / 1.7e-2 for the WORST of the 50 hypotheses of a draw -- and the T0 = 1.0 case is
    # gated on the statistics those draws pin down (below).  T0 = 1.0 is not an evaluation setting.
    self_rot, self_trans = pose_errors(g["x_1thread"], g["x"])
    rot_tol, trans_tol = ROT_TOL, TRANS_TOL
    rot, trans = pose_errors(x.cpu().numpy(), g["x"])
    print(f"{name}: rot {rot:.3e} trans {trans:.3e} (reference self-distance {self_rot:.3e} / {self_trans:.3e})")
    if name == "ode_c1_T1":
        # The tail is heavy (now and then one of the 50 hypotheses lands in another basin), so the worst hypothesis of one
        # run is not bounded by the worst of 32 reference draws.  What the reference's own draws do pin down is the
        # TYPICAL hypothesis: its median over the 50 rows stays below 7.7e-5 rad / 1.1e-2 (|t| ~ sigma = 50 there) in all
        # 32 draws and the 90 % quantile below 2.1e-4 rad / 1.5e-2.  Gates: median and 90 % quantile of OUR rows within
        # twice the largest reference draw of the same statistic (rotation also within the north-star 1e-3 rad), and a
        # loose guard on the worst row.
        from tests.util import geodesic_6d
        sens = load_golden("ode_c1_T1_sens")
        rr = geodesic_6d(x.cpu().numpy()[:, :6], g["x"][:, :6])
        tr = np.linalg.norm(x.cpu().numpy()[:, 6:] - g["x"][:, 6:], axis=1)
        ref_med = (np.median(sens["rot_rows"], axis=1).max(), np.median(sens["trans_rows"], axis=1).max())
        ref_q90 = (np.quantile(sens["rot_rows"], 0.9, axis=1).max(), np.quantile(sens["trans_rows"], 0.9, axis=1).max())
        print(f"   per hypothesis: median {np.median(rr):.3e} / {np.median(tr):.3e} (reference draws <= {ref_med[0]:.3e} / "
              f"{ref_med[1]:.3e}); q90 {np.quantile(rr, 0.9):.3e} / {np.quantile(tr, 0.9):.3e} (<= {ref_q90[0]:.3e} / {ref_q90[1]:.3e})")
        assert self_rot <= sens["rot"].max() and self_trans <= sens["trans"].max()  # the reference's own draw fits
        assert np.median(rr) <= min(ROT_TOL, 2 * ref_med[0]) and np.median(tr) <= 2 * ref_med[1]
        assert np.quantile(rr, 0.9) <= min(ROT_TOL, 2 * ref_q90[0]) and np.quantile(tr, 0.9) <= 2 * ref_q90[1]
        assert rot <= 0.1 and trans <= 0.1, (rot, trans)     # guard only: see above
        rot_tol, trans_tol = 0.1, 0.1
    else:
        assert rot <= rot_tol and trans <= trans_tol, (rot, trans, self_rot, self_trans)
    for key, s in (("xs_last", -1), ("xs_mid", xs.shape[1] // 2), ("xs_first", 0)):
        if key == "xs_mid" and xs.shape[1] != int(g["S"]):
            continue  # another number of steps: the middle state is at another time
        rot, trans = pose_errors(xs[:, s].cpu().numpy(), g[key])
        # mid-trajectory states still carry sigma(t)-sized translations (up to ~50): bound relative to that
        mag = max(1.0, float(np.abs(g[key][:, 6:]).max()))
        assert rot <= rot_tol and trans <= trans_tol * mag, (key, rot, trans, mag)
    # the repeated-feature (reference-style) call gives identical results to the hoisted one
    data2 = {"pts": torch.zeros(B * R, 1, 3), "pts_center": rep(center, R), "pts_feat": rep(feat, R)}
    _, x2 = samplers.cond_ode_sampler(net, data2, lambda shape, T: noise.clone(), net.sde_fn, atol=1e-5, rtol=1e-5,
                                      device="cuda", eps=1e-5, T=float(g["T0"]), pose_mode="rot_matrix",
                                      init_x=init, return_trajectory=False)
    assert torch.equal(x, x2)


@pytest.mark.parametrize("mlp_mode", ["fp32", "fp32_ffma"])
def test_ode_sampler_vs_oracle_ragged_batch(mlp_mode):
    """a batch size that is not a multiple of the tile, rows_per_object = 7 (more objects per tile than the
    tensor-core evaluator's shared-memory (proj + tq) table holds: exercises its global fallback)"""
    from genpose2_b200 import samplers
    net = make_net(100, mlp_mode=mlp_mode)
    trunk = po.Trunk(synthetic.random_gfobjectpose_state_dict(100))
    g = torch.Generator().manual_seed(5)
    B, R = 13, 7
    feat = torch.relu(torch.randn(B, 1024, generator=g))
    center = torch.randn(B, 3, generator=g) * 0.1
    noise = torch.randn(B * R, 9, generator=g) * po.ve_marginal_std(0.4)
    _, want, stats = po.cond_ode_sampler(trunk, rep(feat, R), rep(center, R), noise, T=0.4)
    data = {"pts": torch.zeros(B * R, 1, 3), "pts_center": rep(center, R).cuda(), "_gp_pts_feat_obj": feat.cuda(),
            "_gp_rows_per_object": R}
    xs, x = samplers.cond_ode_sampler(net, data, lambda shape, T: noise.clone(), net.sde_fn, device="cuda",
                                      T=0.4, pose_mode="rot_matrix")
    st = samplers.ode_stats()
    assert st["nfev"] == stats["nfev"] and st["accepted"] == stats["n_accepted"] and st["rejected"] == stats["n_rejected"]
    rot, trans = pose_errors(x.cpu().numpy(), want.numpy())
    assert rot <= ROT_TOL and trans <= TRANS_TOL, (rot, trans)


@pytest.mark.parametrize("mlp_mode", ["fp32", "fp32_ffma"])
def test_pc_sampler_matches_reference_golden(mlp_mode):
    from genpose2_b200 import samplers
    g = load_golden("pc_b2")
    net = make_net(int(g["score_seed"]), mlp_mode=mlp_mode)
    R, B, steps = int(g["R"]), int(g["B"]), int(g["steps"])
    feat, center = torch.from_numpy(g["feat"]).cuda(), torch.from_numpy(g["center"]).cuda()
    data = {"pts": torch.zeros(B * R, 1, 3), "pts_center": rep(center, R), "_gp_pts_feat_obj": feat,
            "_gp_rows_per_object": R}
    xs, mean_x = samplers.cond_pc_sampler(net, data, None, net.sde_fn, num_steps=steps, snr=0.16, device="cuda",
                                          eps=1e-5, pose_mode="rot_matrix", init_x=torch.from_numpy(g["init"]).cuda(),
                                          noise=torch.from_numpy(g["noises"]).cuda())
    assert xs.shape == (B * R, steps, 9) and xs.dtype == torch.float32
    # early steps live at sigma ~ 50: compare relative to the state magnitude there, tightly at the end
    ref_xs = g["xs"]
    err = np.abs(xs.cpu().numpy() - ref_xs).max(axis=(0, 2))
    mag = np.abs(ref_xs).max(axis=(0, 2))
    assert (err <= 2e-3 * np.maximum(mag, 1.0)).all(), err / np.maximum(mag, 1.0)
    rot, trans = pose_errors(mean_x.cpu().numpy(), g["mean_x"])
    # Random weights make this 25-step run diverge to |t| ~ 440 (float32 ulp there: 3e-5), so an absolute 1e-4 on the
    # translation is not defined by float32 arithmetic.  tests/golden/make_sensitivity.py measures what IS defined, on the
    # reference's own sampler: with its score perturbed by 1e-6 relative (a changed float32 summation order) the result
    # moves by up to 8.9e-4 (median 3.4e-4); perturbed by 2^-17 (the operand rounding of the split-bf16 tensor-core mode,
    # 16 mantissa bits per operand) by 1.9e-3 .. 2.6e-3.  Bounds: north-star 1e-3 rad on the rotation; on the translation
    # twice the largest reference draw of the mode's class: fp32_ffma (true float32) -> 1e-6 class, fp32 (split bf16) ->
    # 2^-17 class (= 1.2e-5 of the state magnitude).
    sens = load_golden("pc_b2_sens")
    env = float((sens["trans"] if mlp_mode == "fp32_ffma" else sens["trans_operand"]).max())
    print(f"pc 25 steps [{mlp_mode}]: rot {rot:.3e} trans {trans:.3e} (reference envelope of this mode's class, max: {env:.3e}; "
          f"|t| max {float(sens['state_magnitude']):.0f}); xs rel err max {float((err / np.maximum(mag, 1.0)).max()):.3e}")
    assert rot <= ROT_TOL and trans <= 2 * env, (rot, trans)


@pytest.mark.parametrize("mlp_mode", ["fp32", "fp32_ffma"])
def test_pc_sampler_500_steps_matches_reference_golden(mlp_mode):
    """cond_pc_sampler at the reference's default num_steps = 500 (samplers.py:118), float32 state like the reference.
    The prior and the 2 x 500 noise tensors are redrawn from the fixture's seed in the reference's call order.
    Rotation is held to the north-star 1e-3 rad.  Translation (the state runs away to |t| ~ 390 under random weights):
    the reference does not reproduce ITSELF to 1e-4 over 500 float32 steps (1 vs 8 CPU threads: 1.1e-4, fixture field
    mean_x_1thread); tests/golden/make_sensitivity.py measures the envelope on the reference's own sampler (score
    perturbed by 1e-6 relative: median 1.9e-4, max 2.5e-4; by 2^-17, the split-bf16 operand rounding: 4e-4 .. 1.1e-3)
    and the bound is twice the largest draw of the mode's class; the observed value is printed and asserted."""
    from genpose2_b200 import samplers
    g = load_golden("pc_b2_500")
    sens = load_golden("pc_b2_500_sens")
    net = make_net(int(g["score_seed"]), mlp_mode=mlp_mode)
    R, B, steps = int(g["R"]), int(g["B"]), int(g["steps"])
    feat, center = torch.from_numpy(g["feat"]).cuda(), torch.from_numpy(g["center"]).cuda()
    torch.manual_seed(int(g["noise_seed"]))
    init = net.prior_fn((B * R, 9))
    noises = torch.stack([torch.stack([torch.randn(B * R, 9), torch.randn(B * R, 9)]) for _ in range(steps)])
    data = {"pts": torch.zeros(B * R, 1, 3), "pts_center": rep(center, R), "_gp_pts_feat_obj": feat,
            "_gp_rows_per_object": R}
    xs, mean_x = samplers.cond_pc_sampler(net, data, None, net.sde_fn, num_steps=steps, snr=0.16, device="cuda",
                                          eps=1e-5, pose_mode="rot_matrix", init_x=init.cuda(), noise=noises.cuda())
    assert xs.shape == (B * R, steps, 9) and xs.dtype == torch.float32
    keep = [int(k) for k in g["keep"]]
    got = xs[:, keep].cpu().numpy()
    err = np.abs(got - g["xs_keep"]).max(axis=(0, 2))
    mag = np.maximum(np.abs(g["xs_keep"]).max(axis=(0, 2)), 1.0)
    rot, trans = pose_errors(mean_x.cpu().numpy(), g["mean_x"])
    self_rot, self_trans = pose_errors(g["mean_x_1thread"], g["mean_x"])
    print(f"pc 500 steps [{mlp_mode}]: rot {rot:.3e} trans {trans:.3e} (reference self-distance {self_rot:.3e} / "
          f"{self_trans:.3e}, envelope max {float(sens['rot'].max()):.3e} / {float(sens['trans'].max()):.3e}); "
          f"xs rel err at steps {keep}: {np.round(err / mag, 6).tolist()}")
    assert self_trans <= 2 * float(sens["trans"].max())   # the reference's own draw fits its envelope
    env = float((sens["trans"] if mlp_mode == "fp32_ffma" else sens["trans_operand"]).max())
    assert rot <= ROT_TOL and trans <= 2 * env, (rot, trans, env)
    assert (err <= 2e-3 * mag).all(), (err / mag)


def test_energy_matches_reference_golden():
    from genpose2_b200.config import get_config
    from genpose2_b200.posenet_agent import PoseNet
    g = load_golden("energy_b3")
    cfg = get_config()
    cfg.agent_type = "energy"
    agent = PoseNet(cfg)
    agent.net.load_state_dict(synthetic.random_gfobjectpose_state_dict(int(g["energy_seed"])))
    poses = torch.from_numpy(g["poses"]).cuda()
    data = {"pts_feat": torch.from_numpy(g["feat"]).cuda(), "pts_center": torch.from_numpy(g["center"]).cuda()}
    e = agent.get_energy(data, poses, T=1e-5, mode="test", extract_feature=False)
    assert e.shape == (3, 50, 2) and e.dtype == torch.float32
    want = g["energy"]
    assert np.abs(e.cpu().numpy() - want).max() <= 2e-5 * np.abs(want).max()


# ---------------------------------------------------------------------------------------------------
# bf16 tensor-core mode (tcgen05): same kernels, MLP operands rounded to bf16, fp32 accumulation.
# Stated bound (north_star: "a stated looser bound in bf16 mode"): the raw score within 2e-2 of its
# scale; final poses within BF16_ROT_TOL / BF16_TRANS_TOL of the reference at the evaluation settings.
# ---------------------------------------------------------------------------------------------------
BF16_ROT_TOL = 2e-2    # rad
BF16_TRANS_TOL = 2e-3


def make_net_mode(seed, mode):
    return make_net(seed, mlp_mode=mode)


def test_bf16_scorenet_eval_close_to_fp32():
    net32, net16 = make_net_mode(100, "fp32_ffma"), make_net_mode(100, "bf16")
    g = torch.Generator().manual_seed(3)
    B, R = 7, 50   # 350 rows: 2 full tiles + a ragged one
    feat = torch.relu(torch.randn(B, 1024, generator=g)).cuda()
    x = torch.randn(B * R, 9, generator=g).cuda()
    for tval in (0.55, 0.05):
        t = torch.full((B * R, 1), tval).cuda()
        d = {"_gp_pts_feat_obj": feat, "_gp_rows_per_object": R, "pts_feat": None, "sampled_pose": x, "t": t}
        a = net32(dict(d), mode="score")
        b = net16(dict(d), mode="score")
        err = float((a - b).abs().max() / a.abs().max())
        print("bf16 eval rel err", tval, err)
        assert err < 2e-2, err
        b2 = net16(dict(d), mode="score")
        assert torch.equal(b, b2)  # deterministic, and the pipeline state survives a second launch


@pytest.mark.parametrize("name", ["ode_b4_T055", "ode_track_T025"])
def test_bf16_ode_sampler_within_stated_bound(name):
    from genpose2_b200 import samplers
    g = load_golden(name)
    net = make_net_mode(int(g["score_seed"]), "bf16")
    R, B = int(g["R"]), int(g["B"])
    feat, center = torch.from_numpy(g["feat"]).cuda(), torch.from_numpy(g["center"]).cuda()
    noise = torch.from_numpy(g["noise"])
    init = rep(torch.from_numpy(g["init_x"]), R).cuda() if "init_x" in g else None
    data = {"pts": torch.zeros(B * R, 1, 3), "pts_center": rep(center, R), "_gp_pts_feat_obj": feat,
            "_gp_rows_per_object": R}
    xs, x = samplers.cond_ode_sampler(net, data, lambda shape, T: noise.clone(), net.sde_fn, device="cuda",
                                      T=float(g["T0"]), pose_mode="rot_matrix", init_x=init)
    st = samplers.ode_stats()
    rot, trans = pose_errors(x.cpu().numpy(), g["x"])
    print(f"bf16 {name}: rot {rot:.3e} rad trans {trans:.3e}; nfev {st['nfev'] + 1} (reference {int(g['nfev'])})")
    assert st["status"] == 0
    assert rot <= BF16_ROT_TOL and trans <= BF16_TRANS_TOL, (rot, trans)


@pytest.mark.parametrize("mlp_mode", ["fp32", "fp32_ffma"])
def test_ode_sampler_dense_output_matches_reference_golden(mlp_mode):
    """num_steps = 20: scipy's t_eval / RkDenseOutput path of cond_ode_sampler (samplers.py:222-249), scope row f2"""
    from genpose2_b200 import samplers
    g = load_golden("ode_b2_T055_steps20")
    net = make_net(int(g["score_seed"]), mlp_mode=mlp_mode)
    R, B, S = int(g["R"]), int(g["B"]), int(g["num_steps"])
    feat, center = torch.from_numpy(g["feat"]).cuda(), torch.from_numpy(g["center"]).cuda()
    noise = torch.from_numpy(g["noise"])
    data = {"pts": torch.zeros(B * R, 1, 3), "pts_center": rep(center, R), "_gp_pts_feat_obj": feat,
            "_gp_rows_per_object": R}
    xs, x = samplers.cond_ode_sampler(net, data, lambda shape, T: noise.clone(), net.sde_fn, atol=1e-5, rtol=1e-5,
                                      device="cuda", eps=1e-5, T=float(g["T0"]), num_steps=S, pose_mode="rot_matrix")
    st = samplers.ode_stats()
    assert st["status"] == 0 and st["nfev"] + 1 == int(g["nfev"])
    assert xs.shape == (B * R, S, 9) and xs.dtype == torch.float64
    for key, arr in (("x", x), ("xs_last", xs[:, -1]), ("xs_mid", xs[:, S // 2]), ("xs_first", xs[:, 0])):
        rot, trans = pose_errors(arr.cpu().numpy(), g[key])
        mag = max(1.0, float(np.abs(g[key][:, 6:]).max()))
        print(f"dense {mlp_mode} {key}: rot {rot:.3e} trans {trans:.3e}")
        assert rot <= ROT_TOL and trans <= TRANS_TOL * mag, (key, rot, trans)
    # the interpolated grid against the oracle at every point
    trunk = po.Trunk(synthetic.random_gfobjectpose_state_dict(int(g["score_seed"])))
    xs_o, x_o, _ = po.cond_ode_sampler(trunk, rep(torch.from_numpy(g["feat"]), R), rep(torch.from_numpy(g["center"]), R), noise,
                                       T=float(g["T0"]), num_steps=S)
    for s_i in range(S):
        rot, trans = pose_errors(xs[:, s_i].cpu().numpy(), xs_o[:, s_i].numpy())
        mag = max(1.0, float(xs_o[:, s_i, 6:].abs().max()))
        assert rot <= ROT_TOL and trans <= TRANS_TOL * mag, (s_i, rot, trans)


def test_failed_integration_is_loud():
    """An integration the controller cannot finish (a NaN in the initial state -> NaN error norm -> every attempt is
    rejected -> step size underflow, scipy status -1) poisons its poses with NaN in band and raises at the next look at
    the statistics: the reference ignores solve_ivp's status (samplers.py:226-236), the accelerated path does not pass
    a failure silently."""
    from genpose2_b200 import samplers
    net = make_net(100)
    g = torch.Generator().manual_seed(9)
    B, R = 2, 8
    feat = torch.relu(torch.randn(B, 1024, generator=g)).cuda()
    center = torch.zeros(B, 3).cuda()
    noise = torch.randn(B * R, 9, generator=g) * po.ve_marginal_std(0.5)
    bad = noise.clone()
    bad[3, 4] = float("nan")
    data = {"pts": torch.zeros(B * R, 1, 3), "pts_center": rep(center, R), "_gp_pts_feat_obj": feat,
            "_gp_rows_per_object": R}
    samplers._pending_stats.clear()
    _, x = samplers.cond_ode_sampler(net, data, lambda shape, T: bad.clone(), net.sde_fn, device="cuda", T=0.5,
                                     pose_mode="rot_matrix", return_trajectory=False)
    assert torch.isnan(x).all()          # every row, not just the one that started as NaN
    with pytest.raises(RuntimeError, match="status -1"):
        samplers.ode_stats()
    # the asynchronous watch raises at the next look once the statistics have reached the host
    _, x = samplers.cond_ode_sampler(net, data, lambda shape, T: bad.clone(), net.sde_fn, device="cuda", T=0.5,
                                     pose_mode="rot_matrix", return_trajectory=False)
    torch.cuda.synchronize()
    with pytest.raises(RuntimeError, match="status -1"):
        samplers.check_pending()
    samplers._pending_stats.clear()
    # and a healthy call afterwards is unaffected
    _, x = samplers.cond_ode_sampler(net, data, lambda shape, T: noise.clone(), net.sde_fn, device="cuda", T=0.5,
                                     pose_mode="rot_matrix", return_trajectory=False)
    assert torch.isfinite(x).all() and samplers.ode_stats()["status"] == 0
