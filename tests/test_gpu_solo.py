"""GPU: the one-CTA-per-tile ("solo") shape of the tensor-core ScoreNet evaluator -- with the A operand in tensor
memory (csrc/trunk_solo_t.cuh, the default) and in shared memory (csrc/trunk_solo.cuh, GP_MODE_SMEM_A) -- forced
through the GP_MODE_SOLO flag of the C ABI at sizes where the default is the cluster shape, against the same
reference goldens / oracle and with the same gates as the cluster shape; and at a batch where CTAs own several
tiles (more tiles than SMs) against the cluster shape."""
import numpy as np
import pytest
import torch

from genpose2_b200 import synthetic
from oracle import pose_oracle as po
from tests.test_gpu_sampler import BF16_ROT_TOL, BF16_TRANS_TOL, ROT_TOL, TRANS_TOL, make_net
from tests.util import load_golden, pose_errors, rep

pytestmark = pytest.mark.gpu


def solo_net(seed, mlp_mode, agent_type="score", shape="solo"):
    net = make_net(seed, agent_type=agent_type, mlp_mode=mlp_mode)
    net.pose_score_net.eval_shape = shape
    return net


SHAPES = ["solo", "solo_smem_a"]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("mlp_mode,tol", [("fp32", 2e-5), ("bf16", 2e-2)])
def test_solo_scorenet_eval_matches_oracle(mlp_mode, tol, shape):
    net = solo_net(100, mlp_mode, shape=shape)
    trunk = po.Trunk(synthetic.random_gfobjectpose_state_dict(100))
    g = torch.Generator().manual_seed(0)
    B, R = 7, 50   # 350 rows: three tiles, the last one partial; tiles span up to four objects
    feat = torch.relu(torch.randn(B, 1024, generator=g))
    x = torch.randn(B * R, 9, generator=g)
    for tval in (1.0, 0.55, 1e-5):
        t = torch.full((B * R, 1), tval)
        want = trunk.score(rep(feat, R), x, t)
        got = net({"_gp_pts_feat_obj": feat.cuda(), "_gp_rows_per_object": R, "pts_feat": None,
                   "sampled_pose": x.cuda(), "t": t.cuda()}, mode="score").cpu()
        err = float((got - want).abs().max() / want.abs().max())
        print(f"solo eval [{mlp_mode}] t={tval}: rel err {err:.2e}")
        assert err <= tol, (tval, err)
    # one object per row (the proj table does not fit the shared-memory slots: global fallback), per-row t
    t = torch.rand(B * R, 1, generator=g)
    want = trunk.score(rep(feat, R), x, t)
    got = net({"pts_feat": rep(feat, R).cuda(), "sampled_pose": x.cuda(), "t": t.cuda()}, mode="score").cpu()
    assert float((got - want).abs().max() / want.abs().max()) <= tol


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("mlp_mode", ["fp32", "bf16"])
@pytest.mark.parametrize("name", ["ode_b4_T055", "ode_track_T025"])
def test_solo_ode_sampler_matches_reference_golden(name, mlp_mode, shape):
    from genpose2_b200 import samplers
    g = load_golden(name)
    net = solo_net(int(g["score_seed"]), mlp_mode, shape=shape)
    R, B = int(g["R"]), int(g["B"])
    feat, center = torch.from_numpy(g["feat"]).cuda(), torch.from_numpy(g["center"]).cuda()
    noise = torch.from_numpy(g["noise"])
    init = rep(torch.from_numpy(g["init_x"]), R).cuda() if "init_x" in g else None
    data = {"pts": torch.zeros(B * R, 1, 3), "pts_center": rep(center, R), "_gp_pts_feat_obj": feat,
            "_gp_rows_per_object": R}
    xs, x = samplers.cond_ode_sampler(net, data, lambda shape, T: noise.clone(), net.sde_fn, atol=1e-5, rtol=1e-5,
                                      device="cuda", eps=1e-5, T=float(g["T0"]), pose_mode="rot_matrix", init_x=init)
    st = samplers.ode_stats()
    rot, trans = pose_errors(x.cpu().numpy(), g["x"])
    print(f"{shape} {name} [{mlp_mode}]: rot {rot:.3e} trans {trans:.3e} nfev {st['nfev'] + 1} (reference {int(g['nfev'])})")
    assert st["status"] == 0
    if mlp_mode == "fp32":
        assert st["nfev"] + 1 == int(g["nfev"]) and xs.shape == (B * R, int(g["S"]), 9)
        assert rot <= ROT_TOL and trans <= TRANS_TOL, (rot, trans)
    else:
        assert rot <= BF16_ROT_TOL and trans <= BF16_TRANS_TOL, (rot, trans)


def test_solo_dense_output_and_energy_match_reference_goldens():
    from genpose2_b200 import samplers
    g = load_golden("ode_b2_T055_steps20")
    net = solo_net(int(g["score_seed"]), "fp32")
    R, B, S = int(g["R"]), int(g["B"]), int(g["num_steps"])
    feat, center = torch.from_numpy(g["feat"]).cuda(), torch.from_numpy(g["center"]).cuda()
    noise = torch.from_numpy(g["noise"])
    data = {"pts": torch.zeros(B * R, 1, 3), "pts_center": rep(center, R), "_gp_pts_feat_obj": feat,
            "_gp_rows_per_object": R}
    xs, x = samplers.cond_ode_sampler(net, data, lambda shape, T: noise.clone(), net.sde_fn, atol=1e-5, rtol=1e-5,
                                      device="cuda", eps=1e-5, T=float(g["T0"]), num_steps=S, pose_mode="rot_matrix")
    st = samplers.ode_stats()
    assert st["status"] == 0 and st["nfev"] + 1 == int(g["nfev"])
    for key, arr in (("x", x), ("xs_last", xs[:, -1]), ("xs_mid", xs[:, S // 2]), ("xs_first", xs[:, 0])):
        rot, trans = pose_errors(arr.cpu().numpy(), g[key])
        mag = max(1.0, float(np.abs(g[key][:, 6:]).max()))
        assert rot <= ROT_TOL and trans <= TRANS_TOL * mag, (key, rot, trans)
    # energy (gp_energy) through the solo shape
    from genpose2_b200.config import get_config
    from genpose2_b200.posenet_agent import PoseNet
    g = load_golden("energy_b3")
    cfg = get_config()
    cfg.agent_type = "energy"
    agent = PoseNet(cfg)
    agent.net.load_state_dict(synthetic.random_gfobjectpose_state_dict(int(g["energy_seed"])))
    agent.net.pose_score_net.eval_shape = "solo"
    data = {"pts_feat": torch.from_numpy(g["feat"]).cuda(), "pts_center": torch.from_numpy(g["center"]).cuda()}
    e = agent.get_energy(data, torch.from_numpy(g["poses"]).cuda(), T=1e-5, mode="test", extract_feature=False)
    err = float(np.abs(e.cpu().numpy() - g["energy"]).max() / np.abs(g["energy"]).max())
    assert err <= 2e-5, err


def test_solo_pc_sampler_matches_reference_golden():
    from genpose2_b200 import samplers
    g = load_golden("pc_b2")
    net = solo_net(int(g["score_seed"]), "fp32")
    R, B, steps = int(g["R"]), int(g["B"]), int(g["steps"])
    feat, center = torch.from_numpy(g["feat"]).cuda(), torch.from_numpy(g["center"]).cuda()
    data = {"pts": torch.zeros(B * R, 1, 3), "pts_center": rep(center, R), "_gp_pts_feat_obj": feat,
            "_gp_rows_per_object": R}
    xs, mean_x = samplers.cond_pc_sampler(net, data, None, net.sde_fn, num_steps=steps, snr=0.16, device="cuda",
                                          eps=1e-5, pose_mode="rot_matrix", init_x=torch.from_numpy(g["init"]).cuda(),
                                          noise=torch.from_numpy(g["noises"]).cuda())
    rot, trans = pose_errors(mean_x.cpu().numpy(), g["mean_x"])
    sens = load_golden("pc_b2_sens")   # reference-derived envelope, see test_gpu_sampler.test_pc_sampler_matches_reference_golden
    print(f"solo pc 25 steps: rot {rot:.3e} trans {trans:.3e}")
    assert rot <= ROT_TOL and trans <= 2 * float(sens["trans_operand"].max()), (rot, trans)


@pytest.mark.parametrize("mlp_mode", ["fp32", "bf16"])
def test_solo_many_tiles_per_cta_vs_cluster_shape(mlp_mode):
    """400 objects x 50 hypotheses = 157 tiles on 148 SMs: some CTAs integrate two tiles.  The two shapes sum the head
    columns in different orders, so they agree like two fp32-class evaluators (north-star gate, same step counts in
    fp32 mode), not bit for bit; each is deterministic."""
    from genpose2_b200 import samplers
    B, R, T0 = 400, 50, 0.55
    g = torch.Generator().manual_seed(21)
    feat = torch.relu(torch.randn(B, 1024, generator=g)).cuda()
    center = (torch.randn(B, 3, generator=g) * 0.1 + torch.tensor([0.0, 0.0, 0.8])).cuda()
    noise = torch.randn(B * R, 9, generator=g) * po.ve_marginal_std(T0)
    data = {"pts": torch.zeros(B * R, 1, 3), "pts_center": rep(center, R), "_gp_pts_feat_obj": feat,
            "_gp_rows_per_object": R}
    res = {}
    for shape in ("solo", "cluster", "auto", "solo_smem_a"):
        net = solo_net(100, mlp_mode, shape=shape)
        _, x = samplers.cond_ode_sampler(net, dict(data), lambda s, T: noise.clone(), net.sde_fn, device="cuda", T=T0,
                                         pose_mode="rot_matrix", return_trajectory=False)
        res[shape] = (x.cpu().numpy(), samplers.ode_stats())
    assert res["solo"][1]["status"] == 0 and res["cluster"][1]["status"] == 0
    assert np.array_equal(res["solo"][0], res["auto"][0])       # 157 tiles > 32: auto = solo; and deterministic
    from tests.util import geodesic_6d
    a, b = res["solo"][0], res["cluster"][0]
    rot = geodesic_6d(a[:, :6], b[:, :6])
    trans = np.linalg.norm(a[:, 6:] - b[:, 6:], axis=1)
    print(f"solo vs cluster [{mlp_mode}] 157 tiles: rot max {rot.max():.3e} median {np.median(rot):.3e}; trans max "
          f"{trans.max():.3e}; nfev {res['solo'][1]['nfev']} / {res['cluster'][1]['nfev']}")
    # the two solo variants differ only in the summation order of the 256 -> 3 output layer
    c = res["solo_smem_a"][0]
    rot2 = geodesic_6d(a[:, :6], c[:, :6])
    trans2 = np.linalg.norm(a[:, 6:] - c[:, 6:], axis=1)
    print(f"solo (TMEM A) vs solo (smem A): rot max {rot2.max():.3e} trans max {trans2.max():.3e}")
    assert rot2.max() <= (ROT_TOL if mlp_mode == "fp32" else BF16_ROT_TOL) and trans2.max() <= (TRANS_TOL if mlp_mode == "fp32" else BF16_TRANS_TOL)
    if mlp_mode == "fp32":
        assert res["solo"][1]["nfev"] == res["cluster"][1]["nfev"]
        assert rot.max() <= ROT_TOL and trans.max() <= TRANS_TOL, (rot.max(), trans.max())
    else:
        assert rot.max() <= BF16_ROT_TOL and trans.max() <= BF16_TRANS_TOL, (rot.max(), trans.max())
