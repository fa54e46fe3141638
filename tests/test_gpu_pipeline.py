"""GPU: the whole path through the reference-facing agents (PoseNet.pred_func -> get_energy ->
aggregate_pose -> pred_scale_func) vs the golden fixtures produced by the reference agents, the
encoder vs the CPU restatement, and the BASELINE config-2 size through size-independent properties."""
import numpy as np
import pytest
import torch

from genpose2_b200 import synthetic
from oracle import pose_oracle as po
from tests.util import geodesic_mats, load_golden, pose_errors

pytestmark = pytest.mark.gpu


def make_pipeline(seeds=(100, 200, 300)):
    from genpose2_b200.pipeline import PosePipeline
    return PosePipeline(device="cuda").load_synthetic_weights(seeds)


def inject_features(pipe, sfeat, efeat):
    """same patch the golden generator applied to the reference: extract_pts_feature returns given features"""
    def patch(agent, feat):
        def fake(data, geometry=None, return_geometry=False):
            return (feat, None) if return_geometry else feat
        agent.net.extract_pts_feature = fake
    patch(pipe.score_agent, sfeat)
    patch(pipe.energy_agent, efeat)


@pytest.mark.parametrize("name", ["full_b3_T055", "full_track_b2_T025"])
def test_full_path_matches_reference_golden(name):
    g = load_golden(name)
    pipe = make_pipeline((int(g["score_seed"]), int(g["energy_seed"]), int(g["scale_seed"])))
    inject_features(pipe, torch.from_numpy(g["score_feat"]).cuda(), torch.from_numpy(g["energy_feat"]).cuda())
    noise = torch.from_numpy(g["noise"])
    pipe.score_agent.net.prior_fn = lambda shape, T=1.0: noise.clone()
    data = {"pts": torch.from_numpy(g["pts"]).cuda(), "pts_center": torch.from_numpy(g["center"]).cuda()}
    init = torch.from_numpy(g["init_x"]).cuda() if "init_x" in g else None
    out = pipe(data, repeat_num=int(g["R"]), T0=float(g["T0"]), init_x=init, return_all=True)
    rot, trans = pose_errors(out["pred_pose"].cpu().numpy(), g["pred_pose"])
    assert rot <= 1e-3 and trans <= 1e-4, (rot, trans)
    assert out["pred_pose"].dtype == torch.float64 and out["pred_pose"].shape == g["pred_pose"].shape
    q_err = np.abs(np.abs(out["pred_pose_q_wxyz"].cpu().numpy()) - np.abs(g["pred_q"])).max()
    assert q_err <= 1e-3
    e, we = out["energy"].cpu().numpy(), g["energy"]
    assert np.abs(e - we).max() <= 2e-3 * np.abs(we).max()
    agg, wagg = out["aggregated_pose"].cpu().numpy().astype(np.float64), g["aggregated_pose"].astype(np.float64)
    assert geodesic_mats(agg[:, :3, :3], wagg[:, :3, :3]).max() <= 1e-3
    assert np.abs(agg[:, :3, 3] - wagg[:, :3, 3]).max() <= 1e-4
    np.testing.assert_allclose(out["length"].cpu().numpy(), g["length"], rtol=0, atol=1e-4)


def test_encoder_matches_cpu_restatement():
    from genpose2_b200.pointnet2 import Pointnet2ClsMSG
    sd = synthetic.random_encoder_state_dict(7, prefix="")
    enc = Pointnet2ClsMSG(0).cuda().eval()
    enc.load_state_dict(sd)
    pts, _ = synthetic.make_point_clouds(3, 1024, seed=9, dup_fraction=0.5)
    with torch.no_grad():
        got, geo = enc(pts.cuda(), return_geometry=True)
        again = enc(pts.cuda(), geometry=geo)
    want, trace = po.pointnet2_encoder({"pts_encoder." + k: v for k, v in sd.items()}, pts, return_indices=True)
    # geometry is bit-exact
    fps = [t for t in trace if t[0] == "fps"]
    bqs = [t for t in trace if t[0] == "bq"]
    for k in range(4):
        np.testing.assert_array_equal(geo[k][0].cpu().numpy(), fps[k][2].numpy())
        for i in range(2):
            np.testing.assert_array_equal(geo[k][2][i].cpu().numpy(), bqs[2 * k + i][3].numpy())
    assert got.shape == (3, 1024)
    assert torch.equal(got, again)
    scale = want.abs().max()
    assert (got.cpu() - want).abs().max() <= 1e-4 * scale, float((got.cpu() - want).abs().max() / scale)


def test_config2_size_properties_and_determinism():
    """BASELINE config 2: 64 objects x 50 hypotheses, full path, real encoder."""
    pipe = make_pipeline()
    pts, center = synthetic.make_point_clouds(64, 1024, seed=0)
    data = {"pts": pts.cuda(), "pts_center": center.cuda()}
    torch.manual_seed(1)
    out = pipe(data, repeat_num=50, T0=0.55, return_all=True)
    torch.manual_seed(1)
    out2 = pipe({"pts": pts.cuda(), "pts_center": center.cuda()}, repeat_num=50, T0=0.55, return_all=True)
    for k in ("pred_pose", "energy", "aggregated_pose", "length"):
        assert torch.equal(out[k], out2[k]), k  # bitwise deterministic
    pp = out["pred_pose"]
    assert pp.shape == (64, 50, 9) and torch.isfinite(pp).all()
    a, b = pp[..., :3], pp[..., 3:6]
    assert (a.norm(dim=-1) - 1).abs().max() < 1e-12 and (b.norm(dim=-1) - 1).abs().max() < 1e-12
    assert (a * b).sum(-1).abs().max() < 1e-12
    agg = out["aggregated_pose"].double()
    RtR = agg[:, :3, :3].transpose(1, 2) @ agg[:, :3, :3]
    assert (RtR - torch.eye(3, device="cuda", dtype=torch.float64)).abs().max() < 1e-5
    assert out["length"].shape == (64, 3) and torch.isfinite(out["length"]).all()
    # sharding by object gives each object the result of "the reference run once per shard" (SURVEY 8e):
    # within tolerance of the full-batch run because the step controller is shared per call
    from genpose2_b200.pipeline import shard_range
    lo, hi = shard_range(64, 1, 2)
    torch.manual_seed(1)
    noise_full = pipe.score_agent.net.prior_fn((64 * 50, 9), T=0.55)
    pipe.score_agent.net.prior_fn = lambda shape, T=1.0: noise_full[lo * 50: hi * 50].clone()
    sh = pipe({"pts": pts[lo:hi].cuda(), "pts_center": center[lo:hi].cuda()}, repeat_num=50, T0=0.55, return_all=True)
    rot, trans = pose_errors(sh["pred_pose"].cpu().numpy(), pp[lo:hi].cpu().numpy())
    assert rot < 1e-3 and trans < 1e-4, (rot, trans)


def test_tracking_sequence_matches_oracle():
    """BASELINE config 4 (tracking): warm-started short-horizon denoising with the aggregated pose fed
    back as next init_x (evaluation_tracking.py:117-127, 210-214), a few frames, vs the CPU oracle run in
    the same loop with the same noise.  Encoder features are injected on both sides."""
    from genpose2_b200.pipeline import PosePipeline
    B, R, T0, frames = 4, 50, 0.25, 3
    pipe = make_pipeline()
    ssd = synthetic.random_gfobjectpose_state_dict(100)
    esd = synthetic.random_gfobjectpose_state_dict(200)
    csd = synthetic.random_scalenet_state_dict(300)
    g = torch.Generator().manual_seed(77)
    R0 = synthetic._random_rotations(np.random.default_rng(77), B)
    init = torch.zeros(B, 9)
    init[:, :3] = torch.from_numpy(R0[:, :, 0]).float()
    init[:, 3:6] = torch.from_numpy(R0[:, :, 1]).float()
    init_ref, init_dev = init.clone(), init.clone().cuda()
    for f in range(frames):
        pts, center = synthetic.make_point_clouds(B, 1024, seed=500 + f)
        sfeat = torch.relu(torch.randn(B, 1024, generator=g))
        efeat = torch.relu(torch.randn(B, 1024, generator=g))
        noise = torch.randn(B * R, 9, generator=g) * po.ve_marginal_std(T0)
        init_ref[:, 6:] = torch.randn(B, 3, generator=g) * 0.01 if f == 0 else init_ref[:, 6:] - center
        if f == 0:
            init_dev[:, 6:] = init_ref[:, 6:].cuda()
        else:
            init_dev[:, 6:] = init_dev[:, 6:] - center.cuda()
        want = po.full_pipeline(ssd, esd, csd, pts, center, noise, repeat_num=R, T0=T0, init_x=init_ref.clone(),
                                integrator="restated", score_feat=sfeat, energy_feat=efeat)
        inject_features(pipe, sfeat.cuda(), efeat.cuda())
        pipe.score_agent.net.prior_fn = lambda shape, T=1.0, n=noise: n.clone()
        got = pipe({"pts": pts.cuda(), "pts_center": center.cuda()}, repeat_num=R, T0=T0, init_x=init_dev.clone(),
                   return_all=True)
        rot, trans = pose_errors(got["pred_pose"].cpu().numpy(), want["pred_pose"].numpy())
        assert rot <= 1e-3 and trans <= 1e-4, (f, rot, trans)
        a, w = got["aggregated_pose"].cpu().numpy().astype(np.float64), want["aggregated_pose"].numpy().astype(np.float64)
        assert geodesic_mats(a[:, :3, :3], w[:, :3, :3]).max() <= 1e-3 and np.abs(a[:, :3, 3] - w[:, :3, 3]).max() <= 1e-4
        np.testing.assert_allclose(got["length"].cpu().numpy(), want["length"].numpy(), rtol=0, atol=1e-4)
        # feedback: aggregated pose re-encoded as 6D + t, in the camera frame (evaluation_tracking.py:210-214)
        init_ref = PosePipeline.next_init_x(want["aggregated_pose"])
        init_dev = PosePipeline.next_init_x(got["aggregated_pose"])


def test_config5_shard_size_properties():
    """BASELINE config 5: 8192 objects x 50 hypotheses sharded by object over 8 GPUs = 1024 objects
    (51 200 rows) per GPU.  One shard through the sampler + energy + aggregation + ScaleNet (features
    injected: the encoder at this size is covered by the C3 sweep), size-independent properties."""
    B, R = 1024, 50
    pipe = make_pipeline()
    g = torch.Generator().manual_seed(5)
    sfeat = torch.relu(torch.randn(B, 1024, generator=g)).cuda()
    efeat = torch.relu(torch.randn(B, 1024, generator=g)).cuda()
    inject_features(pipe, sfeat, efeat)
    center = (torch.randn(B, 3, generator=g) * 0.1 + torch.tensor([0.0, 0.0, 0.8])).cuda()
    data = {"pts": torch.zeros(B, 1, 3, device="cuda"), "pts_center": center}
    torch.manual_seed(3)
    out = pipe(data, repeat_num=R, T0=0.55, return_all=True)
    from genpose2_b200 import samplers
    st = samplers.ode_stats()
    assert st["status"] == 0 and 100 <= st["nfev"] <= 400, st
    pp = out["pred_pose"]
    assert pp.shape == (B, R, 9) and torch.isfinite(pp).all()
    assert ((pp[..., :3].norm(dim=-1) - 1).abs().max() < 1e-12) and ((pp[..., :3] * pp[..., 3:6]).sum(-1).abs().max() < 1e-12)
    agg = out["aggregated_pose"].double()
    assert ((agg[:, :3, :3].transpose(1, 2) @ agg[:, :3, :3]) - torch.eye(3, device="cuda", dtype=torch.float64)).abs().max() < 1e-5
    assert torch.isfinite(out["energy"]).all() and torch.isfinite(out["length"]).all()
    # an object's hypotheses depend on the other objects only through the shared step controller:
    # the first 64 objects alone (same noise rows) land within tolerance of their values in the big batch
    torch.manual_seed(3)
    noise = pipe.score_agent.net.prior_fn((B * R, 9), T=0.55)
    pipe.score_agent.net.prior_fn = lambda shape, T=1.0: noise[: 64 * R].clone()
    inject_features(pipe, sfeat[:64].contiguous(), efeat[:64].contiguous())
    sub = pipe({"pts": torch.zeros(64, 1, 3, device="cuda"), "pts_center": center[:64].contiguous()}, repeat_num=R,
               T0=0.55, return_all=True)
    # (a different batch composition is a different step-size sequence, so agreement is at the level the
    # rtol = atol = 1e-5 controller delivers per row -- tight for almost every hypothesis, looser for the few
    # trajectories that are sensitive under random weights; the like-for-like comparisons above are the parity gate)
    from tests.util import geodesic_6d
    a, b = sub["pred_pose"].cpu().numpy().reshape(-1, 9), pp[:64].cpu().numpy().reshape(-1, 9)
    rot = geodesic_6d(a[:, :6], b[:, :6])
    trans = np.linalg.norm(a[:, 6:] - b[:, 6:], axis=1)
    print("cross-batch-composition distance: median", np.median(rot), np.median(trans), "max", rot.max(), trans.max())
    assert np.median(rot) < 2e-3 and np.median(trans) < 2e-3, (np.median(rot), np.median(trans))
    assert rot.max() < 0.2, rot.max()


def test_cuda_graph_step_equals_eager_step():
    """PosePipeline(use_graph=True): the whole step (two encoders on two streams, cooperative cluster sampler, energy,
    aggregation, ScaleNet) replayed as one CUDA graph gives bit-identical results to the eager step, consumes the CPU
    generator identically, follows new inputs, and re-captures per input signature."""
    from genpose2_b200.pipeline import PosePipeline
    eager = PosePipeline(device="cuda").load_synthetic_weights()
    graph = PosePipeline(device="cuda", use_graph=True).load_synthetic_weights()
    for B, seed in ((8, 0), (8, 1), (5, 2), (8, 3)):
        pts, center = synthetic.make_point_clouds(B, 1024, seed=40 + seed)
        data = {"pts": pts.cuda(), "pts_center": center.cuda()}
        torch.manual_seed(seed)
        a_pose, a_len = eager(dict(data), repeat_num=50, T0=0.55)
        after_eager = torch.rand(1)
        torch.manual_seed(seed)
        g_pose, g_len = graph(dict(data), repeat_num=50, T0=0.55)
        after_graph = torch.rand(1)
        assert torch.equal(a_pose, g_pose) and torch.equal(a_len, g_len), (B, seed)
        assert torch.equal(after_eager, after_graph)      # same number of draws from the global CPU generator
    assert len(graph._graphs) == 2 and graph.graph_launches > 20
    from genpose2_b200 import samplers
    assert samplers.ode_stats()["status"] == 0
    # tracking signature (init_x given) is its own graph
    B = 4
    pts, center = synthetic.make_point_clouds(B, 1024, seed=50)
    R0 = synthetic._random_rotations(np.random.default_rng(3), B)
    init = torch.zeros(B, 9)
    init[:, :3] = torch.from_numpy(R0[:, :, 0]).float()
    init[:, 3:6] = torch.from_numpy(R0[:, :, 1]).float()
    data = {"pts": pts.cuda(), "pts_center": center.cuda()}
    torch.manual_seed(9)
    a_pose, a_len = eager(dict(data), repeat_num=50, T0=0.25, init_x=init.cuda())
    torch.manual_seed(9)
    g_pose, g_len = graph(dict(data), repeat_num=50, T0=0.25, init_x=init.cuda())
    assert torch.equal(a_pose, g_pose) and torch.equal(a_len, g_len)
