"""GPU: energy-ranked outlier rejection + quaternion averaging + DBSCAN (gp_aggregate) and the
ScaleNet head (gp_scalenet) vs the golden fixtures produced by the reference functions + sklearn."""
import numpy as np
import pytest
import torch

from genpose2_b200 import synthetic
from oracle import pose_oracle as po
from tests.util import geodesic_mats, load_golden

pytestmark = pytest.mark.gpu


def test_sort_poses_by_energy_matches_reference():
    from genpose2_b200.aggregation import sort_poses_by_energy
    g = load_golden("aggregate_clusters")
    sp, se = sort_poses_by_energy(torch.from_numpy(g["poses"]).cuda(), torch.from_numpy(g["energy"]).cuda())
    np.testing.assert_array_equal(sp.cpu().numpy(), g["sorted_pose"])
    np.testing.assert_array_equal(se.cpu().numpy(), g["sorted_energy"])


def test_sort_ties_are_stable_descending():
    """torch.sort(stable=False) leaves the order of equal energies unspecified; gp_aggregate is
    stable-descending (equal keys keep hypothesis order)."""
    from genpose2_b200.aggregation import sort_poses_by_energy
    poses = torch.arange(2 * 10 * 9, dtype=torch.float64).reshape(2, 10, 9).cuda()
    energy = torch.tensor([3., 1., 3., 2., 1., 3., 0., 2., 9., 1.]).repeat(2, 1).unsqueeze(-1).repeat(1, 1, 2).cuda()
    energy[1, :, 1] = -energy[1, :, 1]
    sp, se = sort_poses_by_energy(poses, energy)
    order = [8, 0, 2, 5, 3, 7, 1, 4, 9, 6]
    assert sp[0, :, 0].tolist() == [poses[0, i, 0].item() for i in order]
    assert sp[0, :, 8].tolist() == [poses[0, i, 8].item() for i in order]
    rev = [6, 1, 4, 9, 3, 7, 0, 2, 5, 8]
    assert sp[1, :, 8].tolist() == [poses[1, i, 8].item() for i in rev]
    assert sp[1, :, 5].tolist() == [poses[1, i, 5].item() for i in order]


def test_aggregate_clusters_matches_reference():
    from genpose2_b200.aggregation import aggregate_pose
    g = load_golden("aggregate_clusters")
    out, labels = aggregate_pose(torch.from_numpy(g["poses"]).cuda(), torch.from_numpy(g["energy"]).cuda(),
                                 eval_repeat_num=50, return_labels=True)
    np.testing.assert_array_equal(labels.cpu().numpy(), g["labels"])
    want = g["aggregated_pose"]
    got = out.cpu().numpy()
    assert got.dtype == np.float32 and got.shape == want.shape
    assert geodesic_mats(got[:, :3, :3].astype(np.float64), want[:, :3, :3].astype(np.float64)).max() < 1e-5
    np.testing.assert_allclose(got[:, :3, 3], want[:, :3, 3], rtol=0, atol=1e-6)
    np.testing.assert_array_equal(got[:, 3], np.tile(np.array([0, 0, 0, 1], np.float32), (got.shape[0], 1)))


def test_aggregate_random_vs_oracle_many_objects():
    """many objects, several jitter levels (cluster / no cluster / multi-cluster), no-clustering flag"""
    from genpose2_b200.aggregation import aggregate_pose
    for seed, jitter in ((1, 1.0), (2, 4.0), (3, 12.0)):
        poses = synthetic.make_cluster_quaternion_poses(40, 50, seed=seed, jitter_deg=jitter)
        energy = torch.randn(40, 50, 2, generator=torch.Generator().manual_seed(seed))
        for clustering in (1, 0):
            want, wl = po.aggregate_pose(poses, energy, clustering=bool(clustering))
            got, gl = aggregate_pose(poses.cuda(), energy.cuda(), eval_repeat_num=50, clustering=clustering,
                                     return_labels=True)
            if clustering:
                np.testing.assert_array_equal(gl.cpu().numpy(), wl)
            g, w = got.cpu().numpy().astype(np.float64), want.numpy().astype(np.float64)
            assert geodesic_mats(g[:, :3, :3], w[:, :3, :3]).max() < 1e-5
            np.testing.assert_allclose(g[:, :3, 3], w[:, :3, 3], rtol=0, atol=1e-6)


def test_scalenet_matches_reference_golden():
    from genpose2_b200.scalenet import ScaleNet
    g = load_golden("scalenet_b5")
    net = ScaleNet(1024, 0, 180).cuda()
    net.load_state_dict(synthetic.random_scalenet_state_dict(int(g["scale_seed"])))
    out = net({"pts_feat": torch.from_numpy(g["feat"]).cuda(), "axes": torch.from_numpy(g["axes"]).cuda()})
    np.testing.assert_allclose(out.cpu().numpy(), g["length"], rtol=2e-5, atol=2e-6)
    # strided axes view of a [B,4,4] pose
    pose = torch.zeros(5, 4, 4, device="cuda")
    pose[:, :3, :3] = torch.from_numpy(g["axes"]).cuda()
    out2 = net({"pts_feat": torch.from_numpy(g["feat"]).cuda(), "axes": pose[:, :3, :3]})
    assert torch.equal(out, out2)
    # larger batch path (8 objects per block)
    B = 1300
    gen = torch.Generator().manual_seed(0)
    feat = torch.relu(torch.randn(B, 1024, generator=gen))
    axes = torch.from_numpy(synthetic._random_rotations(np.random.default_rng(0), B)).float()
    want = po.scalenet_forward(synthetic.random_scalenet_state_dict(int(g["scale_seed"])), axes, feat)
    got = net({"pts_feat": feat.cuda(), "axes": axes.cuda()}).cpu()
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=2e-5, atol=2e-6)


def test_average_over_all_50_hypotheses_matches_oracle():
    """return_average_res of pred_func (posenet_agent.py:561-570): average_quaternion_batch over all repeat_num = 50
    hypotheses + mean translation, no energies, no clustering (retain > 32 is allowed without clustering)"""
    from genpose2_b200.aggregation import _run
    from oracle import pose_oracle as po
    g = torch.Generator().manual_seed(5)
    B, R = 7, 50
    base = torch.randn(B, 1, 6, generator=g, dtype=torch.float64)
    rot6 = base + 0.05 * torch.randn(B, R, 6, generator=g, dtype=torch.float64)
    trans = torch.randn(B, R, 3, generator=g, dtype=torch.float64)
    # normalise to valid (x-axis, y-axis) pairs like the sampler output
    a1 = torch.nn.functional.normalize(rot6[..., :3], dim=-1)
    a2 = rot6[..., 3:] - (a1 * rot6[..., 3:]).sum(-1, keepdim=True) * a1
    a2 = torch.nn.functional.normalize(a2, dim=-1)
    poses = torch.cat([a1, a2, trans], dim=-1)
    out, _, _ = _run(poses.cuda(), torch.zeros(B, R, 2).cuda(), R, False, 0.0, 1)
    quat = po.matrix_to_quaternion(po.get_rot_matrix(poses.reshape(B * R, 9)[:, :6])).reshape(B, R, 4)
    want_q = po.average_quaternion_batch(quat)
    want_R = po.quaternion_to_matrix(want_q)
    from tests.util import geodesic_mats
    rot_err = float(np.max(geodesic_mats(out[:, :3, :3].double().cpu().numpy(), want_R.numpy())))
    t_err = float((out[:, :3, 3].double().cpu() - trans.mean(1)).abs().max())
    assert rot_err < 1e-6 and t_err < 1e-6, (rot_err, t_err)
