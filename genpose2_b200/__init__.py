"""genpose2_b200 -- B200-native (sm_100a) implementation of GenPose++'s per-object
pose-generation hot path behind the reference's own Python API.

Mirror of the reference interface (reference file -> module here):
    networks/posenet_agent.py              -> posenet_agent.PoseNet
    networks/posenet.py                    -> posenet.GFObjectPose
    networks/gf_algorithms/samplers.py     -> samplers.cond_ode_sampler / cond_pc_sampler
    networks/gf_algorithms/scorenet.py     -> scorenet.PoseScoreNet
    networks/gf_algorithms/energynet.py    -> scorenet.PoseEnergyNet
    networks/gf_algorithms/sde.py          -> sde.init_sde
    networks/scalenet.py                   -> scalenet.ScaleNet
    networks/reward.py + runner block      -> aggregation.sort_poses_by_energy / aggregate_pose
    networks/pts_encoder/pointnet2.py      -> pointnet2.Pointnet2ClsMSG
    .../pointnet2/pointnet2_utils.py       -> pointnet2_utils.*
    configs/config.py                      -> config.get_config

All arithmetic on the path runs in genpose2_b200/libgenpose_b200.so (C ABI: include/genpose_b200.h).
"""
__version__ = "0.1.0"
