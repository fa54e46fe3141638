"""On-disk stage files of the reference's evaluation runner, so that a run of the
accelerated path can be picked up by the reference's later stages (or the other way round):

    score_prediction_<model>.pkl   pickle of (all_pred_pose, all_score_feature)            evaluation_single.py:105-120
                                   all_pred_pose[i]     [bs, R, 9] f64 (as pred_func returns it)
                                   all_score_feature[i] {"pts_feat": [bs,1024] f32 cpu, "rgb_feat": None}
    energy_prediction_<model>.pkl  pickle of all_pred_energy, [i] = [bs, R, 2] f32 cpu      evaluation_single.py:147-157
    aggregated.pkl                 pickle of all_aggregated_pose, [i] = [bs, 4, 4] f32      evaluation_single.py:160-219
    scale_prediction_<model>.pkl   pickle of (all_aggregated_pose, all_final_length)        evaluation_single.py:222-288
                                   all_final_length[i]  [bs, 3] f32 cpu

Plain `pickle` files of python lists of torch tensors -- exactly what the reference's `pickle.load(open(path, "rb"))`
call sites read (evaluation_single.py:127,164-167,226-228,295).  Nothing here touches the GPU path itself.
"""
import os
import pickle

import torch


def _dump(obj, path):
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "wb") as f:
        pickle.dump(obj, f)


def _load(path):
    with open(path, "rb") as f:
        return pickle.load(f)


def save_score_stage(path, all_pred_pose, all_pts_feat):
    """all_pred_pose: list of [bs,R,9] f64; all_pts_feat: list of [bs,1024] f32 (moved to the CPU like the reference)."""
    feats = [{"pts_feat": f.detach().cpu(), "rgb_feat": None} for f in all_pts_feat]
    _dump((list(all_pred_pose), feats), path)


def load_score_stage(path):
    all_pred_pose, all_score_feature = _load(path)
    return all_pred_pose, all_score_feature


def save_energy_stage(path, all_pred_energy):
    _dump([e.detach().cpu() for e in all_pred_energy], path)


def load_energy_stage(path):
    return _load(path)


def save_aggregate_stage(path, all_aggregated_pose):
    _dump(list(all_aggregated_pose), path)


def load_aggregate_stage(path):
    return _load(path)


def save_scale_stage(path, all_aggregated_pose, all_final_length):
    _dump((list(all_aggregated_pose), [l.detach().cpu() for l in all_final_length]), path)


def load_scale_stage(path):
    return _load(path)


def bbox_length_from_points(pcl, aggregated_pose):
    """The runner's fallback when no scale model is given (evaluation_single.py:232-250): points into the object
    frame of the aggregated pose, 2 * max |coordinate|.  pcl [bs,n,3], aggregated_pose [bs,4,4] -> [bs,3]."""
    rotation_t = aggregated_pose[:, :3, :3].transpose(1, 2).to(pcl.dtype)
    local = torch.matmul(pcl - aggregated_pose[:, :3, 3].to(pcl.dtype).unsqueeze(1), rotation_t.transpose(1, 2))
    return 2.0 * local.abs().amax(dim=1)


def stage_paths(result_dir, score_model_name, energy_model_name=None, scale_model_name=None):
    """File names of evaluation_single.py:403-419."""
    root = os.path.join("results", "evaluation_results", result_dir)
    return dict(score=os.path.join(root, f"score_prediction_{score_model_name}.pkl"),
                energy=os.path.join(root, f"energy_prediction_{energy_model_name}.pkl"),
                aggregate=os.path.join(root, "aggregated.pkl"),
                scale=os.path.join(root, f"scale_prediction_{scale_model_name}.pkl"))
