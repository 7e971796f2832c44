from .losses import (STANDARD_H36M_WEIGHTS, STANDARD_HEVA_WEIGHTS, weighted_mse_loss, weighted_mpjpe_loss, mean_velocity_error,
                     wta_with_scoring_loss, wta_l2_loss_and_activate_head, training_loss)
from .regularizations import (smoothness_regularization, measure_bones_length, segments_time_consistency,
                              segments_time_consistency_per_bone, sagittal_symmetry, sagittal_symmetry_per_bone)
from .mean_joint_errors import (mpjpe_error, mse_error, jointwise_error, jointwise_mse, coordwise_error, segments_len_err, p_mpjpe)
from .pck import keypoint_3d_pck, keypoint_3d_auc

__all__ = ["STANDARD_H36M_WEIGHTS", "STANDARD_HEVA_WEIGHTS", "weighted_mse_loss", "weighted_mpjpe_loss", "mean_velocity_error",
           "wta_with_scoring_loss", "wta_l2_loss_and_activate_head", "training_loss", "smoothness_regularization", "mpjpe_error", "mse_error", "jointwise_error", "jointwise_mse", "coordwise_error", "segments_len_err", "p_mpjpe", "keypoint_3d_pck", "keypoint_3d_auc",
           "measure_bones_length", "segments_time_consistency", "segments_time_consistency_per_bone", "sagittal_symmetry",
           "sagittal_symmetry_per_bone"]
