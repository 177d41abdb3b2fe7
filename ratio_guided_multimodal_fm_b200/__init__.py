"""rgfm-b200: B200-native ratio-guided coupled Euler sampler (drop-in for the reference's sampling path)."""
from .sampler import (sample_bimodal_guided, sample_bimodal_guided_mnist_svhn, sample_guidance_sweep, shard_bounds,
                      gather_pairs, clear_plan_cache)
from .flow import CFMSchedule
from .evaluate import (evaluate_coherence, evaluate_coherence_mnist_svhn, classifier_predictions,
                       get_inverse_transform)
from .synthetic import synthetic_module, state_dict as synthetic_state_dict, StateDictModule, TrainableModule
from .train import RatioTrainer, RatioTrainPlan, ratio_loss
from .fm_train import train_flow_matching_epoch, fm_step, library_convs, conv3x3, group_norm_nhwc

__all__ = ["sample_bimodal_guided", "sample_bimodal_guided_mnist_svhn", "sample_guidance_sweep", "CFMSchedule", "shard_bounds", "gather_pairs",
           "clear_plan_cache", "evaluate_coherence", "evaluate_coherence_mnist_svhn", "classifier_predictions",
           "get_inverse_transform", "synthetic_module", "synthetic_state_dict", "StateDictModule", "TrainableModule",
           "RatioTrainer", "RatioTrainPlan", "ratio_loss", "train_flow_matching_epoch", "fm_step", "library_convs", "conv3x3", "group_norm_nhwc"]
