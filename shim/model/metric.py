"""Overlay for RAM_Net/model/metric.py: same function names (train.py:8 star-imports them, train.py:226 eval()s the
names in config['metrics']); each accepts what LSTMTrainer._eval_metrics passes (numpy arrays, lstm_trainer.py:100-106)
or tensors, and reduces on the device."""
from rpg_ramnet_b200.model.metric import (abs_rel_diff, eval_metrics, mean_error, median_error, mse, rms_linear,  # noqa: F401
                                          scale_invariant_error, squ_rel_diff)

__all__ = ['abs_rel_diff', 'squ_rel_diff', 'rms_linear', 'scale_invariant_error', 'mean_error', 'median_error', 'mse',
           'eval_metrics']
