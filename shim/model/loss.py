"""Overlay for RAM_Net/model/loss.py (every public name of :1-70; lstm_trainer.py:5 imports mse_loss and
multi_scale_grad_loss, train.py:7 star-imports the rest and eval()s config['loss']['type'])."""
from rpg_ramnet_b200.model.loss import (MultiScaleGradient, SILossBatch, mse_loss, multi_scale_grad_loss,  # noqa: F401
                                        multi_scale_grad_loss_fn, scale_invariant_log_loss, scale_invariant_loss)

__all__ = ['scale_invariant_loss', 'scale_invariant_log_loss', 'mse_loss', 'MultiScaleGradient',
           'multi_scale_grad_loss_fn', 'multi_scale_grad_loss', 'SILossBatch']
