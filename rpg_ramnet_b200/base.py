"""Host-side base class mirroring the reference's BaseModel surface (RAM_Net/base/base_model.py:6-30):
keeps `config`, a class-named logger and `summary()` so reference drivers (train.py:204-217,
test.py:170-181) can use the model object unchanged."""
import logging

import torch.nn as nn


class BaseModel(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.logger = logging.getLogger(self.__class__.__name__)

    def forward(self, *inputs):
        raise NotImplementedError

    def summary(self):
        n = sum(p.numel() for p in self.parameters() if p.requires_grad)
        self.logger.info('Trainable parameters: {}'.format(n))
        self.logger.info(self)
