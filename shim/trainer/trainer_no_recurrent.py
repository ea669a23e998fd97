"""train.py:12 imports `trainer.trainer_no_recurrent.TrainerNoRecurrent`, a module the reference does not ship
(`ModuleNotFoundError` on import); its only use is commented out (train.py:228-234) and every arch,
ERGB2Depth included, is trained by LSTMTrainer (train.py:236-243).  The alias makes train.py importable."""
from trainer.lstm_trainer import LSTMTrainer


class TrainerNoRecurrent(LSTMTrainer):
    pass
