"""Stand-in for `pytorch_lightning`, used ONLY by oracle/gen_golden.py.

The reference (kaiidams/voice100) subclasses `pl.LightningModule` /
`pl.LightningDataModule` (voice100/models/_base.py:3-7,
voice100/data_modules.py:499-500) but its inference path uses nothing from
Lightning except `save_hyperparameters()` and the no-op loggers.  Lightning is
not installable in this image (no network), so this shim lets the UNMODIFIED
reference import from /root/reference while the golden vectors are generated.
It is test infrastructure: nothing in the product path imports it.
"""
import types

import torch
from torch import nn


class _HParams(dict):
    __getattr__ = dict.get


class LightningModule(nn.Module):
    def save_hyperparameters(self, *args, **kwargs):
        import inspect
        frame = inspect.currentframe().f_back
        names = frame.f_code.co_varnames[1:frame.f_code.co_argcount]
        self.hparams = _HParams({k: frame.f_locals[k] for k in names})

    def log(self, *a, **k):
        pass

    def log_dict(self, *a, **k):
        pass


class LightningDataModule:
    pass


class Trainer:
    pass


def seed_everything(seed):
    torch.manual_seed(seed)
    return seed


callbacks = types.SimpleNamespace(ModelCheckpoint=object, LearningRateMonitor=object)
