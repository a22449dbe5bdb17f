"""Synthetic datasets and a bf16 trainer mode for running the reference's OWN `trainer.py` / `main.py` without the private
ABCD data -- the hooks a maintainer needs to measure a training step of any reference model on a B200.

  use_synthetic_data()   patches `DataHandler.get_dataset` (dataloaders.py:30-50) so that every `--dataset_name` of the
                         reference resolves to `SyntheticDataset`, which serves seeded N(0,1) tensors under the same
                         dictionary keys, shapes and dtypes as the ABCD readers (datasets.py:171-366, 479-545, 607-702), plus
                         `index_l` (dataloaders.py:61) and `get_input_shape()` (trainer.py:177).
  use_bf16_autocast()    replaces `trainer.autocast` (fp16, trainer.py:24,378) by bf16 autocast and `trainer.GradScaler`
                         (trainer.py:25,84) by a disabled scaler: bf16 needs no loss scaling, `scaler.step/update` stay callable.

Call both after `install()` and before `Trainer(...)`; nothing under /root/reference is modified.
"""
from __future__ import annotations

import functools
import sys

import numpy as np
import torch
from torch.utils.data import Dataset

# dictionary keys per reference dataset name: (key -> shape builder)
_SEQ = 368        # padded time series length (datasets.py:666-673)


class SyntheticDataset(Dataset):
    """Stand-in for the reference's dataset classes.  Shapes: time series (368, intermediate_vec) float32, structural
    matrices (intermediate_vec, intermediate_vec) float16 (the readers call `.half()`, datasets.py:541-542,663)."""

    def __init__(self, **kwargs):
        self.dataset_name = kwargs.get("dataset_name")
        self.fmri_type = kwargs.get("fmri_type")
        self.target = kwargs.get("target")
        self.fine_tune_task = kwargs.get("fine_tune_task")
        self.augment = None
        self.roi = int(kwargs.get("intermediate_vec") or 84)
        self.n = int(kwargs.get("synthetic_samples") or 256)
        self.seed = int(kwargs.get("seed") or 0)
        # (index, subject name, path placeholder, ..., target): the reference reads [2] of entry 0 for the input shape
        self.index_l = [(i, f"SYNTH{i:06d}", None, None, self._target(i)) for i in range(self.n)]

    def _target(self, i):
        g = torch.Generator().manual_seed(self.seed * 1_000_003 + i)
        if self.fine_tune_task == "regression":
            return torch.randn((), generator=g)
        return torch.bernoulli(torch.tensor(0.5), generator=g)

    def get_input_shape(self):
        return (self.roi, self.roi) if self.dataset_name in ("struct", "DTI", "sMRI", "DTI+sMRI") else (_SEQ + 2, self.roi)

    def __len__(self):
        return self.n

    def __getitem__(self, index):
        i, name, _, _, target = self.index_l[index]
        g = torch.Generator().manual_seed(self.seed * 7_000_003 + i)
        seq = lambda: torch.randn(_SEQ, self.roi, generator=g)
        mat = lambda: torch.randn(self.roi, self.roi, generator=g).half()
        d = {"subject": i, "subject_name": name, self.target: target}
        ds = self.dataset_name
        if ds == "struct":
            d.update(smri=mat(), dti=mat())
        elif ds == "DTI":
            d.update(dti=mat())
        elif ds == "sMRI":
            d.update(smri=mat())
        elif ds == "DTI+sMRI":
            d.update(struct=mat())
        elif ds in ("multimodal", "multimodal_prs"):
            d.update(fmri_raw_sequence=seq(), fmri_lowfreq_sequence=seq(), fmri_ultralowfreq_sequence=seq(), struct=mat())
            if ds == "multimodal_prs":
                d.update(prs=torch.randn(1, generator=g))
        else:                                      # fMRI_timeseries / hcp
            if self.fmri_type in ("divided_frequency", "timeseries_and_frequency"):
                d.update(fmri_sequence=seq(), fmri_lowfreq_sequence=seq(), fmri_ultralowfreq_sequence=seq())
            else:
                d.update(fmri_sequence=seq())
        return d


def use_synthetic_data(dataloaders_module=None):
    """Make the reference's DataHandler serve SyntheticDataset for every dataset name.  Returns the patched class."""
    mod = dataloaders_module or sys.modules.get("data_preprocess_and_load.dataloaders")
    if mod is None:
        import importlib
        mod = importlib.import_module("data_preprocess_and_load.dataloaders")
    mod.DataHandler.get_dataset = lambda self: SyntheticDataset

    def split(self, index_l, **kwargs):
        """Seeded random subject split by position (dataloaders.py:151-165 matches subject ids as strings against integer
        draws, which numpy >= 2 no longer equates; the synthetic subjects are simply 0..n-1)."""
        n = len(index_l)
        rng = np.random.RandomState(int(kwargs.get("seed") or 0))
        perm = rng.permutation(n)
        n_train, n_val = int(n * kwargs.get("train_split")), int(n * kwargs.get("val_split"))
        return perm[:n_train].tolist(), perm[n_train:n_train + n_val].tolist(), perm[n_train + n_val:].tolist()

    mod.DataHandler.determine_split_randomly = split
    return mod.DataHandler


class _NoScaler:
    """GradScaler with scaling disabled (bf16 has fp32's exponent range): the trainer's calls stay valid."""

    def __init__(self, *a, **k):
        self._inner = torch.amp.GradScaler("cuda", enabled=False)

    def __getattr__(self, name):
        return getattr(self._inner, name)


def use_bf16_autocast(trainer_module=None):
    mod = trainer_module or sys.modules.get("trainer")
    if mod is None:
        import importlib
        mod = importlib.import_module("trainer")
    mod.autocast = functools.partial(torch.autocast, "cuda", dtype=torch.bfloat16)
    mod.GradScaler = _NoScaler
    return mod


def seed_everything(seed: int = 0):
    np.random.seed(seed)
    torch.manual_seed(seed)
