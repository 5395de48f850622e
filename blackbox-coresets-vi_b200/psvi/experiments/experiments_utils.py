"""Dataset / model helpers adjacent to the PSVI hot path -- mirrors the parts of the reference's
psvi/experiments/experiments_utils.py that the BASELINE configs touch: SynthDataset (:81-105),
make_four_class_dataset (:299-343, same RNG call order), set_up_model for the baselines (:346-413, incl. the
missing-n_layers quirk) and read_dataset for the generated datasets (:752-834).  Downloaded datasets (UCI, webspam,
torchvision) are out of scope: there is no network."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
from torch.utils.data import Dataset

from psvi.models.neural_net import (VILinear, make_fc2net, make_fcnet, make_lenet, make_logistic_regression,
                                    make_regressor_net)


class BaseDataset(Dataset):
    """(x, y) tensor dataset of the regression flows (reference experiments_utils.py: BaseDataset)."""

    def __init__(self, x, y=None, randomize=False):
        self.data = x.mean() + 1.0 * torch.randn_like(x) if randomize else x
        self.targets = y

    def __len__(self):
        return len(self.data)

    def __getitem__(self, index):
        return self.data[index], self.targets[index]


class SynthDataset(Dataset):
    """Tensor dataset with `.data` / `.targets` (reference :81-105)."""

    def __init__(self, x, y=None, transforms=None):
        self.data = x
        self.targets = y
        self.transforms = transforms

    def __len__(self):
        return len(self.data)

    def __getitem__(self, index):
        return self.data[index], self.targets[index]

    def subset_where(self, cs=[0, 1]):
        idcs = torch.isin(self.targets, torch.tensor(cs))
        return SynthDataset(self.data[idcs], self.targets[idcs])

    def concatenate(self, u, z):
        return SynthDataset(torch.cat((self.data, u)), y=torch.cat((self.targets, z)))


def make_four_class_dataset(N_K=250):
    """four_blobs (reference :299-343).  Draws from the global torch RNG in the reference's order."""
    X1 = torch.cat([0.8 + 0.4 * torch.randn(N_K, 1), 1.5 + 0.4 * torch.randn(N_K, 1)], dim=-1)
    Y1 = 0 * torch.ones(X1.size(0)).long()
    X2 = torch.cat([0.5 + 0.6 * torch.randn(N_K, 1), -0.2 - 0.1 * torch.randn(N_K, 1)], dim=-1)
    Y2 = 1 * torch.ones(X2.size(0)).long()
    X3 = torch.cat([2.5 - 0.1 * torch.randn(N_K, 1), 1.0 + 0.6 * torch.randn(N_K, 1)], dim=-1)
    Y3 = 2 * torch.ones(X3.size(0)).long()
    X4 = torch.distributions.MultivariateNormal(
        torch.Tensor([-0.5, 1.5]), covariance_matrix=torch.Tensor([[0.2, 0.1], [0.1, 0.1]])).sample(torch.Size([N_K]))
    Y4 = 3 * torch.ones(X4.size(0)).long()
    X = torch.cat([X1, X2, X3, X4], dim=0)
    X[:, 1] -= 1
    X[:, 0] -= 0.5
    Y = torch.cat([Y1, Y2, Y3, Y4])
    perm = torch.randperm(X.size()[0])
    return X[perm, :], Y[perm]


def set_up_model(D=None, n_hidden=None, nc=None, mc_samples=None, architecture=None, **kwargs):
    """Model factory used by the baselines (reference :346-413).  NB it does not forward n_layers, so "fn" gets
    make_fcnet's default of TWO hidden layers here (unlike PSVI.set_up_model) -- kept, it changes the model."""
    if architecture in {"fn", "residual_fn"}:
        return make_fcnet(D, n_hidden, nc, linear_class=VILinear, nonl_class=nn.ReLU, mc_samples=mc_samples,
                          residual=(architecture == "residual_fn"), **kwargs)
    elif architecture in {"fn2"}:
        return make_fc2net(D, n_hidden, nc, mc_samples=mc_samples, **kwargs)
    elif architecture == "lenet":
        return make_lenet(mc_samples=mc_samples)
    elif architecture == "logistic_regression":
        return make_logistic_regression(D, nc, mc_samples=mc_samples)
    elif architecture == "regressor_net":
        return make_regressor_net(D, n_hidden, nc, linear_class=VILinear, nonl_class=nn.ReLU, mc_samples=mc_samples, **kwargs)
    raise ValueError("Architecture should be one of \n'lenet', 'logistic_regression', "
                     "'logistic_regression_fullcov', 'fn', 'fn2', 'residual_fn'")


def make_synthetic_rows(n_rows, D, nc, seed=0, device="cpu"):
    """Seeded synthetic classification rows of the BASELINE shapes: X ~ N(0, I),
    y = argmax(X W*) for a fixed random W*."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    W = torch.randn(D, nc, generator=g)
    X = torch.randn(n_rows, D, generator=g)
    y = (X @ W).argmax(-1)
    return X.to(device), y.to(device)


def read_dataset(dnm, method_args):
    """halfmoon / four_blobs exactly as the reference generates them (:759-804); returns
    (x, y, xt, yt, N, D, train_dataset, test_dataset, num_classes)."""
    if dnm == "halfmoon":
        from sklearn.datasets import make_moons
        (X, Y), num_classes = make_moons(n_samples=1000, noise=0.1, random_state=42), 2
        X, Y = torch.from_numpy(X.astype(np.float32)), torch.from_numpy(Y.astype(np.float32))
    elif dnm == "four_blobs":
        (X, Y), num_classes = make_four_class_dataset(N_K=250), 4
    elif dnm.startswith("synth_rows"):
        # synth_rows_<N>_<D>_<C>: bench / scaling shapes without any download
        _, _, n, d, c = dnm.split("_")
        (X, Y), num_classes = make_synthetic_rows(int(n), int(d), int(c)), int(c)
    else:
        raise NotImplementedError(f"dataset {dnm!r} needs a download or a loader that is out of the hot-path scope "
                                  "(no network in this build; SURVEY.md section 2 row 9)")
    Y[Y == -1] = 0
    test_size = int(method_args["test_ratio"] * X.shape[0])
    x, y, xt, yt = X[:-test_size], Y[:-test_size], X[-test_size:], Y[-test_size:]
    N, D = x.shape
    return x, y, xt, yt, N, D, SynthDataset(x, y), SynthDataset(xt, yt), num_classes
