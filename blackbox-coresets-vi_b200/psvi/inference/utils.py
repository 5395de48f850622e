"""Hot-path-adjacent helpers of the reference's psvi/inference/utils.py: pseudo-data initialisers (:33-77),
make_dataloader (:144-148), compute_empirical_mean (:151-161), LogResource (:1752-1781).  The coreset-selection zoo of
that file (k-means / faiss / submodular) is out of scope."""
from __future__ import annotations

import time

import numpy as np
import torch
from torch.utils.data import DataLoader


def pseudo_subsample_init(x, y, num_pseudo=20, nc=2, seed=0):
    """Class-balanced random subset (reference :33-50, same RNG calls)."""
    torch.manual_seed(seed)
    N, _ = x.shape
    cnt = 0
    u, z = torch.Tensor([]), torch.Tensor([])
    for c in range(nc):
        idx_c, pts_with_c = torch.arange(N)[y == c], (num_pseudo // nc if c < nc - 1 else num_pseudo - cnt)
        u = torch.cat((u, x[idx_c[torch.randperm(len(idx_c))[:pts_with_c]]]))
        z = torch.cat((z, c * torch.ones(pts_with_c)))
        cnt += num_pseudo // nc
    return u.requires_grad_(True), z


def pseudo_rand_init(x, y, num_pseudo=20, nc=2, seed=0, variance=0.1):
    """Noisy data mean + labels split equally among classes (reference :53-77)."""
    torch.manual_seed(seed)
    _, D = x.shape
    u = (x[:, :].mean() + variance * torch.randn(num_pseudo, D)).clone().requires_grad_(True)
    z = torch.Tensor([])
    for c in range(nc):
        z = torch.cat((z, c * torch.ones(num_pseudo // nc if c < nc - 1 else num_pseudo - (nc - 1) * (num_pseudo // nc))))
    return u, z


def make_dataloader(data, minibatch, shuffle=True):
    return DataLoader(data, batch_size=minibatch, pin_memory=torch.cuda.is_available(), shuffle=shuffle)


def compute_empirical_mean(dloader):
    trainsum, nb_samples = 0.0, 0.0
    for data, _ in dloader:
        batch_samples = data.size(0)
        data = data.view(batch_samples, data.size(1), -1)
        trainsum += data.mean(2).sum(0)
        nb_samples += batch_samples
    return trainsum / nb_samples


class LogResource:
    """Wall time per outer step and allocated device memory (reference :1752-1781)."""

    def __init__(self):
        self.curr_time = time.time()
        self.prev_time = time.time()
        self.time_per_epoch = []
        self.memory_per_epoch = []
        self.device = "cuda" if torch.cuda.is_available() else "cpu"

    def update(self):
        self.prev_time = self.curr_time
        self.curr_time = time.time()
        self.time_per_epoch.append(self.curr_time - self.prev_time)
        self.memory_per_epoch.append(torch.cuda.memory_allocated(0) / 1024 ** 3 if self.device == "cuda" else 0)

    def get_resources(self):
        if not self.time_per_epoch:
            return {"time": float("nan"), "memory": float("nan")}
        return {"time": float(np.mean(self.time_per_epoch)), "memory": float(np.mean(self.memory_per_epoch))}
