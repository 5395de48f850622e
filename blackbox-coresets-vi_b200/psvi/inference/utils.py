"""Hot-path-adjacent helpers of the reference's psvi/inference/utils.py: pseudo-data initialisers (:33-77),
make_dataloader (:144-148), compute_empirical_mean (:151-161), MeanFieldVI (:221-450), LogResource (:1752-1781).  The
coreset-selection zoo of that file (k-means / faiss / submodular) is out of scope."""
from __future__ import annotations

import os
import random
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


class MeanFieldVI:
    """Mean-field VI on the full training set, sequential minibatches, as a class (reference psvi/inference/utils.py:221-450:
    "same as run_mfvi, but puts it inside a class"), with the forgetting-score bookkeeping (:356-385,401-406) and the
    save / load of the fitted net (:425-450).  Every minibatch step is the fused sampled forward + backward + Adam kernel
    of the baselines (psvi/inference/baselines.py: psvi_mf_unroll with torch.optim.Adam arithmetic; fn2 / lenet through the
    streaming engine); the per-row accuracies of `after_epoch` come from the module forward (`net(x).mean(0)`)."""

    def __init__(self, xt=None, yt=None, mc_samples=4, data_minibatch=128, num_epochs=100, log_every=10, N=None, D=None,
                 lr0net=1e-3, mul_fact=2, seed=0, distr_fn=None, architecture=None, n_hidden=None, nc=2, log_pseudodata=False,
                 train_dataset=None, test_dataset=None, init_sd=None, forgetting_score_flag=False, data_path=None,
                 load_from_saved=False, dnm=None, noise_source=None, **kwargs):
        self.mc_samples, self.data_minibatch, self.num_epochs, self.log_every = mc_samples, data_minibatch, num_epochs, log_every
        self.N, self.D, self.lr0net, self.seed, self.distr_fn = N, D, lr0net, seed, distr_fn
        self.architecture, self.n_hidden, self.nc, self.log_pseudodata = architecture, n_hidden, nc, log_pseudodata
        self.train_dataset, self.test_dataset, self.init_sd, self.mul_fact = train_dataset, test_dataset, init_sd, mul_fact
        self.forgetting_score_flag, self.data_path, self.load_from_saved, self.dnm = (forgetting_score_flag, data_path,
                                                                                      load_from_saved, dnm)
        self.net_state_dict_fname, self.forgetting_fname = f"net_state_dict_{seed}.pt", f"forgetting_{seed}.pt"
        self.noise_source = noise_source          # exact-noise hook of the parity tests (None: Philox streams)
        self.quiet = bool(kwargs.get("quiet", False))

    # ---- reference :300-339
    def before_train(self):
        from psvi import _native
        from psvi.experiments.experiments_utils import set_up_model
        from psvi.inference.baselines import _make_trainer
        _native.require_cuda()
        self.device = torch.device("cuda")
        random.seed(self.seed), np.random.seed(self.seed), torch.manual_seed(self.seed)
        self.net = set_up_model(architecture=self.architecture, D=self.D, n_hidden=self.n_hidden, nc=self.nc,
                                mc_samples=self.mc_samples, init_sd=self.init_sd).to(self.device)
        self._trainer = _make_trainer(self.net, self.seed, self.noise_source)
        ds, dt = self.train_dataset, self.test_dataset
        self._x = torch.as_tensor(ds.data).to(self.device, torch.float32).reshape(len(ds), -1).contiguous()
        self._y = torch.as_tensor(ds.targets).to(self.device).to(torch.int32).contiguous()
        self._xt = torch.as_tensor(dt.data).to(self.device, torch.float32).reshape(len(dt), -1).contiguous()
        self._yt = torch.as_tensor(dt.targets).to(self.device).to(torch.int32).contiguous()
        self.n_train = self._x.shape[0]
        self.total_iterations = self.mul_fact * self.num_epochs
        self.nlls_mfvi, self.accs_mfvi, self.times_mfvi, self.elbos_mfvi = [], [], [0], []
        self.t_start = time.time()
        self.forgetting_events = torch.zeros(self.n_train, device=self.device)
        self.last_acc = torch.zeros(self.n_train, device=self.device)
        self.never_learnt_events = torch.ones(self.n_train, device=self.device)

    def _batches(self):
        B = int(self.data_minibatch)
        return [(r0, min(r0 + B, self.n_train)) for r0 in range(0, self.n_train, B)]

    # ---- reference :274-296: one Adam step per sequential minibatch on  -(n_train / B) sum log p + sum_VILinear KL
    def train_an_epoch(self):
        for r0, r1 in self._batches():
            losses = self._trainer.train(self._x[r0:r1], self._y[r0:r1], self.n_train / (r1 - r0), 1, self.lr0net)
            self.elbos_mfvi.append(-losses.item())

    # ---- reference :341-354
    def test(self):
        acc, nll = self._trainer.test(self._xt, self._yt, int(self.data_minibatch))
        self.times_mfvi.append(self.times_mfvi[-1] + time.time() - self.t_start)
        self.nlls_mfvi.append(nll)
        self.accs_mfvi.append(acc)
        if not self.quiet:
            print(f"predictive accuracy: {(100*self.accs_mfvi[-1]):.2f}%")

    def _mean_logits(self, x):
        from psvi import _native
        from psvi.models.neural_net import MeanFieldLeNet, MeanFieldMLP
        if self.noise_source is not None and isinstance(self.net, MeanFieldMLP) and not isinstance(self.net, MeanFieldLeNet):
            return self.net(x, noise=_native.make_noise(self.noise_source.take(1, self.device))).mean(0)
        return self.net(x).mean(0)

    # ---- reference :356-385: forgetting events / never-learnt flags from the per-row accuracy after every epoch
    def after_epoch(self):
        if not self.forgetting_score_flag:
            return
        with torch.no_grad():
            for r0, r1 in self._batches():
                curr_acc = self._mean_logits(self._x[r0:r1]).argmax(-1).eq(self._y[r0:r1]).float()
                self.forgetting_events[r0:r1] += (self.last_acc[r0:r1] > curr_acc).float()
                self.last_acc[r0:r1] = curr_acc
                self.never_learnt_events[r0:r1] = torch.min(self.never_learnt_events[r0:r1], 1.0 - curr_acc)

    # ---- reference :388-408
    def run(self):
        self.before_train()
        if self.load_from_saved and self.load():
            return
        for i in range(self.total_iterations):
            self.train_an_epoch()
            self.after_epoch()
            if i % self.log_every == 0 or i == self.total_iterations - 1:
                self.test()
        if self.forgetting_score_flag:
            self.forgetting_events = torch.max(self.total_iterations * self.never_learnt_events, self.forgetting_events)
        if self.data_path is not None:
            self.save()

    def _get_net_fname(self):
        return os.path.join(self.data_path, f"net_state_dict_{self.dnm}_{self.architecture}_{self.num_epochs}_{self.seed}.pt")

    def _get_forgetting_fname(self):
        return os.path.join(self.data_path, f"forgetting_{self.dnm}_{self.architecture}_{self.num_epochs}_{self.seed}.pt")

    def save(self):
        torch.save(self.net.state_dict(), self._get_net_fname())
        torch.save(self.forgetting_events, self._get_forgetting_fname())

    def load(self):
        net_f, forg_f = self._get_net_fname(), self._get_forgetting_fname()
        if not (os.path.exists(net_f) and os.path.exists(forg_f)):
            return False
        self.forgetting_events = torch.load(forg_f, map_location=self.device)
        self.net.load_state_dict(torch.load(net_f, map_location=self.device))
        return True
