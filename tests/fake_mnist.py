"""MNIST-shaped synthetic dataset (there is no network for the real one): uint8 [N, 28, 28] images with the reference's
MNIST normalisation (psvi/experiments/experiments_utils.py:42-46) and class-dependent blobs so the classes are learnable.
Used by oracle/make_goldens.py (to drive the reference) and by the GPU tests / bench (to drive the CUDA path)."""
import numpy as np
import torch


class FakeMNIST(torch.utils.data.Dataset):
    def __init__(self, n, seed):
        import torchvision.transforms as T
        g = np.random.default_rng(seed)
        self.targets = torch.tensor(g.integers(0, 10, n))
        img = g.uniform(0, 60, (n, 28, 28))
        for i, c in enumerate(self.targets.numpy()):
            cy, cx = 6 + 2 * (c // 3), 6 + 5 * (c % 3)
            img[i, cy:cy + 8, cx:cx + 8] += 150
        self.data = torch.tensor(np.clip(img, 0, 255).astype(np.uint8))
        self.transform = T.Compose([T.ToTensor(), T.Normalize((0.1307,), (0.3081,))])

    def __len__(self):
        return len(self.targets)

    def __getitem__(self, i):
        from PIL import Image
        return self.transform(Image.fromarray(self.data[i].numpy(), mode="L")), int(self.targets[i])
