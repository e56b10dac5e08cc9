"""Synthetic stand-in for the reference DataModule (reference data/datamodule.py:20-60, :180-188).

Same constructor keywords as the reference's dataset section (`name, img_size, img_channels, batch_size,
train_val_split`), same per-rank batch rule (`batch_size / world_size`, datamodule.py:33) and the same batch format
`(data[B,C,S,S] fp32, labels[B])`; the images are seeded uniform noise instead of a torchvision dataset (there
is no dataset on the box).  The reference normalises images to [-1,1] with transforms.Normalize(0.5, 0.5) BEFORE
the model normalises again (ddpm.py:945 auto_normalize) — kept, since the loader is meant to be a drop-in.
Batches are staged in pinned host memory and copied with non_blocking H2D copies."""
from __future__ import annotations

import torch


class DataModule:
    def __init__(self, name: str = "synthetic", img_size: int = 32, img_channels: int = 3, batch_size: int = 32,
                 train_val_split: float = 0.8, num_workers: int = 0, pin_memory: bool = True,
                 num_images: int = 2048, seed: int = 10, world_size: int = 1, rank: int = 0, device=None, **_):
        self.name, self.img_size, self.img_channels = str(name), img_size, img_channels
        self.batch_size = max(1, int(batch_size / (world_size if world_size > 1 else 1)))
        self.world_size, self.rank = world_size, rank
        self.device = device
        n_train = max(self.batch_size * world_size, int(num_images * train_val_split))
        n_val = max(self.batch_size * world_size, num_images - n_train)
        self.sizes = {"train": n_train, "val": n_val}
        self.seed = seed
        self.pin = bool(pin_memory) and torch.cuda.is_available()
        self._data = {}

    def setup(self, stage=None):
        for i, split in enumerate(("train", "val")):
            g = torch.Generator().manual_seed(self.seed + 1000 * i)
            x = torch.rand(self.sizes[split], self.img_channels, self.img_size, self.img_size, generator=g) * 2 - 1
            self._data[split] = x.pin_memory() if self.pin else x

    def _loader(self, split: str, shuffle: bool, epoch: int = 0):
        """Yields this rank's batches of one epoch (a DistributedSampler-style strided shard of a seeded permutation)."""
        if not self._data:
            self.setup()
        x = self._data[split]
        n = x.shape[0]
        if shuffle:
            g = torch.Generator().manual_seed(self.seed + 7919 * epoch)
            order = torch.randperm(n, generator=g)
        else:
            order = torch.arange(n)
        order = order[self.rank::self.world_size]
        for i in range(0, order.numel() - self.batch_size + 1, self.batch_size):
            idx = order[i:i + self.batch_size]
            batch = x[idx]
            if self.pin:
                batch = batch.pin_memory()
            if self.device is not None:
                batch = batch.to(self.device, non_blocking=True)
            labels = torch.zeros(batch.shape[0], dtype=torch.long, device=batch.device)
            yield batch, labels

    def train_dataloader(self, epoch: int = 0):
        return self._loader("train", True, epoch)

    def val_dataloader(self):
        return self._loader("val", False)

    def steps_per_epoch(self) -> int:
        return (self.sizes["train"] // self.world_size) // self.batch_size
