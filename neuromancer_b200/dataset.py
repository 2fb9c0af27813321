"""``DictDataset``: the one dataset class every benchmark config uses (reference dataset.py:87-121)."""
import torch
from torch.utils.data import Dataset
from torch.utils.data.dataloader import default_collate


class DictDataset(Dataset):
    def __init__(self, datadict, name="train"):
        super().__init__()
        sizes = {v.shape[0] for v in datadict.values()}
        assert len(sizes) == 1, "Mismatched number of samples in dataset tensors"
        self.datadict, self.length, self.name = datadict, sizes.pop(), name

    def __len__(self):
        return self.length

    def __getitem__(self, i):
        return {k: v[i] for k, v in self.datadict.items()}

    def collate_fn(self, batch):
        out = default_collate(batch)
        out["name"] = self.name
        return out


def move_batch_to_device(batch, device="cpu"):
    return {k: (v.to(device, non_blocking=True) if isinstance(v, torch.Tensor) else v)
            for k, v in batch.items()}
