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


class DevicePrefetcher:
    """Iterate over host batches (dicts of pinned CPU tensors) and hand them out as device batches, with the
    host->device copy of batch i+1 running on a side stream while batch i is being computed.

    The trainer's ``for batch in train_data: move_batch_to_device(...)`` loop (reference trainer.py:240-243)
    serialises copy and compute; on B200 the copy of one C2 batch (16.8 MB) is 0.31 ms next to a 1.3 ms step.
    The prefetcher keeps the order of batches and makes the compute stream wait for the copy it consumes, so
    results are unchanged."""

    # One copy stream per device, shared by every prefetcher: the caching allocator keeps a block pool PER STREAM, so a
    # fresh stream per epoch would start from an empty pool and pay cudaMalloc (measured: 2-4 ms/step instead of 1.05
    # for the first ~40 steps of every new prefetcher).
    _copy_streams = {}

    def __init__(self, batches, device="cuda", depth=2):
        self.batches, self.device = batches, torch.device(device)
        self.stream = None
        if self.device.type == "cuda":
            key = self.device.index if self.device.index is not None else torch.cuda.current_device()
            if key not in DevicePrefetcher._copy_streams:
                DevicePrefetcher._copy_streams[key] = torch.cuda.Stream(self.device)
            self.stream = DevicePrefetcher._copy_streams[key]
        self.depth = max(int(depth), 1)

    def _issue(self, batch):
        if self.stream is None:
            return move_batch_to_device(batch, self.device), None
        with torch.cuda.stream(self.stream):
            dev = move_batch_to_device(batch, self.device)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return dev, ev

    def __iter__(self):
        it = iter(self.batches)
        queue = []
        try:
            for _ in range(self.depth):
                queue.append(self._issue(next(it)))
        except StopIteration:
            pass
        while queue:
            dev, ev = queue.pop(0)
            try:
                queue.append(self._issue(next(it)))          # next copy goes out before this batch is consumed
            except StopIteration:
                pass
            if ev is not None:
                torch.cuda.current_stream(self.device).wait_event(ev)
                for v in dev.values():                        # the caching allocator must not recycle the buffers early
                    if isinstance(v, torch.Tensor):
                        v.record_stream(torch.cuda.current_stream(self.device))
            yield dev
