"""Data loading for training / evaluation runs (reference: data/__init__.py:41-53,
data/base_dataset.py:17-250, data/celebamaskhq_dataset.py).

`create_dataloader(opt)` returns a torch DataLoader over an image / label-map folder pair with the
reference's item format ({'label': float [1,S,S], 'image': float [3,S,S] in [-1,1], 'path': str});
`DevicePrefetcher` wraps it so that the next batch's host->device copy (from pinned memory, on a side
stream) overlaps the current training step.  Neither is on the throughput path bench.py measures
(synthetic data); they make train.py-style runs on a real CelebA(-HQ) folder possible.
"""
import torch

from .folder_dataset import FolderDataset, find_dataset_using_name  # noqa: F401


def create_dataloader(opt):
    """data/__init__.py:41-53, plus pinned host buffers for asynchronous H2D copies."""
    dataset = find_dataset_using_name(getattr(opt, "dataset_mode", "celebamaskhq"))()
    dataset.initialize(opt)
    print("dataset [%s] of size %d was created" % (type(dataset).__name__, len(dataset)))
    return torch.utils.data.DataLoader(
        dataset, batch_size=opt.batchSize, shuffle=not getattr(opt, "serial_batches", False),
        num_workers=int(getattr(opt, "nThreads", 0)), drop_last=bool(opt.isTrain),
        pin_memory=torch.cuda.is_available())


class DevicePrefetcher:
    """Iterates a DataLoader one batch ahead: while the caller trains on batch i, batch i+1 is copied
    to the GPU on a side stream from pinned memory.  Yields dicts whose tensors live on the device;
    non-tensor entries ('path') pass through."""

    def __init__(self, loader, device=None):
        self.loader = loader
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        self.stream = torch.cuda.Stream(device=self.device)

    def __len__(self):
        return len(self.loader)

    def _upload(self, batch):
        out = {}
        with torch.cuda.stream(self.stream):
            for k, v in batch.items():
                if torch.is_tensor(v):
                    if not v.is_pinned():
                        v = v.pin_memory()
                    out[k] = v.to(self.device, non_blocking=True)
                else:
                    out[k] = v
        return out

    def __iter__(self):
        it = iter(self.loader)
        try:
            nxt = self._upload(next(it))
        except StopIteration:
            return
        while nxt is not None:
            torch.cuda.current_stream(self.device).wait_stream(self.stream)
            cur = nxt
            for v in cur.values():
                if torch.is_tensor(v):
                    v.record_stream(torch.cuda.current_stream(self.device))
            try:
                nxt = self._upload(next(it))
            except StopIteration:
                nxt = None
            yield cur
