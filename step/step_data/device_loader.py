"""Device-resident input pipeline.

The reference ships every batch through DataLoader workers as overlapping windows: 160 MB of `long_history` per
STEP_METR-LA batch, although the whole pre-processed series is only `[34272, 207, 3]` fp32 = 85 MB
(step/step_data/forecasting_dataset.py:62-71, step_runner.py:58-63).  Here the series is uploaded ONCE and each batch is
gathered on the device from the window index: only the B sample indices cross PCIe.  Items are identical to
`ForecastingDataset.__getitem__` stacked over the batch (tests/test_host_logic.py).
"""
from typing import Iterator, Optional, Sequence, Tuple

import torch

from .forecasting_dataset import ForecastingDataset


class DeviceWindowLoader:
    """`for future, history, long_history in loader:` with tensors already on `device`.

    dataset: a ForecastingDataset (file-backed or synthetic); batch_size / shuffle / drop_last as in DataLoader;
    seed: shuffling generator seed (epoch e uses seed + e).  `long_channels`: optional channel subset of the long
    history (STEP only reads channel 0 of it, discrete_graph_learning.py:139) - None keeps all channels.
    """

    def __init__(self, dataset: ForecastingDataset, device, batch_size: int, shuffle: bool = False, drop_last: bool = False,
                 seed: int = 0, long_channels: Optional[Sequence[int]] = None):
        self.device = torch.device(device)
        self.batch_size, self.shuffle, self.drop_last, self.seed = batch_size, shuffle, drop_last, seed
        self.seq_len = dataset.seq_len
        self.data = dataset.data.to(self.device)                                   # [T, N, C], resident
        idx = torch.as_tensor([list(t) for t in dataset.index], dtype=torch.long)  # [M, 3] = (history start, split, future end)
        self.hist_len = int(idx[0, 1] - idx[0, 0])
        self.fut_len = int(idx[0, 2] - idx[0, 1])
        if not (bool((idx[:, 1] - idx[:, 0] == self.hist_len).all()) and bool((idx[:, 2] - idx[:, 1] == self.fut_len).all())):
            raise ValueError("DeviceWindowLoader: the window index has varying history / future lengths")
        self.split = idx[:, 1].contiguous().to(self.device)                        # position of the first future step
        self.long_channels = None if long_channels is None else torch.as_tensor(list(long_channels), device=self.device)
        self._hist_off = torch.arange(-self.hist_len, 0, device=self.device)
        self._fut_off = torch.arange(0, self.fut_len, device=self.device)
        self._long_off = torch.arange(-self.seq_len, 0, device=self.device)
        self.epoch = 0

    def __len__(self) -> int:
        m = self.split.numel()
        return m // self.batch_size if self.drop_last else (m + self.batch_size - 1) // self.batch_size

    def gather(self, sample_ids: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """sample_ids: [B] long (any device) -> (future [B,12,N,C], history [B,12,N,C], long_history [B,seq_len,N,C'])."""
        split = self.split[sample_ids.to(self.device)]                             # [B]
        history = self.data[split[:, None] + self._hist_off]
        future = self.data[split[:, None] + self._fut_off]
        rows = split[:, None] + self._long_off                                     # [B, seq_len]; negative = before the series
        short = split < self.seq_len                                               # the reference serves an all-zero window
        long_history = self.data[rows.clamp_(min=0)]
        if self.long_channels is not None:
            long_history = long_history.index_select(-1, self.long_channels)
        long_history = long_history.masked_fill(short.view(-1, 1, 1, 1), 0.0)       # no host sync: applied unconditionally
        return future, history, long_history

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]]:
        m = self.split.numel()
        if self.shuffle:
            g = torch.Generator().manual_seed(self.seed + self.epoch)
            order = torch.randperm(m, generator=g)
        else:
            order = torch.arange(m)
        self.epoch += 1
        for s in range(0, m, self.batch_size):
            ids = order[s:s + self.batch_size]
            if self.drop_last and ids.numel() < self.batch_size:
                break
            yield self.gather(ids)
