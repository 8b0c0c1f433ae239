"""Forecasting dataset with the reference's item contract (step/step_data/forecasting_dataset.py:8-80):
item -> (future [12,N,C], history [12,N,C], long_history [seq_len,N,C]) from `data_in12_out12.pkl` + `index_in12_out12.pkl`.
When the files do not exist (no datasets are shipped and there is no network) `synthetic=True` serves seeded N(0,1) windows
of the same shapes, which is what bench.py measures with."""
import os
import pickle

import torch
from torch.utils.data import Dataset


class ForecastingDataset(Dataset):
    def __init__(self, data_file_path: str = None, index_file_path: str = None, mode: str = "train", seq_len: int = 2016,
                 synthetic: bool = False, num_nodes: int = None, length: int = 256, seed: int = 0):
        assert mode in ["train", "valid", "test"], "error mode"
        self.seq_len = seq_len
        self.synthetic = synthetic
        if synthetic:
            assert num_nodes is not None
            g = torch.Generator().manual_seed(seed)
            self.data = torch.randn(seq_len + 24 + length, num_nodes, 3, generator=g)
            self.index = [(seq_len + i, seq_len + i + 12, seq_len + i + 24) for i in range(length)]
        else:
            for p in (data_file_path, index_file_path):
                if not os.path.isfile(p):
                    raise FileNotFoundError("BasicTS can not find file {0}".format(p))
            with open(data_file_path, "rb") as f:
                self.data = torch.from_numpy(pickle.load(f)["processed_data"]).float()
            with open(index_file_path, "rb") as f:
                self.index = pickle.load(f)[mode]
        self.mask = torch.zeros(self.seq_len, self.data.shape[1], self.data.shape[2])

    def __getitem__(self, index: int) -> tuple:
        idx = list(self.index[index])
        history = self.data[idx[0]:idx[1]]
        future = self.data[idx[1]:idx[2]]
        long_history = self.mask if idx[1] - self.seq_len < 0 else self.data[idx[1] - self.seq_len:idx[1]]
        return future, history, long_history

    def __len__(self):
        return len(self.index)
