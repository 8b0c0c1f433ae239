"""Pre-training dataset (reference step/step_data/pretraining_dataset.py = basicts TimeSeriesForecastingDataset,
basicts/data/dataset.py:9-73): item -> (future [12,N,C], history [L,N,C]) where L = DATASET_INPUT_LEN is the long window
(2016 or 4032 steps) and the index pickle is ``index_in{L}_out12.pkl``.  ``synthetic=True`` serves seeded N(0,1) windows of
the same shapes when no dataset files exist (none are shipped, there is no network)."""
import os
import pickle

import torch
from torch.utils.data import Dataset


class PretrainingDataset(Dataset):
    def __init__(self, data_file_path: str = None, index_file_path: str = None, mode: str = "train", synthetic: bool = False,
                 num_nodes: int = None, seq_len: int = 2016, length: int = 64, seed: int = 0):
        assert mode in ["train", "valid", "test"], "error mode"
        if synthetic:
            assert num_nodes is not None
            g = torch.Generator().manual_seed(seed)
            self.data = torch.randn(seq_len + 12 + length, num_nodes, 3, generator=g)
            self.index = [(i, i + seq_len, i + seq_len + 12) for i in range(length)]
        else:
            for what, p in (("data", data_file_path), ("index", index_file_path)):
                if p is None or not os.path.isfile(p):
                    raise FileNotFoundError("BasicTS can not find {0} file {1}".format(what, p))
            with open(data_file_path, "rb") as f:
                self.data = torch.from_numpy(pickle.load(f)["processed_data"]).float()
            with open(index_file_path, "rb") as f:
                self.index = pickle.load(f)[mode]

    def __getitem__(self, index: int) -> tuple:
        idx = list(self.index[index])
        if isinstance(idx[0], int):
            return self.data[idx[1]:idx[2]], self.data[idx[0]:idx[1]]
        history_index = list(idx[0])                     # discontinuous / custom index (dataset.py:55-61)
        assert idx[1] not in history_index, "current time t should not included in the idx[0]"
        history_index.append(idx[1])
        return self.data[idx[1], idx[2]], self.data[history_index]

    def __len__(self):
        return len(self.index)
