"""Dataset with the on-disk format of the reference's data/bar_dataset.py:9-25: one ``.npz`` per item holding SEVERAL
bars under the keys ``note [n,1,96,60]``, ``pre_note [n,1,96,60]``, ``pre_phrase [n,1,384,60]``, ``position [n]``;
the collate function concatenates items along axis 0 (agent/barGen.py:134-141).  ``SyntheticBars`` produces items of
the same structure without touching the disk (benchmarks, smoke tests)."""
import os

import numpy as np
from torch.utils.data import Dataset


class NoteDataset(Dataset):
    def __init__(self, root_dir, config):
        self.root_dir, self.config = root_dir, config
        self.file_list = sorted(os.listdir(os.path.join(root_dir, config.data_path)))
        self.num_iterations = (len(self.file_list) + config.batch_size - 1) // config.batch_size

    def __len__(self):
        return len(self.file_list)

    def __getitem__(self, idx):
        with np.load(os.path.join(self.root_dir, self.config.data_path, self.file_list[idx])) as d:
            return {k: d[k] for k in ("note", "pre_note", "pre_phrase", "position")}


class SyntheticBars(Dataset):
    """`n_items` items of `bars_per_item` random 5 %-density bars each (same dict layout as NoteDataset)."""

    def __init__(self, n_items, bars_per_item=4, batch_size=8, seed=0, density=0.05):
        self.n_items, self.bars, self.seed, self.density = n_items, bars_per_item, seed, density
        self.num_iterations = (n_items + batch_size - 1) // batch_size

    def __len__(self):
        return self.n_items

    def __getitem__(self, idx):
        r = np.random.RandomState(self.seed * 100003 + idx)
        n = self.bars
        return {"note": (r.rand(n, 1, 96, 60) < self.density).astype(np.float32),
                "pre_note": (r.rand(n, 1, 96, 60) < self.density).astype(np.float32),
                "pre_phrase": (r.rand(n, 1, 384, 60) < self.density).astype(np.float32),
                "position": r.randint(0, 332, size=(n,)).astype(np.int64)}
