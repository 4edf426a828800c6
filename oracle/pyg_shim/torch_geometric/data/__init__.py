"""Oracle stand-in for `torch_geometric.data.{Data,Batch}` (test infrastructure only).

Restates upstream PyG >= 2.4 `data/data.py` + `data/collate.py` as far as the reference uses
them (etpgt/train/dataloader.py:193-202, scripts/pipeline/run_full_pipeline.py:152-174,
tests/conftest.py:36-49, etpgt/serving/recommender.py:115-117):
  * `x` is concatenated along dim 0, `edge_index` along dim 1 after adding the cumulative
    node offset, every other tensor attribute is stacked when 0-dim and concatenated when
    it has >= 1 dim;
  * `batch` holds the graph id of every node and `ptr` the node offsets.
"""

from __future__ import annotations

import torch


class Data:
    def __init__(self, x=None, edge_index=None, **kwargs):
        self._keys: list[str] = []
        if x is not None:
            self.x = x
        if edge_index is not None:
            self.edge_index = edge_index
        for name, value in kwargs.items():
            setattr(self, name, value)

    def __setattr__(self, name, value):
        if not name.startswith("_") and name not in self.__dict__.get("_keys", []):
            self.__dict__.setdefault("_keys", []).append(name)
        self.__dict__[name] = value

    def keys(self):
        return list(self._keys)

    @property
    def num_nodes(self):
        if "num_nodes" in self.__dict__:
            return self.__dict__["num_nodes"]
        if "x" in self.__dict__ and self.__dict__["x"] is not None:
            return int(self.__dict__["x"].size(0))
        ei = self.__dict__.get("edge_index")
        return int(ei.max()) + 1 if ei is not None and ei.numel() else 0

    @num_nodes.setter
    def num_nodes(self, value):
        if "num_nodes" not in self._keys:
            self._keys.append("num_nodes")
        self.__dict__["num_nodes"] = value

    @property
    def num_edges(self):
        ei = self.__dict__.get("edge_index")
        return int(ei.size(1)) if ei is not None else 0

    def to(self, device):
        for name in self._keys:
            value = self.__dict__[name]
            if torch.is_tensor(value):
                self.__dict__[name] = value.to(device)
        return self


class Batch(Data):
    @property
    def num_graphs(self):
        return int(self.__dict__["ptr"].numel()) - 1

    @classmethod
    def from_data_list(cls, data_list):
        out = cls()
        sizes = [d.num_nodes for d in data_list]
        ptr = torch.zeros(len(data_list) + 1, dtype=torch.long)
        ptr[1:] = torch.tensor(sizes, dtype=torch.long).cumsum(0)
        names: list[str] = []
        for d in data_list:
            for name in d.keys():
                if name not in names and name != "num_nodes":
                    names.append(name)
        for name in names:
            values = [getattr(d, name) for d in data_list]
            if not all(torch.is_tensor(v) for v in values):
                setattr(out, name, values)
            elif name == "edge_index":
                shifted = [v + int(ptr[g]) for g, v in enumerate(values)]
                setattr(out, name, torch.cat(shifted, dim=1))
            elif values[0].dim() == 0:
                setattr(out, name, torch.stack(values))
            else:
                setattr(out, name, torch.cat(values, dim=0))
        out.batch = torch.repeat_interleave(
            torch.arange(len(data_list)), torch.tensor(sizes, dtype=torch.long)
        )
        out.ptr = ptr
        return out
