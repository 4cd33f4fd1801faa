"""`Data` graph object handed from Text2GraphTransformer to the model.

The reference emits a `torch_geometric.data.Data` (textgcn/lib/text2graph.py:192-193) with the
fields x, edge_index, edge_attr, y, train_mask/val_mask/test_mask, n_vocab.  When
torch_geometric is importable that class is used unchanged; otherwise this module provides a
minimal stand-in with the same attribute access, `.to(device)`, `.keys`, `num_nodes` and
pickling behaviour, which is all the reference's scripts rely on (flat_amazon.py:80-117).
"""
from __future__ import annotations

import copy
from typing import Any, Dict, Iterator, List

import torch

try:  # pragma: no cover - not installed in the build image
    from torch_geometric.data import Data as _PygData
    HAVE_PYG = True
except Exception:  # ImportError, or a broken install
    _PygData = None
    HAVE_PYG = False


class _Data:
    """Duck-typed stand-in for torch_geometric.data.Data (attribute bag of tensors)."""

    def __init__(self, x=None, edge_index=None, edge_attr=None, y=None, **kwargs):
        self.x = x
        self.edge_index = edge_index
        self.edge_attr = edge_attr
        self.y = y
        for k, v in kwargs.items():
            setattr(self, k, v)

    # -- torch_geometric.data.Data surface used by the reference scripts --
    @property
    def keys(self) -> List[str]:
        return [k for k, v in self.__dict__.items() if v is not None and not k.startswith("_")]

    def __contains__(self, key: str) -> bool:
        return key in self.keys

    def __getitem__(self, key: str) -> Any:
        return getattr(self, key)

    def __setitem__(self, key: str, value: Any) -> None:
        setattr(self, key, value)

    def __iter__(self) -> Iterator:
        for k in sorted(self.keys):
            yield k, getattr(self, k)

    @property
    def num_nodes(self) -> int:
        if self.x is not None:
            return int(self.x.shape[0])
        if self.y is not None:
            return int(self.y.shape[0])
        return int(self.edge_index.max()) + 1

    @property
    def num_edges(self) -> int:
        return int(self.edge_index.shape[1])

    def apply(self, fn) -> "_Data":
        for k in self.keys:
            v = getattr(self, k)
            if torch.is_tensor(v):
                setattr(self, k, fn(v))
        return self

    def to(self, device, non_blocking: bool = False) -> "_Data":
        """In place, returns self (PyG 1.6.3 `Data.to` semantics; flat_amazon.py:86)."""
        return self.apply(lambda t: t.to(device, non_blocking=non_blocking))

    def cpu(self) -> "_Data":
        return self.to("cpu")

    def cuda(self, device=None) -> "_Data":
        return self.to("cuda" if device is None else device)

    def clone(self) -> "_Data":
        out = _Data()
        for k, v in self.__dict__.items():
            out.__dict__[k] = v.clone() if torch.is_tensor(v) else copy.deepcopy(v)
        return out

    def __getstate__(self) -> Dict[str, Any]:
        # device-side caches (CSR handles) never travel in a pickle (text2graph.py:195-202)
        return {k: v for k, v in self.__dict__.items() if not k.startswith("_")}

    def __setstate__(self, state: Dict[str, Any]) -> None:
        self.__dict__.update(state)

    def __repr__(self) -> str:
        parts = []
        for k in self.keys:
            v = getattr(self, k)
            parts.append(f"{k}={list(v.shape)}" if torch.is_tensor(v) else f"{k}={v!r}")
        return f"Data({', '.join(parts)})"


Data = _PygData if HAVE_PYG else _Data
