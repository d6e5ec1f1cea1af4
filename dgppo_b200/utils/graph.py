"""GraphsTuple: the 10-field graph record of the reference
(dgppo/utils/graph.py:47-189), holding torch tensors.

Field order, names and the helper API (`type_nodes`, `type_states`,
`_replace`, `without_edge`, `is_single`, `n_graphs`, `batch_shape`) follow the
reference.  Arrays may carry any number of leading batch axes (the reference
adds them with jax.vmap; here the kernels are batched natively).
"""
from __future__ import annotations

from typing import Any, NamedTuple, Optional

import torch


class GraphsTuple(tuple):
    _FIELDS = ("n_node", "n_edge", "nodes", "edges", "states", "receivers", "senders",
               "node_type", "env_states", "connectivity")

    def __new__(cls, n_node, n_edge, nodes, edges, states, receivers, senders, node_type,
                env_states, connectivity=None):
        tup = (n_node, n_edge, nodes, edges, states, receivers, senders, node_type, env_states, connectivity)
        self = tuple.__new__(cls, tup)
        for k, v in zip(cls._FIELDS, tup):
            object.__setattr__(self, k, v)
        return self

    def __getnewargs__(self):
        return tuple(self)

    @property
    def is_single(self) -> bool:
        return self.n_node.ndim == 0

    @property
    def n_graphs(self) -> int:
        if self.n_node.ndim == 0:
            return 1
        return int(self.n_node.numel())

    @property
    def batch_shape(self):
        return tuple(self.n_node.shape)

    def _type_rows(self, arr: torch.Tensor, type_idx: int, n_type: int) -> torch.Tensor:
        # utils/graph.py:115-141: cumsum + scatter-add == gather of the rows
        # whose node_type matches, in order.  Every graph of a batch shares
        # the same node_type layout (static per env), so one mask serves all.
        nt = self.node_type.reshape(-1, self.node_type.shape[-1])[0]
        idx = torch.nonzero(nt == type_idx, as_tuple=False).flatten()
        assert idx.numel() == n_type, f"expected {n_type} nodes of type {type_idx}, found {idx.numel()}"
        lo, hi = int(idx[0]), int(idx[-1]) + 1
        if hi - lo == n_type:                       # contiguous: a view, no copy
            return arr[..., lo:hi, :]
        return arr.index_select(-2, idx)

    def type_nodes(self, type_idx: int, n_type: int) -> torch.Tensor:
        return self._type_rows(self.nodes, type_idx, n_type)

    def type_states(self, type_idx: int, n_type: int) -> torch.Tensor:
        return self._type_rows(self.states, type_idx, n_type)

    def _replace(self, **kw) -> "GraphsTuple":
        vals = {k: getattr(self, k) for k in self._FIELDS}
        for k, v in kw.items():
            if k not in vals:
                raise ValueError(f"unknown GraphsTuple field {k}")
            vals[k] = v
        return GraphsTuple(**vals)

    def without_edge(self) -> "GraphsTuple":
        return self._replace(edges=None)

    def map_arrays(self, fn) -> "GraphsTuple":
        """Apply ``fn`` to every tensor field (tree_map over the array leaves)."""
        vals = {}
        for k in self._FIELDS:
            v = getattr(self, k)
            vals[k] = fn(v) if isinstance(v, torch.Tensor) else v
        return GraphsTuple(**vals)

    def __str__(self) -> str:
        return "n_node={}, n_edge={}, \n{}\n---------\n{}\n-------\n{}\n  |  \n{}".format(
            self.n_node, self.n_edge, self.nodes, self.edges, self.senders, self.receivers)
