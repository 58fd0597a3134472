import inspect

import torch


class MessagePassing(torch.nn.Module):
    """Minimal `propagate`: gather `*_j` by edge_index[0] (source), call
    message -> aggregate(index=edge_index[1]) -> update, which is the flow
    PyG's MessagePassing(aggr="add", node_dim=0, flow="source_to_target")
    runs for painn_denoising.py:537-567."""

    def __init__(self, aggr="add", node_dim=0, **kw):
        super().__init__()
        self.aggr = aggr
        self.node_dim = node_dim

    def jittable(self):
        return self

    def propagate(self, edge_index, size=None, **kwargs):
        j, i = edge_index[0], edge_index[1]
        dim_size = None
        msg_args = {}
        for name in inspect.signature(self.message).parameters:
            if name.endswith("_j"):
                src = kwargs[name[:-2]]
                dim_size = src.size(self.node_dim)
                msg_args[name] = src.index_select(self.node_dim, j)
            elif name.endswith("_i"):
                src = kwargs[name[:-2]]
                dim_size = src.size(self.node_dim)
                msg_args[name] = src.index_select(self.node_dim, i)
            else:
                msg_args[name] = kwargs[name]
        out = self.message(**msg_args)
        out = self.aggregate(out, index=i, ptr=None, dim_size=dim_size)
        return self.update(out)


def radius_graph(*a, **k):
    raise NotImplementedError("oracle shim: non-PBC radius_graph is off the hot path")
