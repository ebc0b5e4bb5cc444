"""TEST INFRASTRUCTURE ONLY -- minimal restatement of the DGL 0.4.1 graph API.

The reference (JiapengWu/TeMP) executes its message passing through the third-party
package ``dgl-cuda10.1==0.4.1`` (reference README.md:15), which is not vendored under
/root/reference and is not installable here (no network).  This module restates, on
CPU with plain torch ops, exactly the slice of the published DGL 0.4.x semantics that the
reference's own call sites use, so that the UNMODIFIED reference python files can be
imported and run in this container to produce golden vectors
(tests/golden/make_golden.py).  It is never imported by the product (temp_b200/).

Call sites covered (reference file:line):
  DGLGraph(), add_nodes, add_edges            utils/dataset.py:171-216
  ndata / edata dict-like frames (+ .update)  utils/dataset.py:220-224
  in_degrees(range(n))                        utils/utils.py:74-79
  local_var, apply_edges                      utils/utils.py:23-28
  update_all(msg_udf, fn.sum, apply_udf)      models/RGCN.py:100-104
  edges(), nodes(), number_of_nodes()         models/DynamicRGCN.py:82, 114; utils/CorrptTriples.py:21
  edge_subgraph(idx, preserve_nodes=True)     models/DynamicRGCN.py:84
  dgl.batch(list)                             models/DynamicRGCN.py:72, 92

Semantics restated from DGL 0.4.x:
  * edge ids are assigned in insertion order; edges() returns (src, dst) in edge-id order;
  * update_all with a UDF message function calls it ONCE with all edges (edge-id order), the
    builtin ``fn.sum`` reducer writes, for EVERY node, the sum of its incoming messages
    (zero-filled for nodes without in-edges) into ndata[out]; the apply UDF then sees all nodes;
  * local_var() shares structure and feature tensors but feature (re)assignment on the
    returned graph does not leak to the original;
  * batch() forms the block-diagonal union: node/edge features concatenated along dim 0 in list
    order, node ids offset by the running node count;
  * edge_subgraph(idx, preserve_nodes=True) keeps all nodes and takes the listed edges in the
    listed order; features are NOT copied (0.4.x requires copy_from_parent()).
"""
import numpy as np
import torch

from . import function  # noqa: F401

__version__ = "0.4.1-stub"


def _as_long(x):
    if isinstance(x, torch.Tensor):
        return x.long().reshape(-1)
    if isinstance(x, range):
        x = list(x)
    return torch.as_tensor(np.asarray(x), dtype=torch.long).reshape(-1)


class _Lazy(object):
    """edges.src[...] / edges.dst[...]: gather node features on demand."""

    def __init__(self, frame, index):
        self._frame, self._index = frame, index

    def __getitem__(self, key):
        return self._frame[key][self._index]

    def __contains__(self, key):
        return key in self._frame


class EdgeBatch(object):
    def __init__(self, g):
        self.src = _Lazy(g.ndata, g._src)
        self.dst = _Lazy(g.ndata, g._dst)
        self.data = g.edata


class NodeBatch(object):
    def __init__(self, g):
        self.data = g.ndata


class DGLGraph(object):
    def __init__(self):
        self._n = 0
        self._src = torch.zeros(0, dtype=torch.long)
        self._dst = torch.zeros(0, dtype=torch.long)
        self.ndata = {}
        self.edata = {}

    # ---- structure -------------------------------------------------------------------
    def add_nodes(self, num):
        self._n += int(num)

    def add_edges(self, u, v):
        u, v = _as_long(u), _as_long(v)
        assert u.shape == v.shape
        self._src = torch.cat([self._src, u])
        self._dst = torch.cat([self._dst, v])

    def number_of_nodes(self):
        return self._n

    def number_of_edges(self):
        return int(self._src.shape[0])

    def nodes(self):
        return torch.arange(self._n, dtype=torch.long)

    def edges(self):
        return self._src, self._dst

    def in_degrees(self, v=None):
        deg = torch.bincount(self._dst, minlength=self._n)
        if v is None:
            return deg
        return deg[_as_long(v)]

    # ---- views ---------------------------------------------------------------------------
    def local_var(self):
        g = DGLGraph()
        g._n, g._src, g._dst = self._n, self._src, self._dst
        g.ndata = dict(self.ndata)
        g.edata = dict(self.edata)
        return g

    def edge_subgraph(self, edges, preserve_nodes=False):
        assert preserve_nodes, "the reference only uses preserve_nodes=True"
        idx = _as_long(edges)
        g = DGLGraph()
        g._n = self._n
        g._src, g._dst = self._src[idx], self._dst[idx]
        return g

    # ---- computation -------------------------------------------------------------------
    def apply_edges(self, func):
        self.edata.update(func(EdgeBatch(self)))

    def update_all(self, message_func, reduce_func, apply_node_func=None):
        msgs = message_func(EdgeBatch(self))
        assert isinstance(reduce_func, function._SumReducer)
        msg = msgs[reduce_func.msg]
        out = msg.new_zeros((self._n,) + tuple(msg.shape[1:]))
        out = out.index_add(0, self._dst.to(msg.device), msg)
        self.ndata[reduce_func.out] = out
        if apply_node_func is not None:
            self.ndata.update(apply_node_func(NodeBatch(self)))


def batch(graph_list):
    g = DGLGraph()
    srcs, dsts, off = [], [], 0
    for x in graph_list:
        srcs.append(x._src + off)
        dsts.append(x._dst + off)
        off += x._n
    g._n = off
    if graph_list:
        g._src, g._dst = torch.cat(srcs), torch.cat(dsts)
        for k in graph_list[0].ndata:
            if all(k in x.ndata for x in graph_list):
                g.ndata[k] = torch.cat([x.ndata[k] for x in graph_list], dim=0)
        for k in graph_list[0].edata:
            if all(k in x.edata for x in graph_list):
                g.edata[k] = torch.cat([x.edata[k] for x in graph_list], dim=0)
    return g
