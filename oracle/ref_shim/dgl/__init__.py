"""Pure-torch stand-in for the few DGL entry points the reference MeshGraphNet path calls.

TEST INFRASTRUCTURE ONLY.  DGL (the reference's third-party gather / SpMM backend) is not
installable in this image; this shim lets the UNMODIFIED reference under /root/reference be
imported so golden vectors can be generated (tests/golden/make_golden.py).  Only the
gather and the segment sum are restated here; Linear / LayerNorm / autograd stay the
reference's own code on real PyTorch.  Call sites mirrored (reference file:line):

  gnn_layers/utils.py:94-148   apply_edges(concat_message_function)   edges.data/src/dst
  gnn_layers/utils.py:337-378  update_all(fn.copy_e, fn.sum|fn.mean)  -> dstdata
  gnn_layers/graph.py:447-479  dgl.heterograph({(s,r,d): ("coo",(src,dst))}, idtype=int32)
  gnn_layers/graph.py:157-159  graph.adj_tensors("csc")   (stable sort by destination)
  test/models/meshgraphnet/test_meshgraphnet.py:41-65     dgl.graph / dgl.batch
"""
from contextlib import contextmanager
from types import SimpleNamespace

import torch

from . import function  # noqa: F401


class _EdgeBatch:
    def __init__(self, g):
        self.data = g.edata
        self.src = {k: v[g._src] for k, v in g.srcdata.items()}
        self.dst = {k: v[g._dst] for k, v in g.dstdata.items()}


class DGLGraph:
    def __init__(self, src, dst, num_src=None, num_dst=None, bipartite=False):
        self._src = torch.as_tensor(src).long()
        self._dst = torch.as_tensor(dst).long()
        if num_src is None:
            n = int(max(self._src.max().item(), self._dst.max().item())) + 1 if self._src.numel() else 0
            num_src = num_dst = n
        self._num_src, self._num_dst = int(num_src), int(num_dst)
        self._bipartite = bipartite
        self.srcdata = {}
        self.dstdata = self.srcdata if not bipartite else {}
        self.edata = {}

    # homogeneous alias
    @property
    def ndata(self):
        return self.srcdata

    @contextmanager
    def local_scope(self):
        saved = (dict(self.srcdata), dict(self.dstdata), dict(self.edata))
        try:
            yield
        finally:
            self.srcdata.clear()
            self.srcdata.update(saved[0])
            if self._bipartite:
                self.dstdata.clear()
                self.dstdata.update(saved[1])
            self.edata.clear()
            self.edata.update(saved[2])

    def apply_edges(self, udf):
        self.edata.update(udf(_EdgeBatch(self)))

    def update_all(self, msg, red):
        kind, ein, _ = msg
        assert kind == "copy_e"
        rkind, _, out = red
        x = self.edata[ein]
        h = torch.zeros((self._num_dst,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        h = h.index_add(0, self._dst.to(x.device), x)
        if rkind == "mean":
            deg = torch.bincount(self._dst.to(x.device), minlength=self._num_dst).clamp(min=1)
            h = h / deg.to(h.dtype).view(-1, *([1] * (x.dim() - 1)))
        elif rkind != "sum":
            raise RuntimeError(rkind)
        self.dstdata[out] = h

    def edges(self):
        return self._src, self._dst

    def num_nodes(self):
        return self._num_src

    def num_src_nodes(self):
        return self._num_src

    def num_dst_nodes(self):
        return self._num_dst

    def num_edges(self):
        return int(self._src.numel())

    def to(self, device):
        self._src = self._src.to(device)
        self._dst = self._dst.to(device)
        return self

    def cpu(self):
        return self.to("cpu")

    def adj_tensors(self, fmt):
        assert fmt == "csc"
        perm = torch.argsort(self._dst, stable=True)
        deg = torch.bincount(self._dst, minlength=self._num_dst)
        offsets = torch.zeros(self._num_dst + 1, dtype=torch.int64)
        offsets[1:] = torch.cumsum(deg, 0)
        return offsets, self._src[perm], perm


def graph(data, num_nodes=None, **kw):
    src, dst = data
    g = DGLGraph(src, dst)
    if num_nodes is not None:
        g._num_src = g._num_dst = int(num_nodes)
    return g


def batch(graphs):
    srcs, dsts, off = [], [], 0
    for g in graphs:
        srcs.append(g._src + off)
        dsts.append(g._dst + off)
        off += g.num_nodes()
    out = DGLGraph(torch.cat(srcs), torch.cat(dsts))
    out._num_src = out._num_dst = off
    return out


def heterograph(data, idtype=None, **kw):
    ((_s, _r, _d), (fmt, (src, dst))), = data.items()
    assert fmt == "coo"
    n_src = int(src.max().item()) + 1 if src.numel() else 0
    n_dst = int(dst.max().item()) + 1 if dst.numel() else 0
    return DGLGraph(src, dst, num_src=n_src, num_dst=n_dst, bipartite=True)


def seed(s):
    torch.manual_seed(s)


def rand_graph(num_nodes, num_edges):
    src = torch.randint(0, num_nodes, (num_edges,))
    dst = torch.randint(0, num_nodes, (num_edges,))
    return graph((src, dst), num_nodes=num_nodes)


backend = SimpleNamespace(name="pytorch")


def to_bidirected(g, copy_ndata=False):
    """DGL semantics (dgl.to_bidirected, v2.4: add_reverse_edges + to_simple): the union of every edge and its
    reverse with duplicates removed; the result is coalesced, i.e. edges ordered by (src, dst)."""
    n = g.num_nodes()
    s = torch.cat([g._src, g._dst])
    d = torch.cat([g._dst, g._src])
    key = torch.unique(s * n + d)
    out = DGLGraph(torch.div(key, n, rounding_mode="floor"), key % n)
    out._num_src = out._num_dst = n
    return out
