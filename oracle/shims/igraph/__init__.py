"""
TEST INFRASTRUCTURE -- minimal pure-Python stand-in for ``python-igraph==0.8.2`` (pinned at
/root/reference/requirements.txt:4; absent here and not installable offline).

It implements exactly the subset of the igraph API that /root/reference/graph_ltpl/data_objects/GraphBase.py touches
(call sites GB:122-123,163,179,190,255,266,274,338,355,368,407-417,468-469,508-509,538-541,564-565,615-622,633-644,
658-660,672,705-711,744-745,772-775,818-821,838,883,917) so that the UNMODIFIED reference Python can be executed in this
container to generate golden vectors (oracle/gen_golden.py).  Never imported by the product.

Semantics restated from the igraph C core (from memory -> "parity unpinned" for this file):
  * ``get_shortest_paths(weights=...)`` = igraph_get_shortest_paths_dijkstra, mode OUT: binary heap keyed by tentative
    distance, relax with ``alt = dist[u] + w`` (double), first visit sets the parent, later visits only on STRICT
    ``alt < dist[v]``; out-edges visited in incidence order (sorted by target vertex id, then edge id).
  * ``induced_subgraph`` keeps the relative order of vertices and edges.
  * edge ids: this shim keeps ids stable after ``delete_edges`` (tombstones) -- the reference never caches an edge id
    across a deletion (every id is obtained by get_eid/get_eids immediately before use).
"""

import heapq

OUT = 1
IN = 2
ALL = 3


class InternalError(Exception):
    pass


class Vertex(object):
    __slots__ = ("_g", "index")

    def __init__(self, g, index):
        self._g = g
        self.index = index

    def __getitem__(self, attr):
        return self._g._vattr[attr][self.index]

    def __setitem__(self, attr, value):
        self._g._set_vattr(attr, self.index, value)

    def attributes(self):
        return {k: v[self.index] for k, v in self._g._vattr.items()}


class Edge(object):
    __slots__ = ("_g", "index")

    def __init__(self, g, index):
        self._g = g
        self.index = index

    @property
    def source(self):
        return self._g._src[self.index]

    @property
    def target(self):
        return self._g._dst[self.index]

    @property
    def tuple(self):
        return self._g._src[self.index], self._g._dst[self.index]

    def __getitem__(self, attr):
        return self._g._eattr[attr][self.index]

    def __setitem__(self, attr, value):
        self._g._set_eattr(attr, self.index, value)


class VertexSeq(object):
    def __init__(self, g, indices=None):
        self._g = g
        self._idx = indices  # None == all vertices

    def _indices(self):
        return range(self._g._nv) if self._idx is None else self._idx

    def __len__(self):
        return self._g._nv if self._idx is None else len(self._idx)

    def __iter__(self):
        g = self._g
        for i in self._indices():
            yield Vertex(g, i)

    def __getitem__(self, key):
        if isinstance(key, str):
            col = self._g._vattr[key]
            if self._idx is None:
                return list(col)
            return [col[i] for i in self._idx]
        if self._idx is None:
            if key < 0 or key >= self._g._nv:
                raise IndexError("vertex index out of range")
            return Vertex(self._g, int(key))
        return Vertex(self._g, self._idx[key])

    def find(self, name):
        """vs.find(<name string>) -- ValueError if no such vertex (GB:255-258, 883-885)."""
        idx = self._g._name2idx.get(name)
        if idx is None or (self._idx is not None and idx not in self._idx):
            raise ValueError("no such vertex: %r" % (name,))
        return Vertex(self._g, idx)

    def select(self, **kwds):
        g = self._g
        idx = list(self._indices())
        for key, val in kwds.items():
            attr, _, op = key.rpartition("_")
            if op not in ("eq", "ne", "lt", "gt", "le", "ge", "in", "notin"):
                attr, op = key, "eq"
            col = g._vattr[attr]
            if op == "ge":
                idx = [i for i in idx if col[i] is not None and col[i] >= val]
            elif op == "le":
                idx = [i for i in idx if col[i] is not None and col[i] <= val]
            elif op == "gt":
                idx = [i for i in idx if col[i] is not None and col[i] > val]
            elif op == "lt":
                idx = [i for i in idx if col[i] is not None and col[i] < val]
            elif op == "eq":
                idx = [i for i in idx if col[i] == val]
            elif op == "ne":
                idx = [i for i in idx if col[i] != val]
            elif op == "in":
                sval = set(val)
                idx = [i for i in idx if col[i] in sval]
            elif op == "notin":
                sval = set(val)
                idx = [i for i in idx if col[i] not in sval]
        return VertexSeq(g, idx)

    @property
    def indices(self):
        return list(self._indices())


class EdgeSeq(object):
    def __init__(self, g, indices=None):
        self._g = g
        self._idx = indices  # None == all alive edges

    def _indices(self):
        if self._idx is not None:
            return self._idx
        g = self._g
        if g._n_dead == 0:
            return range(len(g._src))
        return [e for e in range(len(g._src)) if g._alive[e]]

    def __len__(self):
        if self._idx is not None:
            return len(self._idx)
        return len(self._g._src) - self._g._n_dead

    def __iter__(self):
        g = self._g
        for e in self._indices():
            yield Edge(g, e)

    def __call__(self, key):
        if isinstance(key, (list, tuple)):
            return EdgeSeq(self._g, list(key))
        return EdgeSeq(self._g, [int(key)])

    def __getitem__(self, key):
        if isinstance(key, str):
            col = self._g._eattr[key]
            return [col[e] for e in self._indices()]
        if self._idx is None:
            return Edge(self._g, int(key))
        return Edge(self._g, self._idx[key])


class Graph(object):
    def __init__(self, directed=True):
        self._nv = 0
        self._vattr = {"name": []}
        self._name2idx = {}
        self._src = []
        self._dst = []
        self._alive = []
        self._n_dead = 0
        self._eattr = {}
        self._pair2eid = {}
        self._out = None  # lazily built adjacency: v -> [(target, eid), ...] sorted
        self._in = None

    # -- structure -----------------------------------------------------------------------------------------------------
    def to_directed(self, *args, **kwds):
        return None

    def is_directed(self):
        return True

    def vcount(self):
        return self._nv

    def ecount(self):
        return len(self._src) - self._n_dead

    def _set_vattr(self, attr, idx, value):
        col = self._vattr.get(attr)
        if col is None:
            col = [None] * self._nv
            self._vattr[attr] = col
        col[idx] = value
        if attr == "name":
            self._name2idx[value] = idx

    def _set_eattr(self, attr, idx, value):
        col = self._eattr.get(attr)
        if col is None:
            col = [None] * len(self._src)
            self._eattr[attr] = col
        col[idx] = value

    def add_vertex(self, name=None, **kwds):
        idx = self._nv
        self._nv += 1
        for col in self._vattr.values():
            col.append(None)
        if name is not None:
            self._set_vattr("name", idx, name)
        for k, v in kwds.items():
            self._set_vattr(k, idx, v)

    def _vid(self, v):
        if isinstance(v, Vertex):
            return v.index
        if isinstance(v, str):
            idx = self._name2idx.get(v)
            if idx is None:
                raise ValueError("no such vertex: %r" % (v,))
            return idx
        return int(v)

    def add_edge(self, source, target, **kwds):
        s = self._vid(source)
        t = self._vid(target)
        eid = len(self._src)
        self._src.append(s)
        self._dst.append(t)
        self._alive.append(True)
        for col in self._eattr.values():
            col.append(None)
        for k, v in kwds.items():
            self._set_eattr(k, eid, v)
        self._pair2eid[(s, t)] = eid
        self._out = None
        self._in = None

    def get_eid(self, v1, v2, directed=True, error=True):
        s = self._vid(v1)
        t = self._vid(v2)
        eid = self._pair2eid.get((s, t), -1)
        if eid == -1 and error:
            raise InternalError("Cannot get edge id, no such edge")
        return eid

    def get_eids(self, pairs=None, path=None, directed=True, error=True):
        return [self.get_eid(a, b, error=error) for (a, b) in pairs]

    def delete_edges(self, edges):
        if isinstance(edges, (int,)) or not hasattr(edges, "__iter__"):
            edges = [edges]
        for e in set(int(x) for x in edges):
            if self._alive[e]:
                self._alive[e] = False
                self._n_dead += 1
                del self._pair2eid[(self._src[e], self._dst[e])]
        self._out = None
        self._in = None

    def _build_adj(self):
        out = [[] for _ in range(self._nv)]
        inn = [[] for _ in range(self._nv)]
        src, dst, alive = self._src, self._dst, self._alive
        for e in range(len(src)):
            if alive[e]:
                out[src[e]].append((dst[e], e))
                inn[dst[e]].append((src[e], e))
        for lst in out:
            lst.sort()
        for lst in inn:
            lst.sort()
        self._out = out
        self._in = inn

    def successors(self, v):
        if self._out is None:
            self._build_adj()
        return [t for (t, _) in self._out[self._vid(v)]]

    def predecessors(self, v):
        if self._in is None:
            self._build_adj()
        return [s for (s, _) in self._in[self._vid(v)]]

    @property
    def vs(self):
        return VertexSeq(self)

    @property
    def es(self):
        return EdgeSeq(self)

    def copy(self):
        return self._subgraph(None)

    def induced_subgraph(self, vertices, implementation="auto"):
        if isinstance(vertices, VertexSeq):
            keep = list(vertices._indices())
        else:
            keep = [self._vid(v) for v in vertices]
        return self._subgraph(sorted(keep))

    subgraph = induced_subgraph

    def _subgraph(self, keep):
        g = Graph()
        if keep is None:
            g._nv = self._nv
            g._vattr = {k: list(v) for k, v in self._vattr.items()}
            g._name2idx = dict(self._name2idx)
            remap = None
        else:
            g._nv = len(keep)
            g._vattr = {k: [v[i] for i in keep] for k, v in self._vattr.items()}
            names = g._vattr["name"]
            g._name2idx = {names[i]: i for i in range(g._nv) if names[i] is not None}
            remap = [-1] * self._nv
            for new, old in enumerate(keep):
                remap[old] = new

        src, dst, alive = self._src, self._dst, self._alive
        if remap is None:
            eidx = [e for e in range(len(src)) if alive[e]]
            g._src = [src[e] for e in eidx]
            g._dst = [dst[e] for e in eidx]
        else:
            eidx = [e for e in range(len(src)) if alive[e] and remap[src[e]] >= 0 and remap[dst[e]] >= 0]
            g._src = [remap[src[e]] for e in eidx]
            g._dst = [remap[dst[e]] for e in eidx]
        g._alive = [True] * len(eidx)
        g._eattr = {k: [v[e] for e in eidx] for k, v in self._eattr.items()}
        g._pair2eid = {(g._src[i], g._dst[i]): i for i in range(len(eidx))}
        return g

    # -- shortest path ---------------------------------------------------------------------------------------------------
    def get_shortest_paths(self, v, to=None, weights=None, mode=OUT, output="vpath"):
        """igraph_get_shortest_paths_dijkstra restated (single source, mode OUT)."""
        if self._out is None:
            self._build_adj()
        src = self._vid(v)
        if to is None:
            targets = list(range(self._nv))
        elif isinstance(to, (list, tuple)):
            targets = [self._vid(t) for t in to]
        else:
            targets = [self._vid(to)]

        if isinstance(weights, str):
            w = self._eattr[weights]
        elif weights is None:
            w = None
        else:
            w = list(weights)

        dist = {src: 0.0}
        parent_edge = {src: -1}
        done = set()
        to_reach = set(targets)
        counter = 0
        heap = [(0.0, counter, src)]
        out = self._out
        while heap and to_reach:
            d, _, u = heapq.heappop(heap)
            if u in done or d != dist[u]:
                continue
            done.add(u)
            to_reach.discard(u)
            for (t, e) in out[u]:
                alt = d + (1.0 if w is None else w[e])
                cur = dist.get(t)
                if cur is None or alt < cur:
                    dist[t] = alt
                    parent_edge[t] = e
                    counter += 1
                    heapq.heappush(heap, (alt, counter, t))

        res = []
        for t in targets:
            if t not in done:
                res.append([])
                continue
            if output == "vpath":
                path = [t]
                cur = t
                while parent_edge[cur] != -1:
                    cur = self._src[parent_edge[cur]]
                    path.append(cur)
                path.reverse()
            else:
                path = []
                cur = t
                while parent_edge[cur] != -1:
                    path.append(parent_edge[cur])
                    cur = self._src[parent_edge[cur]]
                path.reverse()
            res.append(path)
        return res
