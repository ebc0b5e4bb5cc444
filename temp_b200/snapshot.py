"""Per-timestamp snapshot graphs without DGL, plus the seeded synthetic generator of the benchmark.

A ``Snapshot`` is the read-only slice of a reference ``DGLGraph`` that the hot path and its callers
touch (SURVEY.md section 8b "Graph objects"): ``edges()``, ``edata['type_s'|'norm']``,
``ndata['id'|'norm']``, ``nodes()``, ``number_of_nodes()``, ``ids``.  Construction follows the
reference's ``get_train_val_test_graph_at_t`` (utils/dataset.py:151-232): the node set of a
timestamp is the sorted unique of every subject/object in train+valid+test, the train graph keeps
only the original direction (``add_reverse = False``), edges stay in file order and
``norm = 1 / in_degree`` (``inf -> 0``, utils/utils.py:74-79).

In addition each snapshot carries a CSR-by-destination view (stable in edge order, so the
per-row summation order equals the reference's edge-id order) which is what the kernels consume.
"""
from __future__ import annotations

import os
from typing import Dict, Iterable, List, Optional, Tuple

import numpy as np

__all__ = ["Snapshot", "SnapshotStore", "SHAPES"]


class Snapshot(object):
    __slots__ = ("time", "node_ids", "src", "dst", "rel", "norm", "row_ptr", "csr_src", "csr_rel", "_ids",
                 "_torch", "_packed")

    def __init__(self, time: int, node_ids: np.ndarray, src: np.ndarray, dst: np.ndarray, rel: np.ndarray):
        n = int(node_ids.shape[0])
        self.time = int(time)
        self.node_ids = np.ascontiguousarray(node_ids, dtype=np.int64)
        self.src = np.ascontiguousarray(src, dtype=np.int64)
        self.dst = np.ascontiguousarray(dst, dtype=np.int64)
        self.rel = np.ascontiguousarray(rel, dtype=np.int64)
        deg = np.bincount(self.dst, minlength=n)
        with np.errstate(divide="ignore"):
            norm = (1.0 / deg.astype(np.float32)).astype(np.float32)
        norm[deg == 0] = 0.0
        self.norm = norm
        order = np.argsort(self.dst, kind="stable")
        self.row_ptr = np.zeros(n + 1, dtype=np.int32)
        np.cumsum(deg, out=self.row_ptr[1:])
        self.csr_src = self.src[order].astype(np.int32)
        self.csr_rel = self.rel[order].astype(np.int32)
        self._ids = None
        self._torch = None
        self._packed = None

    def packed_parts(self):
        """Per-snapshot arrays the window planner concatenates (cached: a snapshot appears in many windows)."""
        if self._packed is None:
            ids32 = self.node_ids.astype(np.int32)
            self._packed = (ids32, np.full(self.num_nodes, self.time, dtype=np.int32), np.diff(self.row_ptr),
                            ids32[self.csr_src], np.full(self.num_nodes, -1, dtype=np.int32))
        return self._packed

    # ---- sizes ------------------------------------------------------------------------------
    @property
    def num_nodes(self) -> int:
        return int(self.node_ids.shape[0])

    @property
    def num_edges(self) -> int:
        return int(self.src.shape[0])

    def number_of_nodes(self) -> int:
        return self.num_nodes

    def number_of_edges(self) -> int:
        return self.num_edges

    # ---- DGL-like read-only surface (torch tensors, built lazily) -------------------------------
    def _t(self):
        if self._torch is None:
            import torch
            self._torch = {
                "src": torch.from_numpy(self.src), "dst": torch.from_numpy(self.dst),
                "type_s": torch.from_numpy(self.rel),
                "enorm": torch.from_numpy(self.norm[self.dst]).view(-1, 1),
                "id": torch.from_numpy(self.node_ids).view(-1, 1),
                "nnorm": torch.from_numpy(self.norm).view(-1, 1),
            }
        return self._torch

    def nodes(self):
        import torch
        return torch.arange(self.num_nodes, dtype=torch.long)

    def edges(self):
        t = self._t()
        return t["src"], t["dst"]

    @property
    def edata(self):
        t = self._t()
        return {"type_s": t["type_s"], "norm": t["enorm"]}

    @property
    def ndata(self):
        t = self._t()
        return {"id": t["id"], "norm": t["nnorm"]}

    @property
    def ids(self) -> Dict[int, int]:
        """local node index -> global entity id (utils/dataset.py:226-231)."""
        if self._ids is None:
            self._ids = {i: int(v) for i, v in enumerate(self.node_ids.tolist())}
        return self._ids

    def edge_subset(self, edge_idx: np.ndarray) -> "Snapshot":
        """``edge_subgraph(idx, preserve_nodes=True)`` + recomputed norms (models/DynamicRGCN.py:82-89)."""
        return Snapshot(self.time, self.node_ids, self.src[edge_idx], self.dst[edge_idx], self.rel[edge_idx])


# (num_ents, num_rels, mean edges, std edges, nodes/edges ratio, dst zipf exponent)  -- SURVEY Appendix C
SHAPES = {
    "icews14": dict(M=7128, R=230, mu_e=199.5, sd_e=60.9, node_ratio=1.14, dst_exp=0.75, pop_exp=1.2),
    "icews05-15": dict(M=10488, R=251, mu_e=91.9, sd_e=32.7, node_ratio=1.175, dst_exp=0.75, pop_exp=1.2),
    "gdelt": dict(M=500, R=20, mu_e=7474.5, sd_e=1843.2, node_ratio=0.0667, dst_exp=0.85, pop_exp=0.2),
}


class SnapshotStore(object):
    """time -> Snapshot for the train / valid / test splits (the three ``graph_dict``s of main.py:40)."""

    def __init__(self, num_ents: int, num_rels: int, train: Dict[int, Snapshot],
                 valid: Optional[Dict[int, Snapshot]] = None, test: Optional[Dict[int, Snapshot]] = None,
                 name: str = ""):
        self.num_ents, self.num_rels = int(num_ents), int(num_rels)
        self.train = train
        self.valid = valid if valid is not None else {}
        self.test = test if test is not None else {}
        self.times: List[int] = list(train.keys())
        self.name = name

    # ---- from the reference's on-disk format ----------------------------------------------------
    @staticmethod
    def _read_quads(path: str) -> np.ndarray:
        if not os.path.exists(path) or os.path.getsize(path) == 0:
            return np.zeros((0, 4), dtype=np.int64)
        try:
            import pandas as pd
            arr = pd.read_csv(path, sep=r"\s+", header=None, usecols=[0, 1, 2, 3], dtype=np.int64).values
        except ImportError:  # pragma: no cover
            arr = np.loadtxt(path, dtype=np.int64, usecols=(0, 1, 2, 3), ndmin=2)
        return np.ascontiguousarray(arr, dtype=np.int64)

    @classmethod
    def from_quadruple_dir(cls, path: str) -> "SnapshotStore":
        with open(os.path.join(path, "stat.txt")) as f:
            parts = f.readline().split()
        num_ents, num_rels = int(parts[0]), int(parts[1])
        splits = [cls._read_quads(os.path.join(path, n + ".txt")) for n in ("train", "valid", "test")]
        all_times = np.unique(np.concatenate([s[:, 3] for s in splits]))
        # group rows per time keeping file order (stable sort on the time column)
        grouped = []
        for s in splits:
            order = np.argsort(s[:, 3], kind="stable")
            ss = s[order]
            lo = np.searchsorted(ss[:, 3], all_times, side="left")
            hi = np.searchsorted(ss[:, 3], all_times, side="right")
            grouped.append((ss, lo, hi))
        train, valid, test = {}, {}, {}
        for k, tim in enumerate(all_times.tolist()):
            parts = [g[0][g[1][k]:g[2][k]] for g in grouped]
            uniq = np.unique(np.concatenate([np.concatenate([p[:, 0], p[:, 2]]) for p in parts]))
            for part, out in zip(parts, (train, valid, test)):
                out[tim] = Snapshot(tim, uniq, np.searchsorted(uniq, part[:, 0]), np.searchsorted(uniq, part[:, 2]),
                                    part[:, 1])
        return cls(num_ents, num_rels, train, valid, test, name=os.path.basename(os.path.normpath(path)))

    # ---- binary cache of a parsed dataset (quadruple text -> cached binary) -------
    def to_npz(self, path: str) -> None:
        """One compressed .npz holding the three splits as flat arrays (per split: times, node / edge offsets per
        snapshot, node ids, local src / dst, relation ids).  ``from_npz`` rebuilds the same snapshots without parsing
        text; norms and CSR orders are derived data and recomputed on load exactly as from text."""
        out = {"meta": np.asarray([self.num_ents, self.num_rels], dtype=np.int64), "name": np.asarray(self.name)}
        for split, gd in (("train", self.train), ("valid", self.valid), ("test", self.test)):
            snaps = list(gd.values())
            cat = lambda xs: np.concatenate(xs) if xs else np.zeros(0, dtype=np.int64)
            out[split + "_times"] = np.asarray([s.time for s in snaps], dtype=np.int64)
            out[split + "_node_off"] = np.cumsum([0] + [s.num_nodes for s in snaps]).astype(np.int64)
            out[split + "_edge_off"] = np.cumsum([0] + [s.num_edges for s in snaps]).astype(np.int64)
            out[split + "_node_ids"] = cat([s.node_ids for s in snaps])
            out[split + "_src"] = cat([s.src for s in snaps])
            out[split + "_dst"] = cat([s.dst for s in snaps])
            out[split + "_rel"] = cat([s.rel for s in snaps])
        np.savez_compressed(path, **out)

    @classmethod
    def from_npz(cls, path: str) -> "SnapshotStore":
        with np.load(path, allow_pickle=False) as z:
            num_ents, num_rels = (int(x) for x in z["meta"])
            splits = []
            for split in ("train", "valid", "test"):
                times, no, eo = z[split + "_times"], z[split + "_node_off"], z[split + "_edge_off"]
                ids, src, dst, rel = (z[split + "_" + k] for k in ("node_ids", "src", "dst", "rel"))
                splits.append({int(t): Snapshot(int(t), ids[no[k]:no[k + 1]], src[eo[k]:eo[k + 1]], dst[eo[k]:eo[k + 1]],
                                                rel[eo[k]:eo[k + 1]]) for k, t in enumerate(times.tolist())})
            return cls(num_ents, num_rels, *splits, name=str(z["name"]))

    # ---- seeded synthetic sequences of the benchmark ------------------------------
    @classmethod
    def synthetic(cls, shape: str = "icews14", num_times: int = 16, scale: int = 1, seed: int = 20201116,
                  num_ents: Optional[int] = None) -> "SnapshotStore":
        """ICEWS14- / ICEWS05-15- / GDELT-shaped snapshot sequence; ``scale`` multiplies the entity
        count and the per-snapshot node and edge counts while keeping the degree shape."""
        cfg = SHAPES[shape]
        rng = np.random.default_rng(seed)
        M = int(num_ents if num_ents is not None else cfg["M"] * scale)
        R = cfg["R"]
        # entity popularity (Zipf): Gumbel top-k draws the active set without replacement
        log_pop = -cfg["pop_exp"] * np.log(np.arange(1, M + 1, dtype=np.float64))
        rel_p = 1.0 / np.arange(1, R + 1, dtype=np.float64)
        rel_p /= rel_p.sum()
        train = {}
        for tim in range(num_times):
            e_t = int(np.clip(rng.normal(cfg["mu_e"], cfg["sd_e"]), cfg["mu_e"] * 0.3, cfg["mu_e"] * 1.6) * scale)
            n_t = int(min(M, max(2, round(e_t * cfg["node_ratio"]))))
            keys = log_pop + rng.gumbel(size=M)
            active = np.sort(np.argpartition(-keys, n_t - 1)[:n_t])
            dst_p = 1.0 / np.arange(1, n_t + 1, dtype=np.float64) ** cfg["dst_exp"]
            dst_p /= dst_p.sum()
            perm = rng.permutation(n_t)
            dst = perm[rng.choice(n_t, size=e_t, p=dst_p)]
            src = rng.integers(0, n_t, size=e_t)
            rel = rng.choice(R, size=e_t, p=rel_p)
            train[tim] = Snapshot(tim, active, src, dst, rel)
        return cls(M, R, train, name="synthetic-%s-x%d" % (shape, scale))
