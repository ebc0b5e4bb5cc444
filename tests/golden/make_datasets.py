"""Generates the two small quadruple datasets the golden tests run on (committed output).

  data/tiny/         seeded synthetic: 40 entities, 6 relations, 10 timestamps; contains
                     multi-edges, zero-in-degree nodes, a near-empty snapshot and entities that
                     only occur in valid/test (exercises utils/dataset.py:151-232 of the reference).
  data/icews14/      the whole public ICEWS14 interpolation split (365 timestamps), same source, ids unchanged
  data/icews0515_head/ the first 17 timestamps of the public ICEWS05-15 interpolation split (M = 10 488, 251 relations)
  data/gdelt_head/   the first 17 timestamps of the public GDELT interpolation split, same source
  data/icews14_head/ the first 12 timestamps of the public ICEWS14 interpolation split as shipped
                     with the reference (/root/reference/interpolation/icews14, DATA not source),
                     ids unchanged (M = 7128, 230 relations).

Run from the repo root:  python tests/golden/make_datasets.py
"""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def write_quads(path, quads):
    with open(path, "w") as f:
        for h, r, t, tim in quads:
            f.write("%d\t%d\t%d\t%d\n" % (h, r, t, tim))


def make_tiny():
    rng = np.random.default_rng(20201116)
    M, R, T = 40, 6, 10
    out = os.path.join(HERE, "data", "tiny")
    os.makedirs(out, exist_ok=True)
    splits = {"train": [], "valid": [], "test": []}
    pop = 1.0 / np.arange(1, M + 1) ** 0.8
    pop /= pop.sum()
    for tim in range(T):
        n_train = 3 if tim == 4 else int(rng.integers(18, 34))
        act = rng.choice(M, size=int(rng.integers(10, 24)), replace=False, p=pop)
        for mode, n in (("train", n_train), ("valid", int(rng.integers(2, 6))), ("test", int(rng.integers(2, 6)))):
            for _ in range(n):
                pool = act if mode == "train" or rng.random() < 0.7 else np.arange(M)
                h, t = int(rng.choice(pool)), int(rng.choice(pool[: max(3, len(pool) // 2)]))
                splits[mode].append((h, int(rng.integers(0, R)), t, tim))
            if mode == "train" and n_train > 3:      # duplicate a few edges: multiplicity must be kept
                for q in list(splits["train"][-3:]):
                    splits["train"].append(q)
    for mode, quads in splits.items():
        write_quads(os.path.join(out, mode + ".txt"), quads)
    with open(os.path.join(out, "stat.txt"), "w") as f:
        f.write("%d\t%d\t%d\n" % (M, R, T))


def make_icews_head(src="/root/reference/interpolation/icews14", n_times=12, name="icews14_head"):
    out = os.path.join(HERE, "data", name)
    os.makedirs(out, exist_ok=True)
    for mode in ("train", "valid", "test"):
        quads = []
        with open(os.path.join(src, mode + ".txt")) as f:
            for line in f:
                p = line.split()
                if int(p[3]) < n_times:
                    quads.append((int(p[0]), int(p[1]), int(p[2]), int(p[3])))
        write_quads(os.path.join(out, mode + ".txt"), quads)
    with open(os.path.join(src, "stat.txt")) as f:
        m, r = f.readline().split()[:2]
    with open(os.path.join(out, "stat.txt"), "w") as f:
        f.write("%s\t%s\t%d\n" % (m, r, n_times))


if __name__ == "__main__":
    make_tiny()
    if os.path.isdir("/root/reference/interpolation/icews14"):
        make_icews_head()
        make_icews_head(n_times=365, name="icews14")      # the whole public split (1.5 MB of text): real windows at any t
    if os.path.isdir("/root/reference/interpolation/icews05-15"):
        make_icews_head(src="/root/reference/interpolation/icews05-15", n_times=17, name="icews0515_head")
    if os.path.isdir("/root/reference/interpolation/gdelt"):
        # the first 17 timestamps of the public GDELT split (500 entities, ~7 500 train facts per snapshot, in-degrees in the
        # hundreds, ~20 % duplicate facts): room for seq_len = 15 windows (BASELINE config 5)
        make_icews_head(src="/root/reference/interpolation/gdelt", n_times=17, name="gdelt_head")
