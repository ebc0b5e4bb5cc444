"""Shared test plumbing: rebuild a golden case's inputs for the oracle (and for the CUDA path)."""
import functools
import os

import numpy as np

from oracle import temp_oracle as orc
from tests.golden.cases import CASES, SAMPLER_CASES, dataset_path

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASE_BY_NAME = {c["name"]: c for c in CASES}


@functools.lru_cache(maxsize=None)
def oracle_graphs(dataset):
    path = dataset_path(dataset)
    m, r = orc.read_stat(path)
    train, valid, test = orc.build_graph_dicts(path)
    return m, r, train, valid, test


def oracle_config(case):
    m, r, train, _, _ = oracle_graphs(case["dataset"])
    return orc.OracleConfig(module=case["module"], num_ents=m, num_rels=r, num_times=len(train),
                            embed_size=case["D"], n_bases=case["n_bases"], seq_len=case["L"],
                            rec_only_last_layer=case["rec_only_last_layer"],
                            use_time_embedding=True if case["module"].endswith("SARGCN") else case["use_time_embedding"],
                            type1=case.get("type1", False), learnable_lambda=case.get("learnable_lambda", False),
                            use_embed_for_non_active=case.get("use_embed_for_non_active", False))


def oracle_model(case):
    cfg = oracle_config(case)
    _, _, train, _, _ = oracle_graphs(case["dataset"])
    return orc.OracleModel(cfg, orc.make_params(cfg), train)


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


# ---- product side ------------------------------------------------------------------------------
from argparse import Namespace


@functools.lru_cache(maxsize=None)
def product_store(dataset):
    from temp_b200.snapshot import SnapshotStore
    return SnapshotStore.from_quadruple_dir(dataset_path(dataset))


def product_args(case, impute=False, post_ensemble=False):
    return Namespace(module=case["module"], embed_size=case["D"], hidden_size=case["D"], n_bases=case["n_bases"],
                     train_seq_len=case["L"], test_seq_len=case["L"], dropout=0.1, num_layers=1, lr=1e-3,
                     rec_only_last_layer=case["rec_only_last_layer"], use_time_embedding=case["use_time_embedding"],
                     inv_temperature=0.1, type1=case.get("type1", False),
                     learnable_lambda=case.get("learnable_lambda", False), score_function="complex",
                     negative_rate=case.get("negative_rate", 5), num_pos_facts=case.get("num_pos_facts", 3000),
                     use_cuda=True, impute=impute, post_ensemble=post_ensemble, post_aggregation=False,
                     use_embed_for_non_active=case.get("use_embed_for_non_active", False))


def product_model(case, device="cuda", impute=False, post_ensemble=False):
    """The CUDA-backed shell for a golden case, parameters filled with the same exact-integer hash the
    golden generator and the oracle use (keyed by state_dict name)."""
    import torch
    from temp_b200.models import build_module
    store = product_store(case["dataset"])
    model = build_module(product_args(case, impute, post_ensemble), store.num_ents, store.num_rels, store.train, store.valid,
                         store.test)
    with torch.no_grad():
        for name, prm in model.state_dict().items():
            prm.copy_(torch.from_numpy(orc.fill_values(name, tuple(prm.shape))))
    return model.to(device).eval()
