"""Case table shared by the golden generator (make_golden.py, runs the reference) and the tests
(which rebuild the same inputs for the oracle and for the CUDA path)."""
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def dataset_path(name):
    return os.path.join(HERE, "data", name)


def _c(name, module, dataset="tiny", D=128, n_bases=128, L=4, t_list=(9, 5, 2, 0), rec_only_last_layer=True,
       use_time_embedding=True, **kw):
    d = dict(name=name, module=module, dataset=dataset, D=D, n_bases=n_bases, L=L, t_list=list(t_list),
             rec_only_last_layer=rec_only_last_layer, use_time_embedding=use_time_embedding)
    d.update(kw)
    return d


CASES = [
    # --- config 1: static RGCN, seq_len 1 (order of t_list is kept, baselines/StaticRGCN.py:23-28)
    _c("srgcn_tiny_d128", "SRGCN", L=1, t_list=(3, 9, 0, 5)),
    _c("srgcn_tiny_d32_nb8_note", "SRGCN", D=32, n_bases=8, L=1, t_list=(4, 1), use_time_embedding=False),
    # --- config 2 family: GRU recurrence, uni-directional
    _c("grrgcn_tiny_d128_last", "GRRGCN"),
    _c("grrgcn_tiny_d128_last_note", "GRRGCN", use_time_embedding=False, t_list=(8, 8, 3)),
    _c("grrgcn_tiny_d128_full", "GRRGCN", rec_only_last_layer=False),
    _c("grrgcn_tiny_d128_full_note", "GRRGCN", rec_only_last_layer=False, use_time_embedding=False),
    _c("grrgcn_tiny_d32_nb8_type1", "GRRGCN", D=32, n_bases=8, type1=True),
    _c("grrgcn_tiny_d128_lambda", "GRRGCN", learnable_lambda=True, L=5),
    _c("grrgcn_tiny_d200_nb100", "GRRGCN", D=200, n_bases=100),
    _c("grrgcn_tiny_d128_L1", "GRRGCN", L=1, t_list=(6, 2)),
    # --use-embed-for-non-active: entities without an edge at t keep their input embedding in the all-entity table
    _c("grrgcn_tiny_d128_embed_nonactive", "GRRGCN", use_embed_for_non_active=True),
    _c("bigrrgcn_tiny_d128_embed_nonactive", "BiGRRGCN", use_embed_for_non_active=True, t_list=(8, 4, 1)),
    _c("sargcn_tiny_d128_embed_nonactive", "SARGCN", use_embed_for_non_active=True, t_list=(9, 3)),
    _c("srgcn_tiny_d128_embed_nonactive", "SRGCN", L=1, t_list=(2, 7), use_embed_for_non_active=True),
    # --- linear recurrence flavour
    _c("rrgcn_tiny_d128_full", "RRGCN", rec_only_last_layer=False),
    _c("rrgcn_tiny_d128_last", "RRGCN"),
    # --- config 3 family: bidirectional
    _c("bigrrgcn_tiny_d128_last", "BiGRRGCN"),
    _c("bigrrgcn_tiny_d128_full", "BiGRRGCN", rec_only_last_layer=False, t_list=(9, 6, 1)),
    _c("bigrrgcn_tiny_d200_nb100", "BiGRRGCN", D=200, n_bases=100, t_list=(7, 4)),
    _c("birrgcn_tiny_d128_full", "BiRRGCN", rec_only_last_layer=False),
    # --- config 4 family: attention over the time axis
    _c("sargcn_tiny_d128_last", "SARGCN"),
    _c("sargcn_tiny_d128_full", "SARGCN", rec_only_last_layer=False),
    _c("sargcn_tiny_d128_lambda", "SARGCN", learnable_lambda=True),
    _c("bisargcn_tiny_d128_last", "BiSARGCN"),
    _c("bisargcn_tiny_d64_full", "BiSARGCN", D=64, n_bases=64, rec_only_last_layer=False, t_list=(8, 5, 0)),
    # --- real ICEWS14 snapshots (first 12 timestamps), reference shapes D=128, n_bases=128, L=8
    _c("grrgcn_icews_d128_L8", "GRRGCN", dataset="icews14_head", L=8, t_list=(11, 10, 3)),
    _c("bigrrgcn_icews_d128_L8", "BiGRRGCN", dataset="icews14_head", L=8, t_list=(11, 6, 3)),
    _c("bisargcn_icews_d128_L8", "BiSARGCN", dataset="icews14_head", L=8, t_list=(11, 5)),
    # --- the whole real ICEWS14 split: windows at the start, in the middle and at the very end of the timeline
    #     (t in {0, 3, 7, 200, 364})
    _c("grrgcn_icews14_real_L8", "GRRGCN", dataset="icews14", L=8, t_list=(364, 200, 7, 3, 0)),
    _c("bigrrgcn_icews14_real_L8", "BiGRRGCN", dataset="icews14", L=8, t_list=(364, 200, 7, 3, 0)),
    # BASELINE config 5 on real data: GDELT snapshots (in-degrees in the hundreds, duplicate facts), seq_len 15, batch 2
    _c("grrgcn_gdelt_real_L15", "GRRGCN", dataset="gdelt_head", L=15, t_list=(16, 15)),
    # BASELINE config 3 as the reference ships it (n_bases = 100 => D = 200, 2 x 2 blocks) on real ICEWS05-15 snapshots
    _c("bigrrgcn_icews0515_real_nb100", "BiGRRGCN", dataset="icews0515_head", D=200, n_bases=100, L=8, t_list=(8, 7, 3)),
]

# (round 1 kept the two real-data cases above as oracle-only pins; both suites iterate over them now)
CPU_CASES = []

SAMPLER_CASES = [
    dict(name="sampler_tiny_seed123", dataset="tiny", times=[0, 4, 7], seed=123, negative_rate=5, num_pos_facts=3000),
    dict(name="sampler_tiny_subsample", dataset="tiny", times=[2, 9], seed=7, negative_rate=6, num_pos_facts=10),
    dict(name="sampler_icews_seed123", dataset="icews14_head", times=[3], seed=123, negative_rate=500, num_pos_facts=3000),
]

# Training-mode forward (models/DynamicRGCN.py:176-194) with dropout p = 0: edge sub-sampling of the final step (and of
# the history steps with --random-dropout), negative sampling, tail + head cross-entropy -> the scalar loss.
TRAIN_CASES = [
    dict(name="train_grrgcn_tiny", base="grrgcn_tiny_d128_last", seed=11, random_dropout=False),
    dict(name="train_grrgcn_tiny_embed_nonactive", base="grrgcn_tiny_d128_embed_nonactive", seed=17, random_dropout=False),
    dict(name="train_grrgcn_tiny_random_dropout", base="grrgcn_tiny_d128_last", seed=12, random_dropout=True),
    dict(name="train_rrgcn_tiny", base="rrgcn_tiny_d128_last", seed=13, random_dropout=False),
    dict(name="train_grrgcn_icews", base="grrgcn_icews_d128_L8", seed=5, random_dropout=True, negative_rate=20),
    # bidirectional (models/BiDynamicRGCN.py:165-187): forward history, backward history, centre step
    dict(name="train_bigrrgcn_tiny", base="bigrrgcn_tiny_d128_last", seed=21, random_dropout=False),
    dict(name="train_bigrrgcn_tiny_random_dropout", base="bigrrgcn_tiny_d128_last", seed=22, random_dropout=True),
    # static and attention families (baselines/StaticRGCN.py:36-47, models/SelfAttentionRGCN.py:122-139,
    # models/BiSelfAttentionRGCN.py:48-69)
    dict(name="train_srgcn_tiny", base="srgcn_tiny_d128", seed=31, random_dropout=False),
    dict(name="train_sargcn_tiny", base="sargcn_tiny_d128_last", seed=32, random_dropout=False),
    dict(name="train_sargcn_tiny_full_random_dropout", base="sargcn_tiny_d128_full", seed=33, random_dropout=True),
    dict(name="train_bisargcn_tiny_random_dropout", base="bisargcn_tiny_d128_last", seed=34, random_dropout=True),
    dict(name="train_bisargcn_icews", base="bisargcn_icews_d128_L8", seed=35, random_dropout=False, negative_rate=20),
    # BASELINE configs 5 and 3 on real snapshots (GDELT seq_len 15; ICEWS05-15 at D = 200 / n_bases = 100)
    dict(name="train_grrgcn_gdelt_real", base="grrgcn_gdelt_real_L15", seed=51, random_dropout=True, negative_rate=20,
         num_pos_facts=500),
    dict(name="train_bigrrgcn_icews0515_real_nb100", base="bigrrgcn_icews0515_real_nb100", seed=52, random_dropout=False,
         negative_rate=20),
]

# evaluate(t_list, val=True) of the unmodified reference -> filtered ranks (subject side then object side per graph) and the
# mean link-classification loss (models/DynamicRGCN.py:118-130, 196-220 over utils/evaluation.py:6-106).  ``empty_first``:
# the validation graph of the latest target timestamp -- the first one calc_metrics visits -- has its edges removed, which
# makes the reference's history index lag for every later graph of the batch (`continue` before `i += 1`).
RANK_CASES = [
    dict(name="rank_grrgcn_tiny", base="grrgcn_tiny_d128_last"),
    dict(name="rank_grrgcn_tiny_embed_nonactive", base="grrgcn_tiny_d128_embed_nonactive"),
    dict(name="rank_rrgcn_tiny_full", base="rrgcn_tiny_d128_full"),
    dict(name="rank_bigrrgcn_tiny", base="bigrrgcn_tiny_d128_last"),
    dict(name="rank_sargcn_tiny", base="sargcn_tiny_d128_last"),
    dict(name="rank_bisargcn_tiny", base="bisargcn_tiny_d128_last"),
    dict(name="rank_srgcn_tiny", base="srgcn_tiny_d128"),
    dict(name="rank_grrgcn_icews", base="grrgcn_icews_d128_L8"),
    dict(name="rank_bigrrgcn_icews", base="bigrrgcn_icews_d128_L8"),
    dict(name="rank_bisargcn_icews", base="bisargcn_icews_d128_L8"),
    dict(name="rank_grrgcn_icews14_real", base="grrgcn_icews14_real_L8"),
    dict(name="rank_grrgcn_tiny_empty_first", base="grrgcn_tiny_d128_last", empty_first=True),
    dict(name="rank_bigrrgcn_tiny_empty_first", base="bigrrgcn_tiny_d128_last", empty_first=True),
    dict(name="rank_sargcn_tiny_empty_first", base="sargcn_tiny_d128_last", empty_first=True),
    dict(name="rank_grrgcn_icews_empty_first", base="grrgcn_icews_d128_L8", empty_first=True),
    dict(name="rank_grrgcn_gdelt_real", base="grrgcn_gdelt_real_L15"),
    dict(name="rank_bigrrgcn_icews0515_real_nb100", base="bigrrgcn_icews0515_real_nb100"),
]

# Training-loss pins the oracle alone is held to for now (the GPU suite reaches these configurations through the oracle:
# tests/test_gpu_parity.py::test_autograd_fallback_covers_every_recurrent_configuration).
CPU_TRAIN_CASES = [
    dict(name="train_grrgcn_tiny_full", base="grrgcn_tiny_d128_full", seed=41, random_dropout=True),
    dict(name="train_rrgcn_tiny_full", base="rrgcn_tiny_d128_full", seed=42, random_dropout=True),
    dict(name="train_bigrrgcn_tiny_full", base="bigrrgcn_tiny_d128_full", seed=43, random_dropout=True),
    dict(name="train_birrgcn_tiny_full", base="birrgcn_tiny_d128_full", seed=44, random_dropout=False),
    dict(name="train_grrgcn_tiny_type1", base="grrgcn_tiny_d32_nb8_type1", seed=45, random_dropout=True),
    dict(name="train_grrgcn_tiny_lambda", base="grrgcn_tiny_d128_lambda", seed=46, random_dropout=True),
    dict(name="train_bigrrgcn_tiny_nb100", base="bigrrgcn_tiny_d200_nb100", seed=47, random_dropout=False),
]


# Post-ensemble / impute variants of the GRU families (models/PostDynamicRGCN.py, models/PostBiDynamicRGCN.py; main.py:57-72
# selects them with --post-ensemble / --impute): evaluate_embed (local + recurrent stream of the target graphs),
# get_all_embeds_Gt of every batch item (the imputed table, or the local / recurrent pair) and evaluate() ranks.
POST_CASES = [
    dict(name="impute_grrgcn_tiny", base="grrgcn_tiny_d128_last", impute=True, post_ensemble=False),
    dict(name="post_grrgcn_tiny", base="grrgcn_tiny_d128_last", impute=False, post_ensemble=True),
    dict(name="post_impute_grrgcn_tiny_full", base="grrgcn_tiny_d128_full", impute=True, post_ensemble=True),
    dict(name="impute_bigrrgcn_tiny", base="bigrrgcn_tiny_d128_last", impute=True, post_ensemble=False),
    dict(name="post_bigrrgcn_tiny", base="bigrrgcn_tiny_d128_last", impute=False, post_ensemble=True),
    dict(name="post_impute_bigrrgcn_tiny_full", base="bigrrgcn_tiny_d128_full", impute=True, post_ensemble=True),
    dict(name="post_impute_grrgcn_icews", base="grrgcn_icews_d128_L8", impute=True, post_ensemble=True),
]
