"""Bit-exact index streams: inducing-point selection on numpy's legacy global MT19937 stream and the
DataLoader(shuffle=True) minibatch order, checked against the REAL numpy / torch objects and the committed golden arrays."""
import os

import numpy as np
import torch
from torch.utils.data import DataLoader, TensorDataset

import ggp_b200.synthetic as syn

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "rng_streams.npz"))


def test_inducing_index_selection_matches_reference_stream():
    for split, N, prop, M in [(0, 9568, 0.8, 500), (3, 545, 0.9, 100), (7, 1000, 0.8, 20)]:
        rng, tr, te = syn.dataset_split_stream(N, split, prop)
        idx = syn.select_inducing_indices(len(tr), M, rng)
        assert np.array_equal(tr[:64], GOLD[f"split{split}_train_head"])
        assert np.array_equal(idx, GOLD[f"split{split}_zidx"])
        assert idx.dtype == GOLD[f"split{split}_zidx"].dtype
        # and against the live global stream, exactly as experiments/regression.py:203,83 + utils/dataset.py:62-63 consume it
        np.random.seed(999)
        ind = np.arange(N)
        np.random.seed(173 + split)
        np.random.shuffle(ind)
        live = np.random.randint(0, int(N * prop), M)
        assert np.array_equal(idx, live)
    assert len(set(GOLD["split0_zidx"].tolist())) < 500  # with replacement: duplicates are the norm


def test_minibatch_order_matches_dataloader():
    for seed, n, bs in [(42, 7654, 1024), (7, 100, 32)]:
        torch.manual_seed(seed)
        # DataLoader.__iter__ draws its base_seed first, then RandomSampler draws the permutation seed
        for ep in range(2):
            _base_seed = torch.empty((), dtype=torch.int64).random_()
            batches = syn.minibatch_indices(n, bs)
            got = torch.cat(batches).numpy()
            assert np.array_equal(got, GOLD[f"loader_seed{seed}_n{n}_bs{bs}_epoch{ep}"])
            assert len(batches) == (n + bs - 1) // bs and len(batches[-1]) == n - bs * (len(batches) - 1)
        torch.manual_seed(seed)
        dl = DataLoader(TensorDataset(torch.arange(n), torch.arange(n)), batch_size=bs, shuffle=True)
        live = torch.cat([a for a, _ in dl]).numpy()
        assert np.array_equal(live, GOLD[f"loader_seed{seed}_n{n}_bs{bs}_epoch0"])


def test_synthetic_configs_have_the_named_shapes():
    c1 = syn.config1_demo_1d()
    assert c1["X"].shape == (1000, 1) and c1["Z"].shape == (20, 1)
    c2 = syn.config2_co2_shaped()
    assert c2["X"].shape == (545, 1) and c2["Z"].shape == (100, 1)
    c3 = syn.config3_power_shaped()
    assert c3["X"].shape == (7654, 4) and c3["X_test"].shape == (1914, 4) and c3["Z"].shape == (500, 4)
    c4 = syn.config4_large(N=4096, M=64)
    assert c4["X"].shape == (4096, 8) and c4["Z"].shape == (64, 8)
    c5 = syn.config5_classification(N=2048, M=32)
    assert set(np.unique(c5["y"])) <= {0.0, 1.0} and c5["X"].shape == (2048, 16)
