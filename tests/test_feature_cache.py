"""CPU: the host-side boundary pieces that need no GPU — the versioned feature cache and the adaptor that turns the
reference's image DataManager into the cached-feature manager the trainers consume (train.py:89, utils/data_manager.py)."""
import os
import types

import pytest
import torch
from torch.utils.data import DataLoader, Dataset

from clip_gp_b200.trainers import CACHE_VERSION, FeatureDataManager, as_feature_manager


def make_dm():
    g = torch.Generator().manual_seed(0)
    return FeatureDataManager(text_embeddings=torch.randn(5, 3, 16, generator=g), features_train=torch.randn(20, 16, generator=g),
                              labels_train=torch.randint(0, 5, (20,), generator=g), features_test=torch.randn(11, 16, generator=g),
                              labels_test=torch.randint(0, 5, (11,), generator=g), classnames=[f"c{i}" for i in range(5)])


def test_cache_round_trip_and_validation(tmp_path):
    dm = make_dm()
    path = os.path.join(tmp_path, "sub", "cache.pt")
    dm.save(path, meta={"backbone": "ViT-B/16", "shots": 4, "seed": 1})
    back = FeatureDataManager.load(path, expect_meta={"backbone": "ViT-B/16", "shots": 4})
    for k in ("text_embeddings", "features_train", "labels_train", "features_test", "labels_test"):
        assert torch.equal(getattr(back, k), getattr(dm, k))
    assert back.features_val is None and back.classnames == dm.classnames and back.meta["seed"] == 1
    assert back.labels_train.dtype == torch.int64 and back.features_train.dtype == torch.float32
    assert as_feature_manager(path).num_classes == 5                        # a cache path is accepted wherever a manager is
    with pytest.raises(ValueError, match="shots"):
        FeatureDataManager.load(path, expect_meta={"shots": 16})            # stale cache detected through the metadata
    blob = torch.load(path, weights_only=True)
    blob["version"] = CACHE_VERSION + 1
    torch.save(blob, path)
    with pytest.raises(ValueError, match="version"):
        FeatureDataManager.load(path)
    torch.save({"something": 1}, path)
    with pytest.raises(ValueError, match="not a clipgp feature cache"):
        FeatureDataManager.load(path)
    blob["version"] = CACHE_VERSION
    blob["tensors"]["labels_train"] = blob["tensors"]["labels_train"][:-1]
    torch.save(blob, path)
    with pytest.raises(ValueError, match="inconsistent"):
        FeatureDataManager.load(path)
    assert not os.path.exists(path + ".tmp")


class _DS(Dataset):
    def __init__(self, f, y):
        self.f, self.y = f, y

    def __len__(self):
        return len(self.y)

    def __getitem__(self, i):
        return {"img": self.f[i], "label": int(self.y[i])}


def test_adaptor_from_reference_data_manager():
    """The reference's DataManager exposes train_loader_x (shuffled, drop_last) / val_loader / test_loader of {"img","label"}
    batches and dataset.classnames (utils/data_manager.py); the adaptor runs the frozen encoder over them WITHOUT dropping the last
    partial train batch (adapter.py:895-903)."""
    dm = make_dm()
    enc = lambda x: 2.0 * x                                                  # stands in for clip_model.visual
    ref = types.SimpleNamespace(
        train_loader_x=DataLoader(_DS(dm.features_train, dm.labels_train), batch_size=8, shuffle=True, drop_last=True),
        test_loader=DataLoader(_DS(dm.features_test, dm.labels_test), batch_size=4), val_loader=None,
        dataset=types.SimpleNamespace(classnames=dm.classnames), num_classes=5)
    out = FeatureDataManager.from_reference(ref, dm.text_embeddings, encode_image=enc)
    assert out.features_train.shape[0] == 20                                # 20 = 2 * 8 + 4: nothing dropped
    assert torch.equal(out.features_train, 2.0 * dm.features_train) and torch.equal(out.labels_train, dm.labels_train)
    assert torch.equal(out.features_test, 2.0 * dm.features_test) and out.features_val is None and out.classnames == dm.classnames
    ref.text_embeddings = dm.text_embeddings
    assert as_feature_manager(ref).features_train.shape == (20, 16)         # what `build_trainer(config, data_manager)` does
    with pytest.raises(TypeError):
        as_feature_manager(object())
