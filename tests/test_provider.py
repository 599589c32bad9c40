"""Row N2: the data provider.  tests/golden/dataset_tiny is a six-example dataset in the
reference's on-disk format, written with the reference's own `np_to_str`; tests/golden/reference_provider.npz is what
the reference's OWN `provider.load_batch` + `jitter_point_cloud` return for it with `np.random.seed(77)`
(generator: tests/golden/make_reference_provider_golden.py).  With the same seed the host half here must draw the
same points, and the device half must assemble the same batch."""
import json
import os

import numpy as np
import pytest

from helpers import GOLDEN

BASE = os.path.join(GOLDEN, "dataset_tiny")


def _expected():
    return np.load(os.path.join(GOLDEN, "reference_provider.npz"))


def _host(jitter):
    from alignnet_b200 import provider
    r = _expected()
    idx = provider.get_data_files(os.path.join(BASE, "split", "val.txt"))
    np.random.seed(77)
    return provider.read_host_batch(BASE, [idx[i] for i in r["order"]], int(r["num_points"]), jitter=jitter), r


def _gather(host, w):
    pts, off, idx = host[f"points{w}"], host[f"offsets{w}"], host[f"sample_idx{w}"]
    out = np.zeros(idx.shape + (3,), np.float32)
    for b in range(idx.shape[0]):
        ok = idx[b] >= 0
        out[b, ok] = pts[off[b] + idx[b, ok]]
    return out


def test_host_half_draws_what_the_reference_draws():
    host, r = _host(jitter=True)
    for k in ("translations", "rel_angles", "pc1_centers", "pc2_centers", "pc1_angles", "pc2_angles"):
        np.testing.assert_allclose(host[k], r[k], atol=1e-12, err_msg=k)
    for w in (1, 2):
        got = _gather(host, w)
        np.testing.assert_array_equal(got, r[f"pcs{w}"].astype(np.float32))             # same points, bit for bit
        np.testing.assert_allclose(got + host[f"jitter{w}"], r[f"pcs{w}_jittered"], atol=1e-6)
    assert (host["sample_idx1"][0] == -1).all() and np.abs(r["pcs1"][0]).max() == 0      # the empty cloud -> zeros


@pytest.mark.gpu
@pytest.mark.parametrize("jitter", [False, True])
def test_device_batch_matches_reference_load_batch(jitter):
    import __graft_entry__ as ge
    ge.build()
    from alignnet_b200 import provider
    host, r = _host(jitter)
    batch = provider.assemble_on_device(host)
    for w in (1, 2):
        ref = r[f"pcs{w}_jittered"] if jitter else r[f"pcs{w}"]
        np.testing.assert_allclose(batch[f"pcs{w}"].cpu().numpy(), ref, atol=1e-6 if jitter else 0, rtol=0)
    for k in ("translations", "rel_angles", "pc1_centers", "pc2_centers", "pc1_angles", "pc2_angles"):
        np.testing.assert_allclose(batch[k].cpu().numpy(), r[k], atol=1e-6, err_msg=k)


@pytest.mark.gpu
def test_prefetcher_feeds_the_engine():
    import __graft_entry__ as ge
    ge.build()
    import torch
    from alignnet_b200 import engine, provider
    idx = provider.get_data_files(os.path.join(BASE, "split", "val.txt"))
    np.random.seed(1)
    eng = engine.Engine(engine.shipped_arch(), "cuda:0", "fp32")
    n = 0
    for batch in provider.Prefetcher(BASE, idx * 2, batch_size=4, num_points=32, jitter=True):
        ep = eng.forward(batch["pcs1"], batch["pcs2"], True, 0.5, None, seed=n)
        loss = eng.backward(batch["pcs1"], batch["pcs2"], batch, ep)
        assert torch.isfinite(loss[0]).item()
        n += 1
    assert n == 3


def test_writer_round_trip(tmp_path):
    """The writer half of the format (pointcloud.py:247-258, 979-997): what `save_example` / `synth.write_dataset` write is
    what `load_meta` / `read_host_batch` read, bit for bit (np.savetxt's %.18e is exact for float64)."""
    from alignnet_b200 import provider, synth
    rng = np.random.default_rng(0)
    c1, c2 = rng.normal(size=(7, 4)).astype(np.float32), rng.normal(size=(0, 4)).astype(np.float32)
    start, end, tr = rng.normal(size=3), rng.normal(size=3), rng.normal(size=3)
    provider.save_example(str(tmp_path), 12, c1, c2, start, 0.25, end, -1.5, tr, 1 / 3, additional_meta={"trackids": [2]})
    t, ra, s0, s1, a0, a1 = provider.load_meta(str(tmp_path), 12)
    assert np.array_equal(t, tr) and np.array_equal(s0, start) and np.array_equal(s1, end)
    assert (ra, a0, a1) == (1 / 3, 0.25, -1.5)
    assert np.array_equal(np.load(tmp_path / "pointcloud1" / "00000012.npy"), c1)
    assert np.load(tmp_path / "pointcloud2" / "00000012.npy").shape == (0, 4)
    assert json.load(open(tmp_path / "meta" / "00000012.json"))["trackids"] == [2]
    assert provider.str_to_np(provider.np_to_str(tr)).tolist() == tr.tolist()
    base = tmp_path / "ds"
    synth.write_dataset(str(base), 10, seed=3, points_range=(20, 40), persons_prob=0.5)
    train, val = (provider.get_data_files(str(base / "split" / f"{k}.txt")) for k in ("train", "val"))
    assert train == list(range(8)) and val == [8, 9]
    np.random.seed(1)
    host = provider.read_host_batch(str(base), val, 16)
    assert host["translations"].shape == (2, 3) and np.isfinite(host["translations"]).all()
    for i in val:
        for w in (1, 2):
            c = np.load(base / f"pointcloud{w}" / f"{i:08d}.npy")
            assert c.ndim == 2 and c.shape[1] == 4 and 20 <= len(c) <= 40 and c.dtype == np.float32
        t, ra, s0, s1, a0, a1 = provider.load_meta(str(base), i)
        np.testing.assert_allclose(s1, s0 + t, atol=1e-12)                     # end = start + translation
        assert abs(a1 - (a0 + ra)) < 1e-12 and abs(ra) <= np.pi / 2
