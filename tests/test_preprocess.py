"""min-max normalisation of the preprocessing step (DataManager.cpp:270-301) and genre ids.
CPU part: the oracle against the reference's own output (golden fixture, and live when the reference
binary is here).  GPU part: the engine's kernels against the oracle, bit for bit, through the C ABI."""
import json
import os
import subprocess

import numpy as np
import pytest

from oracle_lib import ORACLE_DIR
from preprocess_ref import parse_csv, read_songs_bin
from spotify_recommender_b200 import synth

KAT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "preprocess_kat.json")
REF_CLI = os.path.join(ORACLE_DIR, "_ref", "recommender_cpu")


def first_appearance(names):
    seen = {}
    return np.array([seen.setdefault(g, len(seen)) for g in names], np.int32), len(seen)


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def edge_matrix(n=20000, seed=3):
    """Raw features with everything the arithmetic can meet: negative columns, a constant column, a column
    whose range is just below / just above the 1e-4 cut, signed zeros, NaN, +-inf, huge and tiny values."""
    rng = np.random.Generator(np.random.Philox(key=[seed, 9]))
    raw = rng.random((n, 11), dtype=np.float32)
    raw[:, 3] = -60.0 + 60.0 * raw[:, 3]          # loudness-like, negative
    raw[:, 4] = 1.0                                # constant -> 0.5
    raw[:, 5] = 0.25 + np.float32(9.9e-5) * raw[:, 5]   # range <= 1e-4 -> 0.5
    raw[:, 6] = 0.25 + np.float32(1.1e-4) * raw[:, 6]   # range just above the cut
    raw[:, 7] = np.where(raw[:, 7] < 0.5, np.float32(-0.0), raw[:, 7])  # minimum is a signed zero
    raw[:, 8] *= np.float32(1e-30)
    raw[:, 9] *= np.float32(3e38)
    raw[5, 0] = np.nan
    raw[6, 1] = np.inf
    raw[7, 2] = -np.inf
    raw[8, 10] = np.nan
    genre = rng.integers(0, 114, n).astype(np.int32)
    return raw, genre


# ---- CPU: the oracle is pinned on the reference ----------------------------------------
def test_oracle_matches_reference_fixture(oracle):
    kat = json.load(open(KAT))
    n = kat["n"]
    raw = np.array(kat["raw_bits"], np.uint32).view(np.float32).reshape(n, 11)
    want = np.array(kat["ref_features_bits"], np.uint32).reshape(n, 12)
    ids, ng = first_appearance(kat["genres"])
    assert ids.tolist() == kat["ref_genre_ids"] and ng == len(kat["ref_genre_map"])
    got, mm = oracle.minmax_normalize(raw, ids, ng)
    assert (bits(got) == want).all()
    assert (got[:, 4] == 0.5).all() and (got[:, 8] == 0.5).all()  # constant mode, liveness range <= 1e-4
    assert mm[3] < 0 and (mm[:11] <= mm[11:]).all()


@pytest.mark.skipif(not os.path.exists(REF_CLI), reason="reference CLI not built (oracle/Makefile, needs /root/reference)")
def test_oracle_matches_live_reference_preprocess(oracle, tmp_path):
    csv = tmp_path / "tracks.csv"
    synth.spotify_csv(str(csv), n_rows=3000)
    env = dict(os.environ, OMP_NUM_THREADS="1")  # the reference's genre ids are only deterministic at 1 thread
    r = subprocess.run([REF_CLI, "--preprocess", str(csv)], cwd=tmp_path, env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    ids, genre_ids, feats, genre_map = read_songs_bin(str(tmp_path / "songs_data.bin"))
    cids, genres, raw = parse_csv(str(csv))
    assert ids == cids
    fa, ng = first_appearance(genres)
    assert (fa == genre_ids).all() and ng == len(genre_map)
    got, _ = oracle.minmax_normalize(raw, fa, ng)
    assert (bits(got) == bits(feats)).all()


def test_oracle_edge_semantics(oracle):
    raw, genre = edge_matrix(2000)
    got, mm = oracle.minmax_normalize(raw, genre, 114)
    assert (got[:, 4] == 0.5).all() and (got[:, 5] == 0.5).all() and not (got[:, 6] == 0.5).all()
    assert np.isnan(got[5, 0]) and not np.isnan(np.delete(got[:, 0], 5)).any()   # NaN never becomes a min / max
    assert mm[11 + 1] == np.inf and (got[np.arange(2000) != 6, 1] == 0).all()     # finite / inf
    assert np.signbit(mm[7]) == False and mm[7] == 0                              # zero minimum is +0
    assert (got[:, 11] == genre.astype(np.float32) / np.float32(113)).all()
    one, _ = oracle.minmax_normalize(raw[:1], genre[:1], 1)                       # single row, single genre
    assert (one[0, :11] == 0.5).all() and one[0, 11] == np.float32(genre[0])


def test_genre_ids_modes():
    from spotify_recommender_b200.engine import genre_ids
    names = ["rock", "ambient", "rock", "jazz", "ambient", "blues", ""]
    fa, n1 = genre_ids(names, sorted_ids=False)
    so, n2 = genre_ids(names, sorted_ids=True)
    assert n1 == n2 == 5
    assert fa.tolist() == [0, 1, 0, 2, 1, 3, 4]
    assert so.tolist() == [4, 1, 4, 3, 1, 2, 0]           # "" < ambient < blues < jazz < rock
    perm = [5, 3, 0, 6, 1, 2, 4]
    so_p, _ = genre_ids([names[i] for i in perm], sorted_ids=True)
    assert so_p.tolist() == [so[i] for i in perm]          # independent of row order
    kat = json.load(open(KAT))
    assert genre_ids(kat["genres"], sorted_ids=False)[0].tolist() == kat["ref_genre_ids"]


# ---- GPU: the engine against the oracle, through the C ABI -------------------------------
@pytest.fixture()
def engine_factory():
    from spotify_recommender_b200.engine import Engine
    made = []

    def make():
        made.append(Engine(0))
        return made[-1]
    yield make
    for e in made:
        e.close()


@pytest.mark.gpu
def test_normalize_matches_reference_fixture_on_gpu(engine_factory):
    kat = json.load(open(KAT))
    n = kat["n"]
    raw = np.array(kat["raw_bits"], np.uint32).view(np.float32).reshape(n, 11)
    e = engine_factory()
    got, _ = e.normalize_features(raw, np.array(kat["ref_genre_ids"], np.int32), len(kat["ref_genre_map"]))
    assert (bits(got) == np.array(kat["ref_features_bits"], np.uint32).reshape(n, 12)).all()


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 7, 2048, 20000, 300001])
def test_normalize_matches_oracle_bitwise(engine_factory, oracle, n):
    raw, genre = edge_matrix(max(n, 16))
    raw, genre = raw[:n].copy(), genre[:n].copy()
    e = engine_factory()
    got, mm = e.normalize_features(raw, genre, 114)
    want, wmm = oracle.minmax_normalize(raw, genre, 114)
    assert (bits(mm) == bits(wmm)).all()
    same = (bits(got) == bits(want)) | (np.isnan(got) & np.isnan(want))
    assert same.all(), np.argwhere(~same)[:5]


@pytest.mark.gpu
def test_normalize_device_path_feeds_the_store(engine_factory, oracle):
    """raw values -> normalised features -> store, without leaving the GPU; queries agree with the oracle
    run on the oracle-normalised matrix."""
    import torch
    n = 1_000_000
    rng = np.random.Generator(np.random.Philox(key=[11, 1]))
    raw = rng.random((n, 11), dtype=np.float32)
    raw[:, 3] = -60 + 60 * raw[:, 3]
    raw[:, 10] = 50 + 170 * raw[:, 10]
    genre = (np.arange(n) * 114 // n).astype(np.int32)
    e = engine_factory()
    d_raw, d_genre = torch.from_numpy(raw).cuda(), torch.from_numpy(genre).cuda()
    d_out = torch.empty((n, 12), dtype=torch.float32, device="cuda")
    d_mm = torch.empty(22, dtype=torch.float32, device="cuda")
    e.normalize_features_dev(d_raw, d_genre, n, 114, d_out, d_mm, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    want, wmm = oracle.minmax_normalize(raw, genre, 114)
    assert (bits(d_out.cpu().numpy()) == bits(want)).all() and (bits(d_mm.cpu().numpy()) == bits(wmm)).all()
    col = d_out.cpu().numpy()
    assert col.min() == 0.0 and col.max() == 1.0
    e.load_features(d_out)
    q = synth.query_indices(16, n)
    gi, gs = e.query_by_index(q, 10)
    wi, ws = oracle.query_index(want, q, 10, threads=oracle.max_threads)
    assert (gi == wi).all() and (bits(gs) == bits(ws)).all()
