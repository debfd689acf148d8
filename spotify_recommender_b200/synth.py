"""Seeded synthetic inputs in the Spotify-tracks schema.

Feature order follows the reference (Song.h:17-19, DataManager.cpp:156-159,299):
danceability, energy, key, loudness, mode, speechiness, acousticness,
instrumentalness, liveness, valence, tempo, genre_id/(G-1).

The dense generator is block-counter based: block b (ROWS_PER_BLOCK rows) is
drawn from its own generator seeded with (seed, b), so any row-shard can be
produced on any rank without generating the rows before it.
"""
from __future__ import annotations

import numpy as np

FEATURE_COUNT = 12  # reference Song.h:12
ROWS_PER_BLOCK = 1 << 18
N_GENRES = 114
_SKEWED = (5, 7, 8)  # speechiness, instrumentalness, liveness: mass near 0
_KEY, _MODE, _GENRE = 2, 4, 11


def _block(b: int, n_total: int, seed: int) -> np.ndarray:
    lo = b * ROWS_PER_BLOCK
    rows = min(ROWS_PER_BLOCK, n_total - lo)
    rng = np.random.Generator(np.random.Philox(key=[seed, b]))
    u = rng.random((ROWS_PER_BLOCK, FEATURE_COUNT), dtype=np.float32)[:rows]
    out = np.empty((rows, FEATURE_COUNT), dtype=np.float32)
    out[:] = np.floor(u * np.float32(1000.0)) / np.float32(1000.0)
    for j in _SKEWED:
        out[:, j] = np.floor((u[:, j] ** 3) * np.float32(1000.0)) / np.float32(1000.0)
    out[:, _KEY] = np.floor(u[:, _KEY] * 12.0).clip(0, 11) / np.float32(11.0)
    out[:, _MODE] = (u[:, _MODE] < 0.64).astype(np.float32)
    per_genre = max(1, -(-n_total // N_GENRES))
    gid = (np.arange(lo, lo + rows, dtype=np.int64) // per_genre).clip(0, N_GENRES - 1)
    out[:, _GENRE] = gid.astype(np.float32) / np.float32(N_GENRES - 1)
    return out


def features(n_total: int, lo: int = 0, hi: int | None = None, seed: int = 42) -> np.ndarray:
    """Rows [lo, hi) of the n_total x 12 synthetic store, C-contiguous float32."""
    hi = n_total if hi is None else hi
    assert 0 <= lo <= hi <= n_total
    out = np.empty((hi - lo, FEATURE_COUNT), dtype=np.float32)
    b0, b1 = lo // ROWS_PER_BLOCK, -(-hi // ROWS_PER_BLOCK) if hi > lo else lo // ROWS_PER_BLOCK
    for b in range(b0, b1):
        blk = _block(b, n_total, seed)
        s = b * ROWS_PER_BLOCK
        a, z = max(lo, s), min(hi, s + blk.shape[0])
        out[a - lo:z - lo] = blk[a - s:z - s]
    return out


def rows(idx, n_total: int, seed: int = 42) -> np.ndarray:
    """Arbitrary rows of the synthetic store (regenerates only their blocks)."""
    idx = np.asarray(idx, dtype=np.int64)
    out = np.empty((idx.size, FEATURE_COUNT), dtype=np.float32)
    for b in np.unique(idx // ROWS_PER_BLOCK):
        blk = _block(int(b), n_total, seed)
        m = (idx // ROWS_PER_BLOCK) == b
        out[m] = blk[idx[m] - b * ROWS_PER_BLOCK]
    return out


def uniform(n: int, seed: int = 42) -> np.ndarray:
    """Plain U[0,1) n x 12 matrix (the survey's tie-free stress input)."""
    rng = np.random.Generator(np.random.Philox(key=[seed, 0x5eed]))
    return rng.random((n, FEATURE_COUNT), dtype=np.float32)


def query_indices(nq: int, n: int) -> np.ndarray:
    """In-database query songs: (q*7919 + 13) mod n."""
    q = np.arange(nq, dtype=np.int64)
    return ((q * 7919 + 13) % n).astype(np.int32)


def adversarial(n: int = 4096, seed: int = 7) -> np.ndarray:
    """Tie / degenerate cases the reference semantics must survive:
    exact duplicate rows, an all-zero row, rows scaled by a constant (cos == 1
    after clamping), tiny-norm rows around the 1e-8 denominator cut, negative
    values, and a block of identical rows (mass ties)."""
    rng = np.random.Generator(np.random.Philox(key=[seed, 1]))
    f = (np.floor(rng.random((n, FEATURE_COUNT), dtype=np.float32) * 1000) / 1000).astype(np.float32)
    f[5] = f[3]                       # duplicates of a likely query
    f[n // 2] = f[3]
    f[17] = 0.0                       # zero vector: score 0 (Recommender.cu:271)
    f[21] = f[3] * np.float32(2.0)    # scaled copies
    f[22] = f[3] * np.float32(0.5)
    f[23] = f[3] * np.float32(3.0)
    f[30] = np.float32(1e-6) * f[4]   # small but valid denominator
    f[31] = np.float32(1e-9) * f[4]   # denominator <= 1e-8 => 0
    f[32] = np.float32(3e-5) * f[4]
    f[40] = -f[3]                     # cos == -1
    f[41, ::2] *= np.float32(-1.0)
    f[100:164] = f[99]                # 65 identical rows
    f[200:232] = np.float32(0.5)      # constant rows: all mutually cos == 1
    return np.ascontiguousarray(f)


def mt19937_uniform(count: int, seed: int = 42) -> np.ndarray:
    """libstdc++ `std::uniform_real_distribution<float>(0,1)` over
    `std::mt19937(seed)` -- the generator of SURVEY Appendix A cases D1/D2.
    generate_canonical<float,24> draws one 32-bit word, converts it to float
    (round-to-nearest), divides by 2^32 and maps a rounded-up 1.0 to
    nextafter(1, 0)."""
    bg = np.random.MT19937()
    bg._legacy_seeding(seed)
    raw = bg.random_raw(count).astype(np.uint32)
    v = raw.astype(np.float32) / np.float32(4294967296.0)
    v[v >= np.float32(1.0)] = np.nextafter(np.float32(1.0), np.float32(0.0))
    return v


def spotify_csv(path: str, n_rows: int = 114000, seed: int = 42, bad_rows: bool = True) -> None:
    """A Spotify-tracks-schema CSV (21 columns; the 15 the reference requires at
    DataManager.cpp:121-125 among them), genres in equal consecutive blocks.
    With bad_rows one row has an empty track_name so the reference reports
    "Valid songs: n-1 out of n" (README.md:280)."""
    rng = np.random.Generator(np.random.Philox(key=[seed, 2]))
    cols = ["", "track_id", "artists", "album_name", "track_name", "popularity", "duration_ms",
            "explicit", "danceability", "energy", "key", "loudness", "mode", "speechiness",
            "acousticness", "instrumentalness", "liveness", "valence", "tempo", "time_signature",
            "track_genre"]
    per_genre = max(1, -(-n_rows // N_GENRES))
    u = rng.random((n_rows, 12))
    with open(path, "w", encoding="utf-8") as fh:
        fh.write(",".join(cols) + "\n")
        for i in range(n_rows):
            g = min(i // per_genre, N_GENRES - 1)
            name = f"Track {i}"
            if bad_rows and i == n_rows // 3:
                name = ""
            if i % 97 == 0:
                name = f'"Track {i}, Pt. 2"' if name else name
            r = u[i]
            fh.write(
                f"{i},id{i:07d},Artist {i % 5003};Feat {i % 13},Album {i % 9001},{name},"
                f"{int(r[0] * 100)},{120000 + int(r[1] * 240000)},{'True' if r[2] < 0.1 else 'False'},"
                f"{r[3]:.3f},{r[4]:.3f},{int(r[5] * 12) % 12},{-60.0 + 60.0 * r[6]:.3f},"
                f"{1 if r[7] < 0.64 else 0},{r[8] ** 3:.4f},{r[9]:.4f},{r[10] ** 3:.6f},"
                f"{r[11] ** 2:.4f},{r[0] * r[3]:.3f},{50.0 + 170.0 * r[1]:.3f},{3 + int(r[2] * 3) % 3},"
                f"genre_{g:03d}\n")
