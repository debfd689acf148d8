"""Generates tests/golden/preprocess_kat.json: a small Spotify-schema CSV run through the UNMODIFIED
reference's `--preprocess` (oracle/_ref/recommender_cpu, built by oracle/Makefile from /root/reference,
OMP_NUM_THREADS=1), with what went in (raw values as std::stof reads them, genre names) and what came out
(genre ids, 12 normalised features per song, bit patterns).  SURVEY 8 f4 / DataManager.cpp:270-301.

    python tests/golden/make_preprocess_kat.py        (needs /root/reference; run in the build container)"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from preprocess_ref import parse_csv, read_songs_bin  # noqa: E402

COLS = ["", "track_id", "artists", "album_name", "track_name", "popularity", "duration_ms", "explicit",
        "danceability", "energy", "key", "loudness", "mode", "speechiness", "acousticness", "instrumentalness",
        "liveness", "valence", "tempo", "time_signature", "track_genre"]


def write_csv(path, n=320, seed=7):
    rng = np.random.Generator(np.random.Philox(key=[seed, 5]))
    genres = ["rock", "ambient", "jazz", "k-pop", "acoustic", "techno", "blues"]
    with open(path, "w", encoding="utf-8") as fh:
        fh.write(",".join(COLS) + "\n")
        for i in range(n):
            u = rng.random(12)
            g = genres[int(u[11] * len(genres)) % len(genres)]   # not in sorted order of first appearance
            name = f"Song {i}" if i != 100 else ""                # one invalid row (empty name)
            tempo = "n/a" if i == 200 else f"{40.0 + 180.0 * u[10]:.3f}"  # one invalid row (bad number)
            fh.write(f"{i},kat{i:05d},Artist {i % 17},Album {i % 29},{name},{int(u[0] * 100)},{180000 + i},False,"
                     f"{u[1]:.3f},{u[2]:.4f},{int(u[3] * 12) % 12},{-45.0 + 44.0 * u[4]:.3f},1,"      # mode constant -> 0.5
                     f"{u[5] ** 3:.4f},{u[6]:.5f},{u[7] ** 4:.6f},{0.25 + 5e-5 * u[8]:.6f},"           # liveness range <= 1e-4 -> 0.5
                     f"{u[9]:.3f},{tempo},4,{g}\n")


def main():
    exe = os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle", "_ref", "recommender_cpu")
    with tempfile.TemporaryDirectory() as d:
        csv_path = os.path.join(d, "kat.csv")
        write_csv(csv_path)
        env = dict(os.environ, OMP_NUM_THREADS="1")
        r = subprocess.run([exe, "--preprocess", csv_path], cwd=d, env=env, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        ids, genre_ids, feats, genre_map = read_songs_bin(os.path.join(d, "songs_data.bin"))
        cids, genres, raw = parse_csv(csv_path)
        assert ids == cids, "the test-side CSV reader keeps different rows than the reference"
    out = {
        "source": "reference --preprocess (CPU build of /root/reference, OMP_NUM_THREADS=1) on the CSV of make_preprocess_kat.py",
        "n": len(ids),
        "raw_bits": raw.view(np.uint32).ravel().tolist(),
        "genres": genres,
        "ref_genre_ids": genre_ids.tolist(),
        "ref_genre_map": {str(k): v for k, v in sorted(genre_map.items())},
        "ref_features_bits": feats.view(np.uint32).ravel().tolist(),
    }
    with open(os.path.join(HERE, "preprocess_kat.json"), "w") as fh:
        json.dump(out, fh)
    print("wrote", len(ids), "songs,", len(genre_map), "genres")


if __name__ == "__main__":
    main()
