"""Test infrastructure for SURVEY 8 f4: read what the reference's `--preprocess` consumed and produced.

parse_csv():      the numeric part of DataManager::preprocessCSV (DataManager.cpp:100-262) for CSVs whose
                  key / mode columns are numeric -- raw feature values exactly as std::stof yields them,
                  genre names, and which rows the reference keeps.
read_songs_bin(): the songs_data.bin layout (DataManager.cpp:315-409, Song.h:36-77)."""
from __future__ import annotations

import csv
import struct

import numpy as np

FEATURE_COLS = ["danceability", "energy", "key", "loudness", "mode", "speechiness", "acousticness",
                "instrumentalness", "liveness", "valence", "tempo"]  # DataManager.cpp:156-159


def stof(text: str) -> np.float32:
    """std::stof = strtof: one rounding from the decimal string (numpy parses float32 strings the same way)."""
    return np.array([text], dtype="S").astype(np.float32)[0]


def parse_csv(path: str):
    ids, genres, raw = [], [], []
    with open(path, newline="", encoding="utf-8") as fh:
        rd = csv.reader(fh)
        header = next(rd)
        col = {name.strip(): i for i, name in enumerate(header)}
        for fields in rd:
            if len(fields) < len(header):
                continue
            tid, name, genre = fields[col["track_id"]].strip(), fields[col["track_name"]].strip(), fields[col["track_genre"]].strip()
            if not tid or not name or not genre:
                continue
            try:
                row = [stof(fields[col[c]].strip()) for c in FEATURE_COLS]
            except ValueError:
                continue
            ids.append(tid)
            genres.append(genre)
            raw.append(row)
    return ids, genres, np.array(raw, np.float32).reshape(len(ids), 11)


def read_songs_bin(path: str):
    blob = open(path, "rb").read()
    pos = 0

    def u64():
        nonlocal pos
        v = struct.unpack_from("<Q", blob, pos)[0]
        pos += 8
        return v

    def text():
        nonlocal pos
        n = u64()
        s = blob[pos:pos + n].decode("utf-8")
        pos += n
        return s

    n_songs, n_genres = u64(), u64()
    genre_map = {}
    for _ in range(n_genres):
        gid = struct.unpack_from("<i", blob, pos)[0]
        pos += 4
        genre_map[gid] = text()
    ids, genre_ids = [], np.empty(n_songs, np.int32)
    feats = np.empty((n_songs, 12), np.float32)
    for i in range(n_songs):
        ids.append(text())
        text()  # track_name
        text()  # artists
        genre_ids[i] = struct.unpack_from("<i", blob, pos)[0]
        pos += 4
        feats[i] = np.frombuffer(blob, np.float32, 12, pos)
        pos += 48
    assert pos == len(blob)
    return ids, genre_ids, feats, genre_map
