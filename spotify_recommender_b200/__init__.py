"""B200-native cosine-similarity + top-K scoring engine: the hot path of
Iamdarika/Spotify_recommender (Recommender.cu) rebuilt for sm_100a behind the
reference's Recommender.h interface.  See DESIGN.md."""
from . import synth  # noqa: F401

__all__ = ["synth", "engine", "build"]
