/*
 * sr_song.h -- the data model the scoring engine's C++ host class consumes.
 *
 * When this header is compiled next to the reference tree (its directory on the
 * include path), the reference's own Song.h is used unchanged -- main.cpp and
 * DataManager.cpp keep their serialisation code (Song.h:35-77).  Stand-alone
 * (tests, other embedders) a field-compatible declaration is provided: same
 * member names, types and order as reference Song.h:21-33, no serialisation.
 */
#ifndef SR_SONG_H
#define SR_SONG_H

#if defined(__has_include)
#if __has_include("Song.h")
#include "Song.h"
#endif
#endif

#ifndef SONG_H
#define SONG_H
#include <string>

const int FEATURE_COUNT = 12; /* reference Song.h:12 */

struct Song {
    std::string track_id;
    std::string track_name;
    std::string artists;
    int genre_id = -1;
    float features[FEATURE_COUNT] = {0};
};
#endif /* SONG_H */

#endif /* SR_SONG_H */
