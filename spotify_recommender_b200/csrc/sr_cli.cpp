// sr_cli.cpp -- batch front end for the B200 engine: what the reference's docs
// list as a wish ("batch mode", PROJECT_SUMMARY.md:305, ARCHITECTURE.md:359-364), next to its
// untouched --song / --id CLI.  Reads the songs_data.bin the reference's --preprocess writes.
//
//   sr_recommend --data songs_data.bin --ids ids.txt      [-n N] [--out recs.csv]
//   sr_recommend --data songs_data.bin --range LO HI      [-n N] [--out recs.csv]
//   sr_recommend --data songs_data.bin --all-pairs        [-n N] [--out recs.csv]
//
// Output CSV: query_index,rank,song_index,similarity  (similarity = the reference's value,
// Recommender.cu:256-273; order = descending similarity, ties by lower index).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "sr_recommender.hpp"

static void usage()
{
    std::cerr << "usage: sr_recommend --data songs_data.bin (--ids FILE | --range LO HI | --all-pairs) [-n N] [--out FILE]\n"
                 "  --ids FILE    one track_id per line\n"
                 "  --range LO HI query songs by position LO <= i < HI\n"
                 "  --all-pairs   every song is a query (neighbour table)\n";
}

int main(int argc, char **argv)
{
    std::string data, ids_file, out_path;
    long lo = -1, hi = -1;
    bool all = false;
    int topn = 10;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        if (a == "--data" && i + 1 < argc) data = argv[++i];
        else if (a == "--ids" && i + 1 < argc) ids_file = argv[++i];
        else if (a == "--range" && i + 2 < argc) { lo = atol(argv[++i]); hi = atol(argv[++i]); }
        else if (a == "--all-pairs") all = true;
        else if (a == "-n" && i + 1 < argc) topn = atoi(argv[++i]);
        else if (a == "--out" && i + 1 < argc) out_path = argv[++i];
        else { usage(); return 2; }
    }
    if (data.empty() || topn <= 0 || (ids_file.empty() && !all && lo < 0)) { usage(); return 2; }

    Recommender rec;
    std::streambuf *cout_buf = std::cout.rdbuf(std::cerr.rdbuf());  // progress lines go to stderr: stdout may be the CSV
    const bool ok = rec.initializeFromFile(data);
    std::cout.rdbuf(cout_buf);
    if (!ok) return 1;
    const int n = rec.getSongCount();
    std::vector<int> queries;
    if (all) { lo = 0; hi = n; }
    if (!ids_file.empty()) {
        std::ifstream in(ids_file);
        if (!in.is_open()) { std::cerr << "Error: Could not open " << ids_file << std::endl; return 1; }
        std::string line;
        while (std::getline(in, line)) {
            while (!line.empty() && (line.back() == '\r' || line.back() == ' ')) line.pop_back();
            if (line.empty()) continue;
            const int idx = rec.findSongByTrackId(line);
            if (idx < 0) { std::cerr << "Error: Song with track_id '" << line << "' not found" << std::endl; return 1; }
            queries.push_back(idx);
        }
    } else {
        if (lo < 0 || hi > n || lo >= hi) { std::cerr << "Error: Invalid range" << std::endl; return 1; }
        for (long i = lo; i < hi; ++i) queries.push_back((int)i);
    }
    FILE *out = out_path.empty() ? stdout : fopen(out_path.c_str(), "w");
    if (!out) { std::cerr << "Error: Could not create " << out_path << std::endl; return 1; }
    fprintf(out, "query_index,rank,song_index,similarity\n");
    const size_t chunk = 8192;  // one engine pass per chunk
    for (size_t off = 0; off < queries.size(); off += chunk) {
        std::vector<int> part(queries.begin() + off, queries.begin() + std::min(queries.size(), off + chunk));
        const auto res = rec.recommendBatch(part, topn);
        if (res.size() != part.size()) return 1;
        for (size_t q = 0; q < part.size(); ++q)
            for (size_t r = 0; r < res[q].size(); ++r)
                fprintf(out, "%d,%zu,%d,%.9g\n", part[q], r + 1, res[q][r].songIndex, res[q][r].similarity);
    }
    if (out != stdout) fclose(out);
    std::cerr << "Wrote recommendations for " << queries.size() << " queries" << std::endl;
    return 0;
}
