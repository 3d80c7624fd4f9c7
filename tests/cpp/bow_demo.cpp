// bow_demo -- drives orbx_shim::DBoW2::OrbVocabulary the way a loop closer built on the reference's vendored DBoW2 would:
// loadFromTextFile, transform(features, BowVector&, FeatureVector&, levelsup) per frame, score() of the first frame against all.
// usage: bow_demo vocabulary.txt descriptors.bin nframes n levelsup out.bin
//        descriptors.bin: nframes x n x 32 bytes.  out.bin, per frame: int32 nbow, nbow x (uint32 word, double value), int32 nfv,
//        per node: uint32 node, int32 count, count x uint32 feature; then nframes doubles: score(frame 0, frame f), twice
//        (one call per pair, then the one-launch database form).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "orbx_shim.hpp"

using namespace orbx_shim;

int main(int argc, char** argv)
{
    if (argc != 7) { std::fprintf(stderr, "usage: %s vocabulary.txt descriptors.bin nframes n levelsup out.bin\n", argv[0]); return 2; }
    const int nframes = std::atoi(argv[3]), n = std::atoi(argv[4]), levelsup = std::atoi(argv[5]);
    std::vector<uint8_t> raw((size_t)nframes * n * 32);
    FILE* f = std::fopen(argv[2], "rb");
    if (!f || std::fread(raw.data(), 1, raw.size(), f) != raw.size()) { std::fprintf(stderr, "cannot read %s\n", argv[2]); return 2; }
    std::fclose(f);
    try {
        DBoW2::OrbVocabulary voc;
        if (!voc.loadFromTextFile(argv[1])) { std::fprintf(stderr, "cannot load %s\n", argv[1]); return 2; }
        FILE* o = std::fopen(argv[6], "wb");
        if (!o) return 2;
        std::vector<DBoW2::BowVector> bows((size_t)nframes);
        for (int fr = 0; fr < nframes; fr++) {
            DBoW2::FeatureVector fv;
            if (fr % 2 == 0) {                              // the descriptor matrix as the extractor returns it
                Mat d(n, 32, raw.data() + (size_t)fr * n * 32);
                voc.transform(d, bows[(size_t)fr], fv, levelsup);
            } else {                                        // the reference's vector of 1 x 32 rows
                std::vector<Mat> rows;
                for (int i = 0; i < n; i++) rows.push_back(Mat(1, 32, raw.data() + ((size_t)fr * n + i) * 32));
                voc.transform(rows, bows[(size_t)fr], fv, levelsup);
                DBoW2::BowVector again;
                voc.transform(rows, again);                 // the overload without a feature vector
                if (again != bows[(size_t)fr]) { std::fprintf(stderr, "the two transform overloads disagree\n"); return 3; }
            }
            const int32_t nb = (int32_t)bows[(size_t)fr].size();
            std::fwrite(&nb, 4, 1, o);
            for (DBoW2::BowVector::const_iterator it = bows[(size_t)fr].begin(); it != bows[(size_t)fr].end(); ++it) {
                std::fwrite(&it->first, 4, 1, o);
                std::fwrite(&it->second, 8, 1, o);
            }
            const int32_t nf = (int32_t)fv.size();
            std::fwrite(&nf, 4, 1, o);
            for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it) {
                const int32_t c = (int32_t)it->second.size();
                std::fwrite(&it->first, 4, 1, o);
                std::fwrite(&c, 4, 1, o);
                std::fwrite(it->second.data(), 4, it->second.size(), o);
            }
        }
        for (int fr = 0; fr < nframes; fr++) { const double s = voc.score(bows[0], bows[(size_t)fr]); std::fwrite(&s, 8, 1, o); }
        const std::vector<double> all = voc.score(bows[0], bows);
        std::fwrite(all.data(), 8, all.size(), o);
        std::fclose(o);
        // word 0 is a node below the root; a single descriptor's word is one of its frame's words
        if (voc.size() == 0 || voc.getParentNode(0, 0) == 0 || voc.getWordWeight(0) < 0 || bows[0].count(voc.transform(Mat(1, 32, raw.data()))) != 1) return 3;
    } catch (const Error& e) {
        std::fprintf(stderr, "bow_demo: %s\n", e.what());
        return 1;
    }
    return 0;
}
