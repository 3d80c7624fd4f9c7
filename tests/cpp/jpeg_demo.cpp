// jpeg_demo -- orbx_shim::JpegDecoder the way a FrameLoader built on it would use it: the files of a batch read from disk,
// decoded together, the frames handed on as matrices.
// usage: jpeg_demo out.bin file1.jpg file2.jpg ...     out.bin: int32 rows, cols, channels, then the frames' bytes one after the other
#include <cstdio>
#include <vector>

#include "orbx_shim.hpp"

using namespace orbx_shim;

int main(int argc, char** argv)
{
    if (argc < 3) { std::fprintf(stderr, "usage: %s out.bin file.jpg ...\n", argv[0]); return 2; }
    std::vector<std::vector<uint8_t> > files;
    for (int i = 2; i < argc; i++) {
        FILE* f = std::fopen(argv[i], "rb");
        if (!f) { std::fprintf(stderr, "cannot read %s\n", argv[i]); return 2; }
        std::vector<uint8_t> buf;
        uint8_t chunk[65536];
        size_t n;
        while ((n = std::fread(chunk, 1, sizeof(chunk), f)) > 0) buf.insert(buf.end(), chunk, chunk + n);
        std::fclose(f);
        files.push_back(buf);
    }
    try {
        int w = 0, h = 0, ch = 0;
        if (!JpegDecoder::supported(files[0], &w, &h, &ch)) { std::fprintf(stderr, "not a file the GPU path takes\n"); return 4; }
        JpegDecoder dec;
        std::vector<Mat> frames;
        dec.decode(files, frames);
        FILE* o = std::fopen(argv[1], "wb");
        if (!o) return 2;
        const int32_t hdr[3] = {frames[0].rows, frames[0].cols, frames[0].channels()};
        if (hdr[0] != h || hdr[1] != w || hdr[2] != ch) return 3;
        std::fwrite(hdr, 4, 3, o);
        for (size_t i = 0; i < frames.size(); i++)
            for (int y = 0; y < frames[i].rows; y++) std::fwrite(frames[i].ptr(y), 1, (size_t)frames[i].cols * frames[i].channels(), o);
        std::fclose(o);
    } catch (const Error& e) {
        std::fprintf(stderr, "jpeg_demo: %s\n", e.what());
        return 1;
    }
    return 0;
}
