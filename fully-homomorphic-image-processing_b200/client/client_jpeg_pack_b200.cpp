// client_jpeg_pack_b200 -- the last step of `client_jpeg --recieve` (homo/client_jpeg.cpp:255-291)
// as a standalone CPU tool: takes the DECRYPTED AND DECODED values of the server's output stream
// (per 8x8 tile: 64 Y, 64 Cb, 64 Cr doubles in natural order, server_jpeg.cpp:146-153), rounds them
// as the reference does, entropy-codes them into a baseline JPEG file and, optionally, prints the
// reference's RMS comparator line.  Decryption itself stays with whoever holds the secret key.
//
//   client_jpeg_pack_b200 pack <width> <height> <quality> <values.f64> <out.jpg>
//   client_jpeg_pack_b200 rms  <a.rgb> <b.rgb>            # "RMSError,<value>," (fhe_image.h:520)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../host/jpeg_entropy_b200.hpp"

// fhe_image.h:99-100
static const int YQT[64] = {16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56,
                            14, 17, 22, 29, 51, 87, 80, 62, 18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113,
                            92, 49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
static const int UVQT[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99,
                             47, 66, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
                             99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};

static std::vector<unsigned char> slurp(const char *path)
{
    std::FILE *f = std::fopen(path, "rb");
    if (!f) { std::fprintf(stderr, "cannot open %s\n", path); std::exit(2); }
    std::vector<unsigned char> v;
    unsigned char buf[1 << 16];
    size_t got;
    while ((got = std::fread(buf, 1, sizeof buf, f)) > 0) v.insert(v.end(), buf, buf + got);
    std::fclose(f);
    return v;
}

int main(int argc, char **argv)
{
    try {
        if (argc == 7 && !std::strcmp(argv[1], "pack")) {
            const int w = std::atoi(argv[2]), h = std::atoi(argv[3]), quality = std::atoi(argv[4]);
            const std::vector<unsigned char> raw = slurp(argv[5]);
            const size_t tiles = (size_t)(w / 8) * (h / 8);
            if (w <= 0 || h <= 0 || raw.size() != tiles * 192 * sizeof(double)) {
                std::fprintf(stderr, "expected %zu doubles for a %dx%d image\n", tiles * 192, w, h);
                return 2;
            }
            const double *v = reinterpret_cast<const double *>(raw.data());
            std::FILE *out = std::fopen(argv[6], "wb");
            if (!out) { std::fprintf(stderr, "cannot create %s\n", argv[6]); return 2; }
            jpeg_b200::JpegWriter jw(out, w, h, YQT, UVQT, quality);
            const uint8_t *zz = jpeg_b200::zigzag_position();
            for (size_t t = 0; t < tiles; ++t) {
                int du[3][64];
                for (int ch = 0; ch < 3; ++ch)
                    for (int j = 0; j < 64; ++j) du[ch][zz[j]] = jpeg_b200::round_coefficient(v[(t * 3 + ch) * 64 + j]);
                jw.write_mcu(du[0], du[1], du[2]);
            }
            jw.finish();
            std::fclose(out);
            return 0;
        }
        if (argc == 4 && !std::strcmp(argv[1], "rms")) {
            const std::vector<unsigned char> a = slurp(argv[2]), b = slurp(argv[3]);
            if (a.size() != b.size()) { std::fprintf(stderr, "size mismatch\n"); return 2; }
            std::printf("RMSError,%g,\n", jpeg_b200::rms_error(a.data(), b.data(), a.size()));
            return 0;
        }
    } catch (const std::exception &e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    std::fprintf(stderr, "usage: %s pack <width> <height> <quality> <values.f64> <out.jpg> | rms <a.rgb> <b.rgb>\n", argv[0]);
    return 2;
}
