// ref_jpeg_harness.cpp -- TEST INFRASTRUCTURE.  A three-command driver around the two image helpers
// the reference vendors and uses in its JPEG comparison (homo/fhe_image.h:14-15, :508-521,
// homo/client_jpeg.cpp:76,295): stb_image's decoder and jo_jpeg's encoder, compiled from the
// sources where they lie (/root/reference/include/stb_image.h, jo_jpeg.h; recipe: oracle/Makefile
// target `ref`, output oracle/_ref/).  Only this file is ours; it is used by
// tests/golden/make_golden_jpeg.py to produce the committed fixtures and by the CPU tests (when
// built) to decode with the very decoder the reference's RMSError was measured with.
//
//   ref_jpeg_harness decode <in.jpg> <out.rgb>              -> prints "<width> <height>"
//   ref_jpeg_harness jo <in.rgb> <width> <height> <quality> <out.jpg>
//   ref_jpeg_harness rms <a.rgb> <b.rgb>                    -> compare_jpeg_jojpeg's arithmetic (int accumulator)
#define STB_IMAGE_IMPLEMENTATION
#include "stb_image.h"
#include "jo_jpeg.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

static std::vector<unsigned char> slurp(const char *p)
{
    std::vector<unsigned char> v;
    FILE *f = fopen(p, "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", p); exit(2); }
    unsigned char buf[65536];
    size_t g;
    while ((g = fread(buf, 1, sizeof buf, f)) > 0) v.insert(v.end(), buf, buf + g);
    fclose(f);
    return v;
}

int main(int argc, char **argv)
{
    if (argc == 4 && !strcmp(argv[1], "decode")) {
        int w, h, comp;
        unsigned char *d = stbi_load(argv[2], &w, &h, &comp, 3);   // client_jpeg.cpp:76 / fhe_image.h:510
        if (!d) { fprintf(stderr, "stbi_load failed\n"); return 1; }
        FILE *f = fopen(argv[3], "wb");
        fwrite(d, 1, (size_t)w * h * 3, f);
        fclose(f);
        printf("%d %d\n", w, h);
        return 0;
    }
    if (argc == 7 && !strcmp(argv[1], "jo")) {
        std::vector<unsigned char> rgb = slurp(argv[2]);
        const int w = atoi(argv[3]), h = atoi(argv[4]), q = atoi(argv[5]);
        if ((size_t)w * h * 3 != rgb.size()) { fprintf(stderr, "size mismatch\n"); return 2; }
        return jo_write_jpg(argv[6], rgb.data(), w, h, 3, q) ? 0 : 1;   // fhe_image.h:511
    }
    if (argc == 4 && !strcmp(argv[1], "rms")) {
        std::vector<unsigned char> a = slurp(argv[2]), b = slurp(argv[3]);
        if (a.size() != b.size()) return 2;
        int running_error = 0;                                            // fhe_image.h:514-519, as written
        for (size_t i = 0; i < a.size(); i++) running_error += (a[i] - b[i]) * (a[i] - b[i]);
        printf("RMSError,%g,\n", std::sqrt(((double)running_error) / (double)a.size()));
        return 0;
    }
    fprintf(stderr, "usage: decode <in.jpg> <out.rgb> | jo <in.rgb> <w> <h> <quality> <out.jpg> | rms <a.rgb> <b.rgb>\n");
    return 2;
}
