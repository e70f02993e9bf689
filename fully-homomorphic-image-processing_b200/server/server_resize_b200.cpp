// server_resize_b200.cpp -- the reference's homo/server_resize.cpp re-hosted on libfhe_b200
// (SURVEY.md 8f-N3): same files, flags and stdout label.
//
//   ./keys/params.txt   "W H C ..." (client_resize --send)           server_resize.cpp:15-23
//   --file/-f, --outfile/-o, --bicubic, --width/-w, --height/-h, --ncoeff, --fcoeff, --cmod, --pmod,
//   --base, --verbose/-v, --help                                       server_resize.cpp:42-87
//   input: RGB-interleaved row-major ciphertext stream (client_resize.cpp:141-154); output: the
//   resized pixels, 3 ciphertexts each (size 6 bicubic / 4 bilinear), server_resize.cpp:141-146
//   stdout: "Cubic,<ms>,<ms>,...\n" or "Linear,..." (server_resize.cpp:124-139)
//
// The op graph is ResizeImage / SampleBicubic / Cubic of homo/fhe_resize.h op for op, evaluated a
// ROW of output pixels at a time on the device (ResizeImageBatched); `--per-pixel` runs the
// reference's pixel-at-a-time order instead (same ciphertexts).  The public fractions are
// encrypted with the key-less transparent Encryptor: this server loads neither the public key nor
// the secret key the reference reads "FOR DEBUGGING ONLY" (server_resize.cpp:99-116).
#include <chrono>
#include <cstdio>
#include <fstream>
#include <sstream>

#include "../host/fhe_resize_b200.hpp"
#include "../host/seal_wire_b200.hpp"
#include "cli_b200.hpp"

using namespace seal_b200;

int main(int argc, const char **argv)
{
    cli_b200::Spec spec;
    spec.alias = {{"-f", "file"}, {"-o", "outfile"}, {"-v", "verbose"}, {"-w", "width"}, {"-h", "height"}};
    spec.valued = {"file", "outfile", "ncoeff", "fcoeff", "cmod", "pmod", "width", "height", "base"};
    spec.switches = {"bicubic", "verbose", "help", "per-pixel"};
    const cli_b200::Parsed opt = cli_b200::parse(argc, argv, spec);
    if (opt.has("help")) {
        std::printf("%s [-f in] [-o out] [--bicubic] -w W -h H [--ncoeff N] [--fcoeff N] [--cmod n] [--pmod t] [--base b] "
                    "[--per-pixel] [-v]\n", argv[0]);
        return 0;
    }
    int original_width = 0, original_height = 0, channels = 0;
    {
        std::ifstream paramfile("./keys/params.txt");
        paramfile >> original_width >> original_height >> channels;
        if (!paramfile || channels != 3 || original_width == 0 || original_height == 0) {
            std::fprintf(stderr, "server_resize_b200: ./keys/params.txt missing or invalid\n");
            return 1;
        }
    }
    if (!opt.has("width") || !opt.has("height")) {
        std::cout << "Please enter the width/height of the resized image..." << std::endl;
        std::cout << "\t Use --help to see the options." << std::endl;
        return 1;
    }
    const std::string ctext_infile = opt.str("file", "./image/nothingpersonnel.txt");
    const std::string ctext_outfile = opt.str("outfile", "./image/zoop.txt");
    const bool bicubic = opt.has("bicubic");
    const int resized_width = opt.num("width", 0), resized_height = opt.num("height", 0);
    try {
        Context context((uint32_t)opt.num("cmod", 8192), (uint64_t)opt.num("pmod", 1 << 14));
        Evaluator evaluator(context);
        FractionalEncoder encoder(context, (uint32_t)opt.num("ncoeff", 100), (uint32_t)opt.num("fcoeff", 100), 2);
        Encryptor encryptor(context);
        std::ifstream infile(ctext_infile.c_str(), std::ios::binary);
        std::ofstream outfile(ctext_outfile.c_str(), std::ios::binary);
        if (!infile || !outfile) {
            std::fprintf(stderr, "server_resize_b200: cannot open %s or %s\n", ctext_infile.c_str(), ctext_outfile.c_str());
            return 1;
        }
        WireHeader hdr;
        CiphertextSource source = [&](Ciphertext &dst) {
            if (!load_ciphertext(infile, context, dst, hdr)) throw std::invalid_argument("ciphertext stream ended early");
        };
        std::cout << (bicubic ? "Cubic," : "Linear,");
        SImageData resize_im;
        const int inter = bicubic ? BICUBIC : BILINEAR;
        if (opt.has("per-pixel")) {
            ResizeImage(source, original_width, original_height, resize_im, resized_width, resized_height, inter, evaluator,
                        encoder, encryptor);  // Cubic / Linear print their own timings, like the reference
        } else {
            const auto t0 = std::chrono::steady_clock::now();
            ResizeImageBatched(source, original_width, original_height, resize_im, resized_width, resized_height, inter,
                               evaluator, encoder, encryptor);
            context.sync();
            const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            const size_t calls = (size_t)resized_width * resized_height * (bicubic ? 15 : 9);
            for (size_t i = 0; i < calls; ++i) std::cout << ms / calls << ',';
        }
        std::cout << std::endl;
        for (size_t i = 0; i < resize_im.pixels.size(); i++)
            for (int j = 0; j < 3; j++) save_ciphertext(outfile, resize_im.pixels[i][(size_t)j], hdr.hash);
        return 0;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "server_resize_b200: %s\n", e.what());
        return 1;
    }
}
