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
// The pixel loop runs on the row engine of host/fhe_resize_b200.hpp (RowResizer: fused Cubic / Linear
// device graphs, fraction powers as operand forms, horizontal passes shared between output rows);
// source rows stream in from the input file and output rows stream out, so device memory holds a
// window of rows, not the image.  The ciphertexts written are the reference's bit for bit (its own
// homo/server_resize.cpp compiles unchanged against include/seal/seal.h: oracle/_ref/server_resize;
// tests compare both with the oracle).  The public fractions are encrypted with the key-less
// transparent Encryptor: this server loads neither the public key nor the secret key the reference
// reads "FOR DEBUGGING ONLY" (server_resize.cpp:99-116).
// stdout: the label, then ONE measured value per output row -- the row's device + stream time in ms
// divided by the Cubic / Linear calls the reference would have made for it (15 / 9 per pixel), repeated
// once per such call so that benchmark/analyze.py's per-label mean stays a per-call time.  The values
// are row averages, not individual call timings (the calls of a row run as five fused launches).
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
    spec.switches = {"bicubic", "verbose", "help"};
    const cli_b200::Parsed opt = cli_b200::parse(argc, argv, spec);
    if (opt.has("help")) {
        std::printf("%s [-f in] [-o out] [--bicubic] -w W -h H [--ncoeff N] [--fcoeff N] [--cmod n] [--pmod t] [--base b] "
                    "[-v]\n  ciphertext stream layout: host/seal_wire_b200.hpp (SEAL 2.3 layout from recall, UNVERIFIED against real SEAL files)\n", argv[0]);
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
        std::vector<uint64_t> host;
        RowSource source = [&](int, bool needed) {
            CiphertextBatch row;
            if (needed) row = CiphertextBatch(context, (size_t)original_width * 3, 2);
            for (int i = 0; i < original_width * 3; ++i) {
                if (!load_ciphertext(infile, context, host, hdr)) throw std::invalid_argument("ciphertext stream ended early");
                if (hdr.size != 2) throw std::invalid_argument("input ciphertexts must have size 2");
                if (needed) check(fhe_upload(context.handle(), row.element((size_t)i), host.data(), host.size() * sizeof(uint64_t)));
            }
            return row;
        };
        std::cout << (bicubic ? "Cubic," : "Linear,");
        const int inter = bicubic ? BICUBIC : BILINEAR;
        const size_t calls_per_row = (size_t)resized_width * (bicubic ? 15 : 9);
        RowResizer resizer(evaluator, encoder, encryptor, original_width, original_height, resized_width, resized_height, inter);
        auto t0 = std::chrono::steady_clock::now();
        std::vector<uint64_t> out;
        resizer.run(source, [&](int, CiphertextBatch &row) {
            context.sync();
            const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            for (size_t i = 0; i < calls_per_row; ++i) std::cout << ms / calls_per_row << ',';
            out.resize(context.ct_words(row.size()) * row.count());
            row.download(out.data());
            for (size_t i = 0; i < row.count(); ++i)
                save_ciphertext(outfile, context, out.data() + i * context.ct_words(row.size()), row.size(), hdr.hash);
            t0 = std::chrono::steady_clock::now();
        });
        std::cout << std::endl;
        return 0;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "server_resize_b200: %s\n", e.what());
        return 1;
    }
}
