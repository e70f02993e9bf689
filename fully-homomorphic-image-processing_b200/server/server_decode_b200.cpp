// server_decode_b200.cpp -- the reference's homo/server_decode.cpp re-hosted on libfhe_b200
// (SURVEY.md 8f-N3): homomorphic run-length decoding of an RLE-encoded image.
//
//   ./keys/params.txt   "W H pairs_R pairs_G pairs_B"                    server_decode.cpp:15-24
//   --file/-f, --outfile/-o, --delta, --order, --degree, --ncoeff, --fcoeff, --cmod, --pmod, --base,
//   --verbose/-v, --help                                                   server_decode.cpp:44-81
//   input: per channel `pairs` records of (element, count) ciphertexts (:127-128); output: W*H
//   pixels, 3 ciphertexts each, channel-interleaved (:139-143)
//
// One deliberate difference: the reference calls the overload of approximated_step that DECRYPTS on
// the server (server_decode.cpp:131 passes `decryptor`; fhe_decode.h:244-282).  A server must not
// hold the secret key, so this driver runs the homomorphic overload (fhe_decode.h:202-242) -- the
// one BASELINE config 5 names -- with the reference's defaults (degree 12, order 64, delta 0.5,
// server_decode.cpp:38-40).  The accumulators the reference creates with encryptor.encrypt(encode(0))
// (:121,124) and the series' starting values (fhe_decode.h:54,134) are transparent encryptions.
#include <cstdio>
#include <fstream>

#include "../host/fhe_decode_b200.hpp"
#include "../host/seal_wire_b200.hpp"
#include "cli_b200.hpp"

using namespace seal_b200;

int main(int argc, const char **argv)
{
    cli_b200::Spec spec;
    spec.alias = {{"-f", "file"}, {"-o", "outfile"}, {"-v", "verbose"}};
    spec.valued = {"file", "outfile", "delta", "order", "degree", "ncoeff", "fcoeff", "cmod", "pmod", "base", "wire"};
    spec.switches = {"verbose", "help"};
    const cli_b200::Parsed opt = cli_b200::parse(argc, argv, spec);
    if (opt.has("wire")) {  // header variant of the ciphertext records (host/seal_wire_b200.hpp)
        const std::string v = opt.str("wire", "capacity");
        if (v != "capacity" && v != "nocapacity") {
            std::cout << "error parsing options: --wire takes capacity or nocapacity" << std::endl;
            return 1;
        }
        wire_has_capacity() = v == "capacity";
    }
    if (opt.has("help")) {
        std::printf("%s [-f in] [-o out] [--delta d] [--order o] [--degree g] [--ncoeff N] [--fcoeff N] [--cmod n] [--pmod t] "
                    "[--base b] [-v]\n  ciphertext stream layout: host/seal_wire_b200.hpp (SEAL 2.3 layout from recall, UNVERIFIED against real SEAL files);\n  --wire nocapacity: records whose header has no size_capacity field\n", argv[0]);
        return 0;
    }
    int pairs[3] = {0, 0, 0}, width = 0, height = 0;
    {
        std::ifstream paramfile("./keys/params.txt");
        paramfile >> width >> height >> pairs[0] >> pairs[1] >> pairs[2];
        if (!paramfile || width <= 0 || height <= 0) {
            std::fprintf(stderr, "server_decode_b200: ./keys/params.txt missing or invalid\n");
            return 1;
        }
    }
    const std::string ctext_infile = opt.str("file", "./image/nothingpersonnel.txt");
    const std::string ctext_outfile = opt.str("outfile", "./image/zoop.txt");
    const int degree = opt.num("degree", 12), order = opt.num("order", 64);
    const double delta = opt.real("delta", 0.5);
    try {
        Context context((uint32_t)opt.num("cmod", 8192), (uint64_t)opt.num("pmod", 1 << 14));
        Evaluator evaluator(context);
        FractionalEncoder encoder(context, (uint32_t)opt.num("ncoeff", 100), (uint32_t)opt.num("fcoeff", 100), 2);
        Encryptor encryptor(context);
        std::ifstream myfile(ctext_infile.c_str(), std::ios::binary);
        std::ofstream outfile(ctext_outfile.c_str(), std::ios::binary);
        if (!myfile || !outfile) {
            std::fprintf(stderr, "server_decode_b200: cannot open %s or %s\n", ctext_infile.c_str(), ctext_outfile.c_str());
            return 1;
        }
        WireHeader hdr;
        Ciphertext c, index, elem, count;
        std::vector<std::vector<Ciphertext>> res;
        for (int i = 0; i < 3; i++) {
            encryptor.encrypt(encoder.encode(0), index);
            std::vector<Ciphertext> channel;
            for (int j = 0; j < width * height; j++) {
                encryptor.encrypt(encoder.encode(0), c);
                channel.push_back(c);
            }
            for (int j = 0; j < pairs[i]; j++) {
                std::vector<Ciphertext> run;
                if (!load_ciphertext(myfile, context, elem, hdr) || !load_ciphertext(myfile, context, count, hdr))
                    throw std::invalid_argument("ciphertext stream ended early");
                approximated_step_batched(elem, index, count, order, degree, delta, width, height, run, evaluator, encoder, encryptor);
                for (int k = 0; k < width * height; k++) evaluator.add(channel[(size_t)k], run[(size_t)k]);
                evaluator.add(index, count);
            }
            res.push_back(channel);
        }
        for (int i = 0; i < width * height; i++)
            for (int j = 0; j < 3; j++) save_ciphertext(outfile, res[(size_t)j][(size_t)i], hdr.hash);
        return 0;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "server_decode_b200: %s\n", e.what());
        return 1;
    }
}
