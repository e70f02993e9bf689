// ref_graphs.cpp -- TEST INFRASTRUCTURE.  Runs the REFERENCE'S OWN op graphs -- homo/fhe_image.h,
// homo/fhe_resize.h and homo/fhe_decode.h included unchanged from /root/reference (one token of
// fhe_decode.h patched at build time, see oracle/Makefile) -- on the GPU through include/seal/seal.h,
// and dumps every result so that tests/test_reference_sources_on_gpu.py can compare them bit for bit
// with the oracle's restatement of the same graphs.  Only this driver is ours.  Output: oracle/_ref/.
//
//   ref_graphs jpeg   <n> <ncoeff> <fcoeff> <in.bin> <out.bin>
//       in : 192 size-2 ciphertexts R64 G64 B64 (server_jpeg.cpp:115-124), raw u64 [poly][limb][n]
//       out: rgb_to_ycc_fhe per pixel, encrypted_dct per channel (server_jpeg.cpp:126-136), then
//            quantize_fhe with YQT / UVQT / UVQT -- 192 ciphertexts
//   ref_graphs resize <n> <ncoeff> <fcoeff> <in.bin> <out.bin> <stream.wire> <ow> <oh> <dw> <dh> [<evk.bin> <dbc>]
//       in : A B C D t; out: Cubic level 1, Cubic level 2 (four copies of level 1), Linear, then the
//            bicubic ResizeImage of the wire-format stream (fhe_resize.h:308-392 reads the file itself);
//            with keys: Cubic and Linear again with every product relinearised
//   ref_graphs decode <n> <ncoeff> <fcoeff> <in.bin> <out.bin> <order> <degree> <delta> <npos> [<evk.bin> <dbc>]
//       in : x_sin x_cos amplitude index count; out: homomorphic_sin, homomorphic_cos,
//            approximated_step (npos x 1); with keys: the same three again with every product
//            relinearised (Evaluator::set_relinearization)
// Server-side encryptions (fhe_resize.h:230-266, fhe_decode.h:54,134) are the transparent ones of seal.h.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sstream>

#include "fhe_image.h"
#include "fhe_resize.h"
#include "fhe_decode.h"

using namespace seal;

static std::vector<uint64_t> slurp(const char *path)
{
    std::ifstream in(path, std::ios::binary);
    if (!in) throw std::runtime_error(std::string("cannot open ") + path);
    in.seekg(0, std::ios::end);
    std::vector<uint64_t> w((size_t)in.tellg() / 8);
    in.seekg(0);
    in.read((char *)w.data(), (std::streamsize)(w.size() * 8));
    return w;
}

int main(int argc, char **argv)
{
    if (argc < 7) {
        std::fprintf(stderr, "usage: %s jpeg|resize|decode n ncoeff fcoeff in.bin out.bin ...\n", argv[0]);
        return 2;
    }
    try {
        const std::string mode = argv[1];
        const int n = std::atoi(argv[2]), ncoeff = std::atoi(argv[3]), fcoeff = std::atoi(argv[4]);
        // server_jpeg.cpp:74-100
        EncryptionParameters params;
        char poly_mod[32];
        std::snprintf(poly_mod, sizeof poly_mod, "1x^%i + 1", n);
        params.set_poly_modulus(poly_mod);
        params.set_coeff_modulus(coeff_modulus_128(n));
        params.set_plain_modulus(PLAIN_MODULUS);
        SEALContext context(params);
        PublicKey public_key;
        Encryptor encryptor(context, public_key);
        Evaluator evaluator(context);
        FractionalEncoder encoder(context.plain_modulus(), context.poly_modulus(), ncoeff, fcoeff, POLY_BASE);

        const seal_b200::Context &dc = context.device_context();
        const size_t W = dc.ct_words(2);
        const std::vector<uint64_t> words = slurp(argv[5]);
        auto ct_at = [&](size_t i) {
            if ((i + 1) * W > words.size()) throw std::runtime_error("input file too short");
            Ciphertext c(dc, 2);
            c.upload(words.data() + i * W);
            return c;
        };
        std::ofstream out(argv[6], std::ios::binary);
        auto dump = [&](const Ciphertext &c) {
            std::vector<uint64_t> h(dc.ct_words(c.size()));
            c.download(h.data());
            out.write((const char *)h.data(), (std::streamsize)(h.size() * 8));
        };
        // the reference's functions print their timings to stdout
        std::ostringstream sink;
        std::streambuf *old = std::cout.rdbuf(sink.rdbuf());

        if (mode == "jpeg") {
            std::vector<Ciphertext> red, green, blue;
            for (int i = 0; i < 64; ++i) red.push_back(ct_at(i));
            for (int i = 0; i < 64; ++i) green.push_back(ct_at(64 + i));
            for (int i = 0; i < 64; ++i) blue.push_back(ct_at(128 + i));
            for (size_t i = 0; i < red.size(); i++) rgb_to_ycc_fhe(red[i], green[i], blue[i], evaluator, encoder, encryptor);
            encrypted_dct(red, evaluator, encoder, encryptor);
            encrypted_dct(green, evaluator, encoder, encryptor);
            encrypted_dct(blue, evaluator, encoder, encryptor);
            const std::vector<double> yq(YQT, YQT + 64), uvq(UVQT, UVQT + 64);
            quantize_fhe(red, yq, evaluator, encoder, encryptor);
            quantize_fhe(green, uvq, evaluator, encoder, encryptor);
            quantize_fhe(blue, uvq, evaluator, encoder, encryptor);
            for (auto &c : red) dump(c);
            for (auto &c : green) dump(c);
            for (auto &c : blue) dump(c);
        } else if (mode == "resize") {
            if (argc < 12) throw std::runtime_error("resize: missing arguments");
            Ciphertext A = ct_at(0), B = ct_at(1), C = ct_at(2), D = ct_at(3), t = ct_at(4);
            Ciphertext c1, c2, lin;
            Cubic(c1, A, B, C, D, t, evaluator, encoder, encryptor);
            dump(c1);
            Ciphertext e1(c1), e2(c1), e3(c1), e4(c1);
            Cubic(c2, e1, e2, e3, e4, t, evaluator, encoder, encryptor);
            dump(c2);
            Linear(lin, A, B, t, evaluator, encoder, encryptor);
            dump(lin);
            SImageData dest;
            ResizeImage(std::string(argv[7]), std::atoi(argv[8]), std::atoi(argv[9]), dest, std::atoi(argv[10]),
                        std::atoi(argv[11]), BICUBIC, evaluator, encoder, encryptor);
            for (auto &px : dest.pixels)
                for (auto &ch : px) dump(ch);
            if (argc > 13) {
                const std::vector<uint64_t> kw = slurp(argv[12]);
                EvaluationKeys evk;
                evk.assign(context, kw.data(), std::atoi(argv[13]), 1);
                evaluator.set_relinearization(&evk.device(dc));
                Ciphertext cr, lr;
                Cubic(cr, A, B, C, D, t, evaluator, encoder, encryptor);
                dump(cr);
                Linear(lr, A, B, t, evaluator, encoder, encryptor);
                dump(lr);
                evaluator.set_relinearization(nullptr);
            }
        } else if (mode == "decode") {
            if (argc < 11) throw std::runtime_error("decode: missing arguments");
            const int order = std::atoi(argv[7]), degree = std::atoi(argv[8]), npos = std::atoi(argv[10]);
            const double delta = std::atof(argv[9]);
            EvaluationKeys evk;
            const bool with_keys = argc > 12;
            if (with_keys) {
                const std::vector<uint64_t> kw = slurp(argv[11]);
                evk.assign(context, kw.data(), std::atoi(argv[12]), 1);
            }
            for (int pass = 0; pass < (with_keys ? 2 : 1); ++pass) {
                evaluator.set_relinearization(pass ? &evk.device(dc) : nullptr);
                Ciphertext xs = ct_at(0), xc = ct_at(1), s, c;
                homomorphic_sin(xs, s, evaluator, encoder, encryptor);
                homomorphic_cos(xc, c, evaluator, encoder, encryptor);
                dump(s);
                dump(c);
                Ciphertext amplitude = ct_at(2), index = ct_at(3), count = ct_at(4);
                std::vector<Ciphertext> run;
                approximated_step(amplitude, index, count, order, degree, delta, npos, 1, run, evaluator, encoder, encryptor);
                for (auto &r : run) dump(r);
            }
        } else {
            throw std::runtime_error("unknown mode " + mode);
        }
        std::cout.rdbuf(old);
        out.close();
        std::printf("PASS\n");
        return 0;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "exception: %s\n", e.what());
        return 1;
    }
}
