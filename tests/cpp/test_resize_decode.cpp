// test_resize_decode.cpp -- the C++ host layer for the resize and decode op libraries
// (host/fhe_resize_b200.hpp, host/fhe_decode_b200.hpp) on a real B200.
//
// Usage: test_resize_decode <n> <in.bin> <out.bin> [<evk.bin> <dbc> [<ncoeff> <fcoeff>]]
// in.bin (raw u64, size-2 ciphertexts of W = 2*k*n words, written by tests/test_host_cpp.py from
// the oracle's generator / encryptor):
//   [0..4]   A B C D t            -> Cubic level 1 (size 4), Cubic level 2 on four copies (size 6),
//                                    Linear(A, B, t) (size 3)
//   [5..7]   amplitude index count
//   [8..15]  8 fresh "encryptions of zero" consumed by sin/cos inside approximated_step
//            (order 8, degree 2, delta 0.25, width*height 2); [8],[9] also feed the stand-alone
//            homomorphic_sin(A) / homomorphic_cos(B)
//   [16..33] 18 encrypted fractions, consumed in call order by SampleBicubic (2 per output pixel)
//   [34..]   5x6 RGB source image, RGB-interleaved row-major (90 ciphertexts)
// out.bin: the results in the order above, then the 3x3 bicubic ResizeImage output (9 pixels x 3
// channels, size 6).  The pytest wrapper compares it with the oracle bit for bit.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>

#include "../../fully-homomorphic-image-processing_b200/host/fhe_decode_b200.hpp"
#include "../../fully-homomorphic-image-processing_b200/host/fhe_resize_b200.hpp"

using namespace seal_b200;

#define REQUIRE(cond)                                                            \
    do {                                                                         \
        if (!(cond)) {                                                           \
            std::fprintf(stderr, "FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            return 1;                                                            \
        }                                                                        \
    } while (0)

int main(int argc, char **argv)
{
    try {
        const uint32_t n = argc > 1 ? (uint32_t)std::atoi(argv[1]) : 2048;
        Context ctx(n);
        if (argc < 4) {
            std::fprintf(stderr, "usage: %s n in.bin out.bin\n", argv[0]);
            return 2;
        }
        const size_t W = ctx.ct_words(2);
        std::ifstream in(argv[2], std::ios::binary);
        REQUIRE((bool)in);
        in.seekg(0, std::ios::end);
        const size_t bytes = (size_t)in.tellg();
        in.seekg(0);
        std::vector<uint64_t> words(bytes / 8);
        in.read((char *)words.data(), (std::streamsize)bytes);
        REQUIRE(words.size() >= (34 + 90) * W);
        auto ct_at = [&](size_t i) {
            Ciphertext c(ctx, 2);
            c.upload(words.data() + i * W);
            return c;
        };
        std::ofstream out(argv[3], std::ios::binary);
        auto dump = [&](const Ciphertext &c) {
            std::vector<uint64_t> h(ctx.ct_words(c.size()));
            c.download(h.data());
            out.write((const char *)h.data(), (std::streamsize)(h.size() * 8));
        };
        // the functions print timings like the reference; keep stdout clean (restored on every path)
        struct Quiet {
            std::ostringstream sink;
            std::streambuf *old;
            Quiet() : old(std::cout.rdbuf(sink.rdbuf())) {}
            ~Quiet() { std::cout.rdbuf(old); }
        } quiet;

        Evaluator evaluator(ctx);
        FractionalEncoder encoder(ctx, argc > 7 ? (uint32_t)std::atoi(argv[6]) : 64, argc > 7 ? (uint32_t)std::atoi(argv[7]) : 32);
        size_t next_zero = 8, next_frac = 16;
        Encryptor zeros([&](const Plaintext &, Ciphertext &dst) { dst = ct_at(next_zero++); });
        Encryptor fracs([&](const Plaintext &, Ciphertext &dst) { dst = ct_at(next_frac++); });

        Ciphertext A = ct_at(0), B = ct_at(1), C = ct_at(2), D = ct_at(3), t = ct_at(4);
        Ciphertext c1, c2, lin;
        Cubic(c1, A, B, C, D, t, evaluator, encoder, zeros);
        REQUIRE(c1.size() == 4);
        dump(c1);
        Ciphertext e1(c1), e2(c1), e3(c1), e4(c1);
        Cubic(c2, e1, e2, e3, e4, t, evaluator, encoder, zeros);
        REQUIRE(c2.size() == 6);
        dump(c2);
        Linear(lin, A, B, t, evaluator, encoder, zeros);
        REQUIRE(lin.size() == 3);
        dump(lin);

        Ciphertext s, c;
        homomorphic_sin(A, s, evaluator, encoder, zeros);
        homomorphic_cos(B, c, evaluator, encoder, zeros);
        REQUIRE(s.size() == 11 && c.size() == 11 && next_zero == 10);
        dump(s);
        dump(c);

        next_zero = 8;
        Ciphertext amplitude = ct_at(5), index = ct_at(6), count = ct_at(7);
        std::vector<Ciphertext> run;
        approximated_step(amplitude, index, count, 8, 2, 0.25, 2, 1, run, evaluator, encoder, zeros);
        REQUIRE(run.size() == 2 && run[0].size() == 22 && next_zero == 16);
        dump(run[0]);
        dump(run[1]);

        size_t next_src = 34;
        SImageData dest;
        ResizeImage([&](Ciphertext &dst) { dst = ct_at(next_src++); }, 5, 6, dest, 3, 3, BICUBIC, evaluator, encoder, fracs);
        REQUIRE(dest.pixels.size() == 9 && next_frac == 34 && next_src == 34 + 90);
        for (auto &px : dest.pixels)
            for (auto &ch : px) {
                REQUIRE(ch.size() == 6);
                dump(ch);
            }

        // the row-batched sampler (one Evaluator call per graph op for a whole output row) gives the
        // same ciphertexts as the reference-style per-pixel sampler, bicubic and bilinear
        auto same = [&](const Ciphertext &x, const Ciphertext &y) {
            if (x.size() != y.size()) return false;
            std::vector<uint64_t> hx(ctx.ct_words(x.size())), hy(hx.size());
            x.download(hx.data());
            y.download(hy.data());
            return hx == hy;
        };
        for (int inter : {BICUBIC, BILINEAR}) {
            size_t src1 = 34, src2 = 34, f1 = 16, f2 = 16;
            Encryptor fr1([&](const Plaintext &, Ciphertext &dst) { dst = ct_at(f1++); });
            Encryptor fr2([&](const Plaintext &, Ciphertext &dst) { dst = ct_at(f2++); });
            SImageData d1, d2;
            ResizeImage([&](Ciphertext &dst) { dst = ct_at(src1++); }, 5, 6, d1, 3, 3, inter, evaluator, encoder, fr1);
            ResizeImageBatched([&](Ciphertext &dst) { dst = ct_at(src2++); }, 5, 6, d2, 3, 3, inter, evaluator, encoder, fr2);
            REQUIRE(d1.pixels.size() == 9 && d2.pixels.size() == 9 && f1 == f2 && src1 == src2);
            for (size_t i = 0; i < 9; ++i)
                for (size_t ch = 0; ch < 3; ++ch) {
                    REQUIRE(d1.pixels[i][ch].size() == (inter == BICUBIC ? 6u : 4u));
                    REQUIRE(same(d1.pixels[i][ch], d2.pixels[i][ch]));
                }
        }
        // batched sine == per-ciphertext sine
        {
            CiphertextBatch xs(ctx, 3, 2), rs;
            for (size_t i = 0; i < 3; ++i) xs.set(i, ct_at(i));
            next_zero = 8;
            homomorphic_sin(xs, rs, evaluator, encoder, zeros);
            REQUIRE(rs.count() == 3 && rs.size() == 11 && next_zero == 11);
            for (size_t i = 0; i < 3; ++i) {
                Ciphertext xi = ct_at(i), ri;
                next_zero = 8 + i;
                homomorphic_sin(xi, ri, evaluator, encoder, zeros);
                REQUIRE(same(ri, rs.get(i)));
            }
        }

        // batched step == the reference-order step, with the same stream of server-side encryptions
        {
            next_zero = 8;
            Ciphertext amp2 = ct_at(5), idx2 = ct_at(6), cnt2 = ct_at(7);
            std::vector<Ciphertext> run2;
            approximated_step_batched(amp2, idx2, cnt2, 8, 2, 0.25, 2, 1, run2, evaluator, encoder, zeros);
            REQUIRE(run2.size() == 2 && next_zero == 16);
            REQUIRE(same(run[0], run2[0]) && same(run[1], run2[1]));
        }

        // the key-less deterministic Encryptor: (Delta*m, 0) decrypts to m and is accepted by the graphs
        Encryptor transparent(ctx);
        Ciphertext tz;
        transparent.encrypt(encoder.encode(0.375), tz);
        REQUIRE(tz.size() == 2);
        dump(tz);
        // a default-constructed Encryptor has no key: std::logic_error
        bool threw = false;
        try {
            Encryptor none;
            none.encrypt(encoder.encode(1.0), tz);
        } catch (const std::logic_error &) {
            threw = true;
        }
        REQUIRE(threw);
        // relinearisation policy (Evaluator::set_relinearization, an extension): every product is
        // brought back to size 2; argv[4] = evaluation keys (coefficient form), argv[5] = dbc
        if (argc > 5) {
            std::ifstream kin(argv[4], std::ios::binary);
            REQUIRE((bool)kin);
            kin.seekg(0, std::ios::end);
            std::vector<uint64_t> kw((size_t)kin.tellg() / 8);
            kin.seekg(0);
            kin.read((char *)kw.data(), (std::streamsize)(kw.size() * 8));
            const uint32_t dbc = (uint32_t)std::atoi(argv[5]);
            EvaluationKeys evk(ctx, kw.data(), dbc, 1);
            evaluator.set_relinearization(&evk);
            Ciphertext s2, lin2, Ar = ct_at(0), Br = ct_at(1), tr = ct_at(4);
            next_zero = 8;
            homomorphic_sin(Ar, s2, evaluator, encoder, zeros);
            REQUIRE(s2.size() == 2);
            dump(s2);
            Linear(lin2, Ar, Br, tr, evaluator, encoder, zeros);
            REQUIRE(lin2.size() == 2);
            dump(lin2);
            Ciphertext cub2, Cr = ct_at(2), Dr = ct_at(3);
            Cubic(cub2, Ar, Br, Cr, Dr, tr, evaluator, encoder, zeros);
            REQUIRE(cub2.size() == 2);
            dump(cub2);
            // the whole step with every product relinearised, reference order and batched
            {
                next_zero = 8;
                Ciphertext amp3 = ct_at(5), idx3 = ct_at(6), cnt3 = ct_at(7);
                std::vector<Ciphertext> run3, run4;
                approximated_step(amp3, idx3, cnt3, 8, 2, 0.25, 2, 1, run3, evaluator, encoder, zeros);
                REQUIRE(run3.size() == 2 && run3[0].size() == 2 && next_zero == 16);
                dump(run3[0]);
                dump(run3[1]);
                next_zero = 8;
                Ciphertext amp4 = ct_at(5), idx4 = ct_at(6), cnt4 = ct_at(7);
                approximated_step_batched(amp4, idx4, cnt4, 8, 2, 0.25, 2, 1, run4, evaluator, encoder, zeros);
                REQUIRE(run4.size() == 2 && same(run3[0], run4[0]) && same(run3[1], run4[1]));
            }
            // batched == single under the same policy
            CiphertextBatch xs(ctx, 2, 2), rs;
            xs.set(0, ct_at(0));
            xs.set(1, ct_at(1));
            next_zero = 8;
            homomorphic_sin(xs, rs, evaluator, encoder, zeros);
            REQUIRE(rs.size() == 2 && same(rs.get(0), s2));
            evaluator.set_relinearization(nullptr);
            Ciphertext s3, Ar2 = ct_at(0);
            next_zero = 8;
            homomorphic_sin(Ar2, s3, evaluator, encoder, zeros);
            REQUIRE(s3.size() == 11);
        }
        std::printf("PASS\n");
        return 0;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "exception: %s\n", e.what());
        return 1;
    }
}
