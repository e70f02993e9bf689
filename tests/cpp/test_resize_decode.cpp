// test_resize_decode.cpp -- the B200-first host layer for the resize and decode servers
// (host/fhe_resize_b200.hpp: RowResizer over the fused Cubic / Linear device graphs;
// host/fhe_decode_b200.hpp: batched Taylor series and step approximation over operand forms) on a
// real B200.  (The reference's own per-ciphertext graphs are exercised from its unchanged headers:
// oracle/ref_graphs.cpp.)
//
// Usage: test_resize_decode <n> <in.bin> <out.bin> [<evk.bin> <dbc> [<ncoeff> <fcoeff>]]
// in.bin (raw u64, size-2 ciphertexts of W = 2*k*n words, written by tests/test_host_resize_decode.py
// from the oracle's generator / encryptor):
//   [0..4]   A B C D t            -> Cubic level 1 (size 4), Cubic level 2 on four copies (size 6),
//                                    Linear(A, B, t) (size 3)
//   [5..7]   amplitude index count
//   [8..15]  8 fresh "encryptions of zero" consumed by the series inside the step approximation
//            (order 8, degree 2, delta 0.25, width*height 2); [8],[9] also feed the stand-alone
//            sine of A / cosine of B
//   [16..33] 18 encrypted fractions, consumed in the reference's call order (x then y per output pixel)
//   [34..]   5x6 RGB source image, RGB-interleaved row-major (90 ciphertexts)
// out.bin: the results in the order above, then the 3x3 bicubic resize (9 pixels x 3 channels, size
// 6), a transparent encryption, and with keys: the sine and the whole step with every product
// relinearised.  The pytest wrapper compares it with the oracle bit for bit.
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
            std::fprintf(stderr, "usage: %s n in.bin out.bin [evk.bin dbc [ncoeff fcoeff]]\n", argv[0]);
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
        auto batch_of = [&](std::initializer_list<size_t> idx) {
            CiphertextBatch b(ctx, idx.size(), 2);
            size_t e = 0;
            for (size_t i : idx) b.set(e++, ct_at(i));
            return b;
        };
        std::ofstream out(argv[3], std::ios::binary);
        auto dump_words = [&](const uint64_t *d, size_t count) { out.write((const char *)d, (std::streamsize)(count * 8)); };
        auto dump = [&](const Ciphertext &c) {
            std::vector<uint64_t> h(ctx.ct_words(c.size()));
            c.download(h.data());
            dump_words(h.data(), h.size());
        };
        auto dump_batch = [&](const CiphertextBatch &b) {
            std::vector<uint64_t> h(ctx.ct_words(b.size()) * b.count());
            b.download(h.data());
            dump_words(h.data(), h.size());
        };
        auto same = [&](const Ciphertext &x, const Ciphertext &y) {
            if (x.size() != y.size()) return false;
            std::vector<uint64_t> hx(ctx.ct_words(x.size())), hy(hx.size());
            x.download(hx.data());
            y.download(hy.data());
            return hx == hy;
        };

        Evaluator evaluator(ctx);
        FractionalEncoder encoder(ctx, argc > 7 ? (uint32_t)std::atoi(argv[6]) : 64, argc > 7 ? (uint32_t)std::atoi(argv[7]) : 32);
        size_t next_zero = 8, next_frac = 16;
        Encryptor zeros([&](const Plaintext &, Ciphertext &dst) { dst = ct_at(next_zero++); });
        Encryptor fracs([&](const Plaintext &, Ciphertext &dst) { dst = ct_at(next_frac++); });
        REQUIRE(!zeros.deterministic());

        // ---- the fused graphs on single calls: Cubic level 1 and 2, Linear
        {
            ResizePlan plan(ctx, encoder);
            CiphertextBatch t = batch_of({4}), abcd = batch_of({0, 1, 2, 3});
            OperandForm ft(ctx, t);
            CiphertextBatch t2 = evaluator.multiply(ft, ft);  // square(t) == multiply(t, t)
            REQUIRE(t2.size() == 3);
            OperandForm ft2(ctx, t2);
            const uint64_t *taps[4] = {abcd.element(0), abcd.element(1), abcd.element(2), abcd.element(3)};
            CiphertextBatch c1(ctx, 1, 4), c2(ctx, 1, 6), lin(ctx, 1, 3);
            check(fhe_cubic(ctx.handle(), plan.handle(), taps, 2, ft.handle(), ft2.handle(), ft2.handle(), 1, c1.device_data(), 1));
            dump_batch(c1);
            const uint64_t *taps2[4] = {c1.element(0), c1.element(0), c1.element(0), c1.element(0)};
            check(fhe_cubic(ctx.handle(), plan.handle(), taps2, 4, ft.handle(), ft2.handle(), ft2.handle(), 1, c2.device_data(), 1));
            dump_batch(c2);
            CiphertextBatch one_minus_t(t);
            evaluator.negate(one_minus_t);
            evaluator.add_plain(one_minus_t, encoder.encode(1.0));
            OperandForm f1mt(ctx, one_minus_t);
            check(fhe_linear(ctx.handle(), taps, 2, f1mt.handle(), ft.handle(), 1, lin.device_data(), 1));
            dump_batch(lin);
        }

        // ---- Taylor series and the step approximation
        CiphertextBatch s, c;
        {
            CiphertextBatch xa = batch_of({0}), xb = batch_of({1});
            homomorphic_sin(xa, s, evaluator, encoder, zeros);
            homomorphic_cos(xb, c, evaluator, encoder, zeros);
            REQUIRE(s.size() == 11 && c.size() == 11 && next_zero == 10);
            dump_batch(s);
            dump_batch(c);
        }
        std::vector<Ciphertext> run;
        {
            next_zero = 8;
            Ciphertext amplitude = ct_at(5), index = ct_at(6), count = ct_at(7);
            approximated_step_batched(amplitude, index, count, 8, 2, 0.25, 2, 1, run, evaluator, encoder, zeros);
            REQUIRE(run.size() == 2 && run[0].size() == 22 && next_zero == 16);
            dump(run[0]);
            dump(run[1]);
        }
        // a batch of three arguments == three single series
        {
            CiphertextBatch xs = batch_of({0, 1, 2}), rs;
            next_zero = 8;
            homomorphic_sin(xs, rs, evaluator, encoder, zeros);
            REQUIRE(rs.count() == 3 && rs.size() == 11 && next_zero == 11);
            REQUIRE(same(rs.get(0), s.get(0)));
            for (size_t i = 1; i < 3; ++i) {
                CiphertextBatch xi = batch_of({i}), ri;
                next_zero = 8 + i;
                homomorphic_sin(xi, ri, evaluator, encoder, zeros);
                REQUIRE(same(ri.get(0), rs.get(i)));
            }
        }

        // ---- the row engine: bicubic 5x6 -> 3x3 with the reference's stream of server-side encryptions
        size_t next_src = 34;
        SImageData dest;
        ResizeImageBatched([&](Ciphertext &dst) { dst = ct_at(next_src++); }, 5, 6, dest, 3, 3, BICUBIC, evaluator, encoder, fracs);
        REQUIRE(dest.pixels.size() == 9 && next_frac == 34 && next_src == 34 + 90);
        for (auto &px : dest.pixels)
            for (auto &ch : px) {
                REQUIRE(ch.size() == 6);
                dump(ch);
            }
        // The deterministic (transparent) Encryptor lets the engine share fractions, their powers and the
        // horizontal passes between rows; a callback Encryptor producing the SAME transparent ciphertexts
        // shares nothing: identical images, bicubic and bilinear, and a band equals the rows of the whole
        Encryptor transparent(ctx);
        REQUIRE(transparent.deterministic());
        Encryptor opaque([&](const Plaintext &p, Ciphertext &dst) { transparent.encrypt(p, dst); });
        for (int inter : {BICUBIC, BILINEAR}) {
            size_t s1 = 34, s2 = 34;
            SImageData d1, d2;
            ResizeImageBatched([&](Ciphertext &dst) { dst = ct_at(s1++); }, 5, 6, d1, 4, 5, inter, evaluator, encoder, transparent);
            ResizeImageBatched([&](Ciphertext &dst) { dst = ct_at(s2++); }, 5, 6, d2, 4, 5, inter, evaluator, encoder, opaque);
            REQUIRE(d1.pixels.size() == 20 && d2.pixels.size() == 20);
            for (size_t i = 0; i < 20; ++i)
                for (size_t ch = 0; ch < 3; ++ch) {
                    REQUIRE(d1.pixels[i][ch].size() == (inter == BICUBIC ? 6u : 4u));
                    REQUIRE(same(d1.pixels[i][ch], d2.pixels[i][ch]));
                }
            // output rows [2, 4) as a band (a rank of a multi-GPU run), fed by row index
            RowResizer band(evaluator, encoder, transparent, 5, 6, 4, 5, inter);
            size_t got_rows = 0;
            band.run(
                [&](int y, bool needed) {
                    CiphertextBatch row;
                    if (needed) {
                        row = CiphertextBatch(ctx, 15, 2);
                        for (size_t i = 0; i < 15; ++i) row.set(i, ct_at(34 + (size_t)y * 15 + i));
                    }
                    return row;
                },
                [&](int y, CiphertextBatch &row) {
                    for (size_t x = 0; x < 4; ++x)
                        for (size_t ch = 0; ch < 3; ++ch)
                            if (!same(row.get(x * 3 + ch), d1.pixels[(size_t)y * 4 + x][ch])) throw std::runtime_error("band row differs");
                    ++got_rows;
                },
                2, 4);
            REQUIRE(got_rows == 2);
        }

        // the key-less deterministic Encryptor: (Delta*m, 0) decrypts to m and is accepted by the graphs
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
            CiphertextBatch xa = batch_of({0}), s2;
            next_zero = 8;
            homomorphic_sin(xa, s2, evaluator, encoder, zeros);
            REQUIRE(s2.size() == 2);
            dump_batch(s2);
            next_zero = 8;
            Ciphertext amp3 = ct_at(5), idx3 = ct_at(6), cnt3 = ct_at(7);
            std::vector<Ciphertext> run3;
            approximated_step_batched(amp3, idx3, cnt3, 8, 2, 0.25, 2, 1, run3, evaluator, encoder, zeros);
            REQUIRE(run3.size() == 2 && run3[0].size() == 2 && next_zero == 16);
            dump(run3[0]);
            dump(run3[1]);
            // the fused resize graphs follow the reference (no relinearisation) and say so
            bool refused = false;
            try {
                RowResizer rr(evaluator, encoder, transparent, 5, 6, 3, 3, BICUBIC);
            } catch (const std::invalid_argument &) {
                refused = true;
            }
            REQUIRE(refused);
            evaluator.set_relinearization(nullptr);
        }
        std::printf("PASS\n");
        return 0;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "exception: %s\n", e.what());
        return 1;
    }
}
