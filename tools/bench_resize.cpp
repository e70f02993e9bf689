// bench_resize.cpp -- BASELINE.json config 3 (server_resize bicubic, FractionalEncoder, n=8192) on
// one B200, through the C++ host layer: ResizeImageBatched (host/fhe_resize_b200.hpp) on synthetic
// uniform ciphertexts with the key-less transparent Encryptor for the public fractions.
// The geometry is scaled down (default 128x16 -> 64x8, 2:1 like 512x512 -> 256x256); the cost per
// output pixel does not depend on the image size, so pixels/s extrapolates to the full config.
// With a 4th argument dbc > 0 every ct x ct product is relinearised (Evaluator::set_relinearization,
// synthetic keys): outputs stay size 2 instead of growing to 6 (bicubic) / 4 (bilinear).
// Not the headline bench (bench.py).  Usage: bench_resize [n] [ow] [oh] [dbc]   -> one JSON line per filter
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>
#include <sstream>

#include "../fully-homomorphic-image-processing_b200/host/fhe_resize_b200.hpp"

using namespace seal_b200;

int main(int argc, char **argv)
{
    try {
        const uint32_t n = argc > 1 ? (uint32_t)std::atoi(argv[1]) : 8192;
        const int ow = argc > 2 ? std::atoi(argv[2]) : 128, oh = argc > 3 ? std::atoi(argv[3]) : 16;
        const int dw = ow / 2, dh = oh / 2;
        const uint32_t dbc = argc > 4 ? (uint32_t)std::atoi(argv[4]) : 0;
        Context ctx(n);
        Evaluator evaluator(ctx);
        std::unique_ptr<EvaluationKeys> evk;
        if (dbc) {
            const uint32_t k = ctx.coeff_mod_count(), digits = fhe_evk_digits(ctx.handle(), dbc);
            std::vector<uint64_t> kw((size_t)k * digits * 2 * k * n);
            uint64_t z = 0x5EED;
            for (auto &w : kw) { z = z * 6364136223846793005ull + 1442695040888963407ull; w = z >> 12; }
            evk.reset(new EvaluationKeys(ctx, kw.data(), dbc, 1));
            evaluator.set_relinearization(evk.get());
        }
        FractionalEncoder encoder(ctx, 100, 100);
        Encryptor encryptor(ctx);  // transparent encryption of the public fractions
        uint64_t next = 0;
        CiphertextSource source = [&](Ciphertext &dst) {
            Ciphertext c(ctx, 2);
            check(fhe_fill_uniform(ctx.handle(), c.device_data(), next++, 1, 2, 0xC0F3));
            dst = std::move(c);
        };
        struct Quiet {
            std::ostringstream sink;
            std::streambuf *old;
            Quiet() : old(std::cout.rdbuf(sink.rdbuf())) {}
            ~Quiet() { std::cout.rdbuf(old); }
        } quiet;
        for (int inter : {BICUBIC, BILINEAR}) {
            // the first pass grows the stream-ordered memory pool through the driver (hundreds of MiB
            // blocks at size 6: slow and noisy, and paid once per process); the second and third are
            // the steady state a server streaming rows is in.  The best of those is reported, the
            // first beside it.
            double s = 1e30, first = 0;
            uint64_t launches = 0;
            SImageData dest;
            for (int rep = 0; rep < 3; ++rep) {
                next = 0;
                dest = SImageData{};
                const uint64_t l0 = fhe_launch_count(ctx.handle());
                ctx.sync();
                auto t0 = std::chrono::steady_clock::now();
                ResizeImageBatched(source, ow, oh, dest, dw, dh, inter, evaluator, encoder, encryptor);
                ctx.sync();
                const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                launches = fhe_launch_count(ctx.handle()) - l0;
                if (rep == 0) first = sec;
                else if (sec < s) s = sec;
            }
            const double px = (double)dw * dh;
            std::printf("{\"what\": \"ResizeImageBatched %s, FractionalEncoder(100,100), n=%u k=%u, %dx%d -> %dx%d RGB, "
                        "synthetic ciphertexts, transparent fractions, relinearisation dbc %u (0 = none, as the reference)\", \"seconds\": %.4f, \"output_pixels_per_s\": %.2f, "
                        "\"%s_calls_per_s\": %.1f, \"first_pass_seconds\": %.4f, \"out_ct_size\": %u, \"kernel_launches\": %llu, "
                        "\"full_config3_512x512_to_256x256_estimate_s\": %.1f}\n",
                        inter == BICUBIC ? "bicubic" : "bilinear", n, ctx.coeff_mod_count(), ow, oh, dw, dh, dbc, s, px / s,
                        inter == BICUBIC ? "Cubic" : "Linear", (inter == BICUBIC ? 15.0 : 9.0) * px / s, first,
                        dest.pixels[0][0].size(), (unsigned long long)launches, 65536.0 / (px / s));
        }
        return 0;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "exception: %s\n", e.what());
        return 1;
    }
}
