// bench_decode.cpp -- BASELINE.json config 5 (server_decode homomorphic step approximation, 8x8 blocks
// = 64 positions, n=8192, deep multiply) on one B200 through the C++ host layer:
// approximated_step_batched (host/fhe_decode_b200.hpp) on synthetic uniform ciphertexts, transparent
// server-side encryptions.  One run-length pair expands to width*height ciphertexts of size 22 through
// 2 * positions * degree Taylor series (15 ciphertext products each in the reference's graph, sizes up
// to 11; the host layer forms each distinct product once: 7 per series, sine series shared over positions).
// With a 5th argument dbc > 0 every ct x ct product is relinearised (Evaluator::set_relinearization,
// "deep-multiply + relinearize" of config 5) with synthetic evaluation keys: all operands stay size 2.
// Not the headline bench (bench.py).  Usage: bench_decode [n] [positions] [degree] [order] [dbc] -> JSON line
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>
#include <vector>

#include "../fully-homomorphic-image-processing_b200/host/fhe_decode_b200.hpp"

using namespace seal_b200;

int main(int argc, char **argv)
{
    try {
        const uint32_t n = argc > 1 ? (uint32_t)std::atoi(argv[1]) : 8192;
        const int npos = argc > 2 ? std::atoi(argv[2]) : 64, degree = argc > 3 ? std::atoi(argv[3]) : 12;
        const int order = argc > 4 ? std::atoi(argv[4]) : 64;
        const uint32_t dbc = argc > 5 ? (uint32_t)std::atoi(argv[5]) : 0;
        Context ctx(n);
        Evaluator evaluator(ctx);
        std::unique_ptr<EvaluationKeys> evk;
        if (dbc) {
            // timing only: any residues below every q_j serve as keys
            const uint32_t k = ctx.coeff_mod_count(), digits = fhe_evk_digits(ctx.handle(), dbc);
            std::vector<uint64_t> kw((size_t)k * digits * 2 * k * n);
            uint64_t z = 0x5EED;
            for (auto &w : kw) { z = z * 6364136223846793005ull + 1442695040888963407ull; w = z >> 12; }
            evk.reset(new EvaluationKeys(ctx, kw.data(), dbc, 1));
            evaluator.set_relinearization(evk.get());
        }
        FractionalEncoder encoder(ctx, 100, 100);
        Encryptor encryptor(ctx);
        Ciphertext amplitude(ctx, 2), index(ctx, 2), count(ctx, 2);
        check(fhe_fill_uniform(ctx.handle(), amplitude.device_data(), 0, 1, 2, 0xDEC));
        check(fhe_fill_uniform(ctx.handle(), index.device_data(), 1, 1, 2, 0xDEC));
        check(fhe_fill_uniform(ctx.handle(), count.device_data(), 2, 1, 2, 0xDEC));
        double best = 1e30;
        unsigned long long launches = 0;
        size_t out_size = 0;
        std::string all;
        for (int rep = 0; rep < 4; ++rep) {
            std::vector<Ciphertext> run;
            const uint64_t l0 = fhe_launch_count(ctx.handle());
            ctx.sync();
            auto t0 = std::chrono::steady_clock::now();
            approximated_step_batched(amplitude, index, count, order, degree, 0.5, npos, 1, run, evaluator, encoder, encryptor);
            ctx.sync();
            const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (s < best) best = s;
            char buf[32];
            std::snprintf(buf, sizeof buf, "%s%.4f", rep ? ", " : "", s);
            all += buf;
            launches = fhe_launch_count(ctx.handle()) - l0;
            out_size = run.at(0).size();
        }
        // the reference's graph: 2 series per (position, j), 15 ciphertext products each, + sin*cos + amplitude;
        // formed here: 7 products per distinct series (cosine: one per (position, j); sine: one per j)
        const double series = 2.0 * npos * degree, products = series * 15 + (double)npos * degree + npos;
        const double formed = 7.0 * ((double)npos * degree + degree) + (double)npos * degree + npos;
        std::printf("{\"what\": \"approximated_step_batched (homomorphic RLE step, fhe_decode.h:202-242), n=%u k=%u, "
                    "%d positions, degree %d, order %d, synthetic ciphertexts, transparent server-side encryptions, relinearisation dbc %u (0 = none, as the reference)\", "
                    "\"seconds_per_run_length_pair\": %.4f, \"taylor_series_per_s\": %.1f, \"reference_graph_ct_products_per_s\": %.1f, \"ct_products_formed_per_s\": %.1f, "
                    "\"seconds_each_pass\": [%s], \"out_ct_size\": %zu, \"kernel_launches\": %llu}\n",
                    n, ctx.coeff_mod_count(), npos, degree, order, dbc, best, series / best, products / best, formed / best, all.c_str(), out_size, launches);
        return 0;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "exception: %s\n", e.what());
        return 1;
    }
}
