// test_host_shim.cpp -- the C++ host layer (host/seal_b200.hpp, host/fhe_image_b200.hpp) over
// the C-ABI, on a real B200:
//   * the reference-style op-by-op evaluation of the 8x8 DCT + quantize graph through
//     Evaluator::{add, sub, multiply_plain} equals the fused device call, bit for bit;
//   * Ciphertext value semantics (copies are independent, block views stay contiguous);
//   * SEAL's error behaviour (std::invalid_argument on a zero plaintext / foreign ciphertext).
// Prints "PASS" and exits 0 on success.  Usage: test_host_shim [n] [dump_file]
// With dump_file: writes the input block and the fused result (raw u64) for the pytest wrapper
// to compare against the CPU oracle.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>

#include "../../fully-homomorphic-image-processing_b200/host/fhe_image_b200.hpp"

using namespace seal_b200;

#define REQUIRE(cond)                                                        \
    do {                                                                     \
        if (!(cond)) {                                                       \
            std::fprintf(stderr, "FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            std::_Exit(1); /* std::cout is redirected into a local: no unwinding */ \
        }                                                                    \
    } while (0)

// One 1-D pass of the LLM "islow" network the reference uses (mirror of dct() at
// homo/fhe_image.h:400-484), written against the Evaluator API the way fhe_image.h does:
// copy-construct, then in-place op.  `scale` = the extra multiply_plain by encode(0.125) of the
// column pass (fhe_image.h:259).
static void dct_pass_by_ops(Ciphertext *d[8], Evaluator &ev, FractionalEncoder &enc, bool scale)
{
    auto add = [&](const Ciphertext &a, const Ciphertext &b) { Ciphertext x(a); ev.add(x, b); return x; };
    auto sub = [&](const Ciphertext &a, const Ciphertext &b) { Ciphertext x(a); ev.sub(x, b); return x; };
    auto mul = [&](const Ciphertext &a, double c) { Ciphertext x(a); ev.multiply_plain(x, enc.encode(c)); return x; };
    auto out = [&](const Ciphertext &a) { return scale ? mul(a, 0.125) : Ciphertext(a); };
    Ciphertext t0 = add(*d[0], *d[7]), t7 = sub(*d[0], *d[7]);
    Ciphertext t1 = add(*d[1], *d[6]), t6 = sub(*d[1], *d[6]);
    Ciphertext t2 = add(*d[2], *d[5]), t5 = sub(*d[2], *d[5]);
    Ciphertext t3 = add(*d[3], *d[4]), t4 = sub(*d[3], *d[4]);
    Ciphertext t10 = add(t0, t3), t13 = sub(t0, t3), t11 = add(t1, t2), t12 = sub(t1, t2);
    *d[0] = out(add(t10, t11));
    *d[4] = out(sub(t10, t11));
    Ciphertext z1 = mul(add(t12, t13), 0.541196100);
    *d[2] = out(add(z1, mul(t13, 0.765366865)));
    *d[6] = out(add(z1, mul(t12, -1.847759065)));
    z1 = add(t4, t7);
    Ciphertext z2 = add(t5, t6), z3 = add(t4, t6), z4 = add(t5, t7);
    Ciphertext z5 = mul(add(z3, z4), 1.175875602);
    t4 = mul(t4, 0.298631336);
    t5 = mul(t5, 2.053119869);
    t6 = mul(t6, 3.072711026);
    t7 = mul(t7, 1.501321110);
    z1 = mul(z1, -0.899976223);
    z2 = mul(z2, -2.562915447);
    z3 = add(mul(z3, -1.961570560), z5);
    z4 = add(mul(z4, -0.390180644), z5);
    *d[7] = out(add(add(t4, z1), z3));
    *d[5] = out(add(add(t5, z2), z4));
    *d[3] = out(add(add(t6, z2), z3));
    *d[1] = out(add(add(t7, z1), z4));
}

static void dct_by_ops(std::vector<Ciphertext> &data, Evaluator &ev, FractionalEncoder &enc)
{
    for (int r = 0; r < 8; ++r) {
        Ciphertext *row[8];
        for (int i = 0; i < 8; ++i) row[i] = &data[8 * r + i];
        dct_pass_by_ops(row, ev, enc, false);
    }
    for (int c = 0; c < 8; ++c) {
        Ciphertext *col[8];
        for (int i = 0; i < 8; ++i) col[i] = &data[8 * i + c];
        dct_pass_by_ops(col, ev, enc, true);
    }
}

// Independent deep copies of a block that are NOT one contiguous allocation: the device allocator
// is a stream-ordered pool, which may well hand consecutive requests adjacent addresses, so a live
// spacer ciphertext is allocated between consecutive copies.
static std::vector<Ciphertext> scattered_copies(const std::vector<Ciphertext> &block, std::vector<Ciphertext> &spacers)
{
    std::vector<Ciphertext> out;
    out.reserve(block.size());
    for (size_t i = 0; i < block.size(); ++i) {
        out.emplace_back(block[i]);
        spacers.emplace_back(block[i]);
    }
    return out;
}

static std::vector<uint64_t> fetch(const std::vector<Ciphertext> &v)
{
    const size_t w = v[0].context()->ct_words(2);
    std::vector<uint64_t> h(v.size() * w);
    for (size_t i = 0; i < v.size(); ++i) v[i].download(h.data() + i * w);
    return h;
}

int main(int argc, char **argv)
{
    const uint32_t n = argc > 1 ? (uint32_t)std::atoi(argv[1]) : 2048;
    std::ostringstream sink;  // the image ops print CSV timings like the reference; keep stdout clean
    std::streambuf *old = std::cout.rdbuf(sink.rdbuf());

    Context ctx(n);  // coeff_modulus_128(n), plain_modulus 1 << 14 (fhe_image.h:26)
    Evaluator evaluator(ctx);
    FractionalEncoder encoder(ctx);  // (t, poly, 100, 100, 2), server_jpeg.cpp:100
    Encryptor encryptor;
    const std::vector<double> quant = {16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57,
                                       69, 56, 14, 17, 22, 29, 51, 87, 80, 62, 18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64,
                                       81, 104, 113, 92, 49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};

    // a block of 64 synthetic ciphertexts, contiguous on the device
    std::vector<Ciphertext> block = load_block(ctx, nullptr, 64);
    REQUIRE(is_contiguous(block));
    check(fhe_fill_uniform(ctx.handle(), block[0].device_data(), 0, 64, 2, 0x5EA1B200ull));
    const std::vector<uint64_t> input = fetch(block);

    // (1) op by op, on independent copies (what the reference does)
    std::vector<Ciphertext> spacers;
    std::vector<Ciphertext> by_ops = scattered_copies(block, spacers);
    REQUIRE(!is_contiguous(by_ops));
    dct_by_ops(by_ops, evaluator, encoder);
    const std::vector<uint64_t> dct_ops = fetch(by_ops);
    quantize_fhe(by_ops, quant, evaluator, encoder, encryptor);
    const std::vector<uint64_t> dctq_ops = fetch(by_ops);
    REQUIRE(fetch(block) == input);  // copies are independent

    // (2) fused encrypted_dct on scattered copies, then quantize_fhe op by op
    std::vector<Ciphertext> scattered = scattered_copies(block, spacers);
    REQUIRE(!is_contiguous(scattered));
    encrypted_dct(scattered, evaluator, encoder, encryptor);
    REQUIRE(fetch(scattered) == dct_ops);
    quantize_fhe(scattered, quant, evaluator, encoder, encryptor);
    REQUIRE(fetch(scattered) == dctq_ops);

    // (3) fully fused, in place on the contiguous block
    encrypted_dct_quantize(block, quant, evaluator, encoder, encryptor);
    REQUIRE(is_contiguous(block));
    const std::vector<uint64_t> fused = fetch(block);
    REQUIRE(fused == dctq_ops);

    // (4) rgb_to_ycc_fhe: linear in (r, g, b) up to the -128 offset -> check Y(r,g,b) + Y(r,g,b) == Y(2r,2g,2b) - 128
    {
        std::vector<Ciphertext> px = load_block(ctx, nullptr, 3);
        check(fhe_fill_uniform(ctx.handle(), px[0].device_data(), 1000, 3, 2, 7));
        Ciphertext r(px[0]), g(px[1]), b(px[2]);
        Ciphertext r2(r), g2(g), b2(b);
        evaluator.add(r2, r); evaluator.add(g2, g); evaluator.add(b2, b);
        rgb_to_ycc_fhe(r, g, b, evaluator, encoder, encryptor);
        rgb_to_ycc_fhe(r2, g2, b2, evaluator, encoder, encryptor);
        Ciphertext yy(r);
        evaluator.add(yy, r);                              // 2Y - 256
        evaluator.add_plain(yy, encoder.encode(128.0));    // 2Y - 128
        std::vector<uint64_t> a(ctx.ct_words()), c(ctx.ct_words());
        yy.download(a.data()); r2.download(c.data());
        REQUIRE(a == c);
        Ciphertext uu(g); evaluator.add(uu, g);
        uu.download(a.data()); g2.download(c.data());
        REQUIRE(a == c);
    }

    // (5) error behaviour
    bool threw = false;
    try { evaluator.multiply_plain(block[0], Plaintext(std::vector<uint64_t>(n, 0))); } catch (const std::invalid_argument &) { threw = true; }
    REQUIRE(threw);
    threw = false;
    try { Ciphertext empty; evaluator.add(block[0], empty); } catch (const std::invalid_argument &) { threw = true; }
    REQUIRE(threw);
    threw = false;
    try { Context bad(n, 1 << 14, {0x7fffffff380003ull}); } catch (const std::invalid_argument &) { threw = true; }
    REQUIRE(threw);

    if (argc > 2) {
        std::ofstream f(argv[2], std::ios::binary);
        f.write((const char *)input.data(), input.size() * 8);
        f.write((const char *)fused.data(), fused.size() * 8);
    }
    std::cout.rdbuf(old);
    // CSV contract: every image op printed "<ms>," (fhe_image.h:285-286)
    REQUIRE(!sink.str().empty() && sink.str().back() == ',');
    std::printf("PASS n=%u ops=%zu\n", n, (size_t)fhe_launch_count(ctx.handle()));
    return 0;
}
