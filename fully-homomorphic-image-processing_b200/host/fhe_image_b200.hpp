// fhe_image_b200.hpp -- the JPEG image-op library of the reference (homo/fhe_image.h) on B200.
//
// Same call surface as the reference's three homomorphic functions, so server_jpeg.cpp's block
// loop (:113-138) compiles against it after swapping the include and the namespace
// (INTEGRATION.md):
//
//   encrypted_dct (data, evaluator, encoder, encryptor)          homo/fhe_image.h:196-288
//   quantize_fhe  (data, quant, evaluator, encoder, encryptor)   homo/fhe_image.h:294-305
//   rgb_to_ycc_fhe(r, g, b, evaluator, encoder, encryptor)       homo/fhe_image.h:310-325
//
// plus the fused entry point the BASELINE metric is defined on:
//
//   encrypted_dct_quantize(data, quant, ...) == encrypted_dct followed by quantize_fhe
//
// The reference evaluates these graphs one Evaluator call at a time (768 ciphertext copies and
// 256 multiply_plain per block).  Here a block goes through ONE fused call of the C-ABI
// (fhe_dct8x8_quantize): forward NTT of the 64 ciphertexts, the whole 2-D butterfly network
// pointwise in the NTT domain, inverse NTT.  Every step of the reference graph is an exact ring
// operation, so the canonical residues -- the bits SEAL would produce -- are identical.
// Like the reference, each function prints its wall time in ms followed by ',' to stdout
// (the CSV contract benchmark/analyze.py parses).
#pragma once
#include <chrono>
#include <iostream>

#include "seal_b200.hpp"

namespace seal_b200 {

// fhe_image.h:221-236,259
static const double kDctConstants[12] = {0.541196100, 0.765366865, -1.847759065, 1.175875602,
                                         0.298631336, 2.053119869, 3.072711026,  1.501321110,
                                         -0.899976223, -2.562915447, -1.961570560, -0.390180644};

namespace detail {

struct DctPlan {
    fhe_dct_plan *h = nullptr;
    ~DctPlan() { fhe_dct_plan_destroy(h); }
};

inline fhe_dct_plan *dct_plan(Evaluator &ev, FractionalEncoder &enc, const std::vector<double> *quant)
{
    std::string key = "dct:" + enc.key() + ":";  // the tables are encodings: another encoder, another plan
    if (quant) key.append((const char *)quant->data(), quant->size() * sizeof(double));
    std::shared_ptr<void> &slot = ev.plan_slot(key);
    if (!slot) {
        const uint32_t n = ev.context().poly_degree();
        std::vector<uint64_t> consts((size_t)12 * n), eighth, q;
        for (int i = 0; i < 12; ++i) {
            Plaintext p = enc.encode(kDctConstants[i]);
            std::memcpy(consts.data() + (size_t)i * n, p.coeffs().data(), n * sizeof(uint64_t));
        }
        eighth = enc.encode(0.125).coeffs();
        if (quant) {
            if (quant->size() != 64) throw std::invalid_argument("quant must hold 64 entries");
            q.resize((size_t)64 * n);
            for (int i = 0; i < 64; ++i) {
                Plaintext p = enc.encode(1 / (*quant)[i]);  // fhe_image.h:301
                std::memcpy(q.data() + (size_t)i * n, p.coeffs().data(), n * sizeof(uint64_t));
            }
        }
        auto plan = std::make_shared<DctPlan>();
        check(fhe_dct_plan_create(ev.context().handle(), consts.data(), eighth.data(), quant ? q.data() : nullptr, &plan->h));
        slot = plan;
    }
    return static_cast<DctPlan *>(slot.get())->h;
}

inline void run_block(std::vector<Ciphertext> &data, Evaluator &ev, fhe_dct_plan *plan)
{
    if (data.size() != 64) throw std::invalid_argument("a block is 64 ciphertexts, rows first");
    const Context &ctx = ev.context();
    for (const Ciphertext &c : data)
        if (!c.valid() || c.size() != 2 || c.context() != &ctx) throw std::invalid_argument("encrypted is not valid for encryption parameters");
    if (is_contiguous(data)) {
        check(fhe_dct8x8_quantize(ctx.handle(), plan, data[0].device_data(), data[0].device_data(), 1));
        return;
    }
    // scattered ciphertexts: gather, transform, scatter (device to device)
    const size_t w = ctx.ct_words(2);
    Arena tmp(ctx, 64 * w);
    for (int i = 0; i < 64; ++i) check(fhe_device_copy(ctx.handle(), tmp.data() + i * w, data[i].device_data(), w * 8));
    check(fhe_dct8x8_quantize(ctx.handle(), plan, tmp.data(), tmp.data(), 1));
    for (int i = 0; i < 64; ++i) check(fhe_device_copy(ctx.handle(), data[i].device_data(), tmp.data() + i * w, w * 8));
    ctx.sync();  // tmp is released on return
}

struct CsvTimer {
    const Context &ctx;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    explicit CsvTimer(const Context &c) : ctx(c) {}
    ~CsvTimer()
    {
        fhe_ctx_sync(ctx.handle());
        auto d = std::chrono::steady_clock::now() - t0;
        std::cout << std::chrono::duration<double, std::milli>(d).count() << ',';
    }
};

}  // namespace detail

// In-place 2-D 8x8 forward DCT of a block of 64 ciphertexts laid out rows first.
inline void encrypted_dct(std::vector<Ciphertext> &data, Evaluator &evaluator, FractionalEncoder &encoder,
                          Encryptor &encryptor)
{
    (void)encryptor;
    detail::CsvTimer t(evaluator.context());
    detail::run_block(data, evaluator, detail::dct_plan(evaluator, encoder, nullptr));
}

// data[i] <- data[i] * encode(1 / quant[i])
inline void quantize_fhe(std::vector<Ciphertext> &data, const std::vector<double> &quant, Evaluator &evaluator,
                         FractionalEncoder &encoder, Encryptor &encryptor)
{
    (void)encryptor;
    detail::CsvTimer t(evaluator.context());
    for (int i = 0; i < 64; i++) evaluator.multiply_plain(data[i], encoder.encode(1 / quant[i]));
}

// encrypted_dct then quantize_fhe as one fused device call (the BASELINE.json unit of work).
inline void encrypted_dct_quantize(std::vector<Ciphertext> &data, const std::vector<double> &quant,
                                   Evaluator &evaluator, FractionalEncoder &encoder, Encryptor &encryptor)
{
    (void)encryptor;
    detail::CsvTimer t(evaluator.context());
    detail::run_block(data, evaluator, detail::dct_plan(evaluator, encoder, &quant));
}

// (R, G, B) -> (Y, Cb, Cr) in place, fhe_image.h:317-319.
inline void rgb_to_ycc_fhe(Ciphertext &r, Ciphertext &g, Ciphertext &b, Evaluator &evaluator,
                           FractionalEncoder &encoder, Encryptor &encryptor)
{
    (void)encryptor;
    detail::CsvTimer t(evaluator.context());
    auto term = [&](const Ciphertext &c, double w) {
        Ciphertext x(c);
        evaluator.multiply_plain(x, encoder.encode(w));
        return x;
    };
    Ciphertext y = term(r, 0.299);
    evaluator.add(y, term(g, 0.587));
    evaluator.add(y, term(b, 0.114));
    evaluator.sub_plain(y, encoder.encode(128.0));
    Ciphertext u = term(r, -0.168736);
    evaluator.sub(u, term(g, 0.331264));
    evaluator.add(u, term(b, 0.5));
    Ciphertext v = term(r, 0.5);
    evaluator.sub(v, term(g, 0.418688));
    evaluator.sub(v, term(b, 0.081312));
    r = y;
    g = u;
    b = v;
}

}  // namespace seal_b200
