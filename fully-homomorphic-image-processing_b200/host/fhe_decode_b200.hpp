// fhe_decode_b200.hpp -- the run-length-decode op library of the reference (homo/fhe_decode.h) on
// B200: the homomorphic Taylor sine / cosine and the Fourier step approximation built on them.
//
//   homomorphic_sin (x, res, evaluator, encoder, encryptor)            homo/fhe_decode.h:48-120
//   homomorphic_cos (x, res, evaluator, encoder, encryptor)            homo/fhe_decode.h:128-200
//   approximated_step(amplitude, index, count, order, degree, delta, width, height, run,
//                     evaluator, encoder, encryptor)                    homo/fhe_decode.h:202-242
//
// Op for op the reference's graphs, quirks included: both series expand
// around 3 pi / 2 (:57, :137); `offset` receives add_plain(encode(i)) INSIDE the j loop (:229);
// the fresh `encrypt(encode(0.0))` each series starts from (:54, :134) comes from the caller's
// Encryptor.  Without relinearisation the ciphertexts grow: sin / cos return size 11, a step
// term size 21, each run element size 22.  The decrypting overload of approximated_step
// (:244-282) needs the SECRET key on the server and is out of scope.
//
// `homomorphic_cos` is declared `Ciphertext` in the reference but never returns (:128, UB); here it
// returns void and delivers the result through `res`, as every caller expects.
#pragma once
#include <cmath>

#include "seal_b200.hpp"

namespace seal_b200 {

namespace detail {
inline void fresh_encrypt(Encryptor &e, const Plaintext &p, Ciphertext &res, const Ciphertext &) { e.encrypt(p, res); }
inline void fresh_encrypt(Encryptor &e, const Plaintext &p, CiphertextBatch &res, const CiphertextBatch &like)
{
    e.encrypt(p, res, like.count());
}
// The part of a series that depends on the argument only: the sum of the five weighted powers of
// (x - 3 pi / 2) (size 11 without relinearisation).
template <class CT>
inline CT taylor_terms(const CT &x, double sign, Evaluator &evaluator, FractionalEncoder &encoder)
{
    CT shifted_x(x);
    evaluator.add_plain(shifted_x, encoder.encode(-3 * M_PI / 2.0));

    // The reference recomputes every power from shifted_x (fhe_decode.h:66-111: 1 + 2 + 4 + 3 + 5 = 15
    // ciphertext products); Evaluator::square / multiply are deterministic, so square(shifted_x) is
    // the same ciphertext every time it is formed and the chain below -- x^2, x^4, x^8 once, x^6 and
    // x^10 from x^4 and x^8 -- gives bit-identical terms with 7 products.
    CT power2(shifted_x);
    evaluator.square(power2);                 // x^2
    CT power4(power2);
    evaluator.square(power4);                 // x^4 = square(square(x))
    CT power8(power4);
    evaluator.square(power8);                 // x^8
    CT power6(power4);
    evaluator.multiply(power6, shifted_x);
    evaluator.multiply(power6, shifted_x);    // x^6 = (x^4 * x) * x
    CT power10(power8);
    evaluator.multiply(power10, shifted_x);
    evaluator.multiply(power10, shifted_x);   // x^10 = (x^8 * x) * x

    evaluator.multiply_plain(power2, encoder.encode(sign * 0.5));
    evaluator.multiply_plain(power4, encoder.encode(sign * (-1.0 / 24.0)));
    evaluator.multiply_plain(power6, encoder.encode(sign * (1.0 / 720.0)));
    evaluator.multiply_plain(power8, encoder.encode(sign * (-1.0 / 40320.0)));
    evaluator.multiply_plain(power10, encoder.encode(sign * (1.0 / 3628800.0)));

    // ciphertext addition is exact modular addition per coefficient, hence associative: summing the
    // terms first and adding the fresh encryption afterwards gives the bits of the reference's
    // ((((res + 1) + p2) + p4) + ...) chain
    evaluator.add(power10, power2);
    evaluator.add(power10, power4);
    evaluator.add(power10, power6);
    evaluator.add(power10, power8);
    return power10;
}
// shared body: the two series differ only in the sign of every coefficient.  CT = Ciphertext (the
// reference's call) or CiphertextBatch (the same graph on many independent arguments at once).
// `res` holds the fresh encryption of 0.0 on entry (fhe_decode.h:54, :134)
template <class CT>
inline void taylor_3pi2_core(CT &x, CT &res, double sign, Evaluator &evaluator, FractionalEncoder &encoder)
{
    const CT terms = taylor_terms(x, sign, evaluator, encoder);
    evaluator.add_plain(res, encoder.encode(sign * -1.0));   // 0th order term
    evaluator.add(res, terms);
}
template <class CT>
inline void taylor_3pi2(CT &x, CT &res, double sign, Evaluator &evaluator, FractionalEncoder &encoder,
                        Encryptor &encryptor)
{
    fresh_encrypt(encryptor, encoder.encode(0.0), res, x);
    taylor_3pi2_core(x, res, sign, evaluator, encoder);
}
}  // namespace detail

inline void homomorphic_sin(Ciphertext &x, Ciphertext &res, Evaluator &evaluator, FractionalEncoder &encoder,
                            Encryptor &encryptor)
{
    detail::taylor_3pi2(x, res, 1.0, evaluator, encoder, encryptor);
}
inline void homomorphic_cos(Ciphertext &x, Ciphertext &res, Evaluator &evaluator, FractionalEncoder &encoder,
                            Encryptor &encryptor)
{
    detail::taylor_3pi2(x, res, -1.0, evaluator, encoder, encryptor);
}

inline void homomorphic_sin(CiphertextBatch &x, CiphertextBatch &res, Evaluator &evaluator, FractionalEncoder &encoder,
                            Encryptor &encryptor)
{
    detail::taylor_3pi2(x, res, 1.0, evaluator, encoder, encryptor);
}
inline void homomorphic_cos(CiphertextBatch &x, CiphertextBatch &res, Evaluator &evaluator, FractionalEncoder &encoder,
                            Encryptor &encryptor)
{
    detail::taylor_3pi2(x, res, -1.0, evaluator, encoder, encryptor);
}

// CT = Ciphertext: the reference's call (one run-length pair).  CT = CiphertextBatch: the same graph
// on `count()` independent (amplitude, index, count) triples at once.
template <class CT>
inline void approximated_step(CT &amplitude, CT &index, CT &count, int order, int degree,
                              double delta, int width, int height, std::vector<CT> &run, Evaluator &evaluator,
                              FractionalEncoder &encoder, Encryptor &encryptor)
{
    CT b(count);
    evaluator.multiply_plain(b, encoder.encode(0.5));
    CT offset(index);
    evaluator.add(offset, b);
    evaluator.add_plain(offset, encoder.encode(-0.5));
    evaluator.negate(offset);
    evaluator.add_plain(b, encoder.encode(delta - 0.5));
    for (int i = 0; i < width * height; i++) {
        CT c(b);
        evaluator.multiply_plain(c, encoder.encode(1.0 / ((double)order)));
        for (int j = 1; j <= degree; j++) {
            double arg_factor = ((float)j) * M_PI / ((double)order);
            CT sin_arg(b);
            evaluator.multiply_plain(sin_arg, encoder.encode(arg_factor));
            CT cos_arg(offset);
            evaluator.add_plain(offset, encoder.encode(i));
            evaluator.multiply_plain(cos_arg, encoder.encode(arg_factor));
            CT sin_factor, cos_factor;
            homomorphic_sin(sin_arg, sin_factor, evaluator, encoder, encryptor);
            homomorphic_cos(cos_arg, cos_factor, evaluator, encoder, encryptor);
            CT term(sin_factor);
            evaluator.multiply(term, cos_factor);
            evaluator.multiply_plain(term, encoder.encode(2.0 / (M_PI * ((float)j))));
            evaluator.add(c, term);
        }
        evaluator.multiply(c, amplitude);
        run.push_back(c);
    }
}

// ---------------------------------------------------------------------------------------------
// B200-first variant of approximated_step: identical arithmetic and identical results for ANY
// Encryptor (the server-side encryptions are drawn in the reference's call order: sin then cos for
// j = 1..degree inside i = 0..width*height-1), but the width*height*degree sine and cosine series --
// the expensive part, 15 ciphertext products each -- run as TWO batched series over all (i, j), and
// the closing products as one batched call per step.  Elements are stored j-major so that the
// per-j plaintext factors act on contiguous slices.
inline void approximated_step_batched(Ciphertext &amplitude, Ciphertext &index, Ciphertext &count, int order, int degree,
                                      double delta, int width, int height, std::vector<Ciphertext> &run,
                                      Evaluator &evaluator, FractionalEncoder &encoder, Encryptor &encryptor)
{
    const Context &ctx = evaluator.context();
    const size_t npos = (size_t)width * height, E = npos * (size_t)degree;
    Ciphertext b(count);
    evaluator.multiply_plain(b, encoder.encode(0.5));
    Ciphertext offset(index);
    evaluator.add(offset, b);
    evaluator.add_plain(offset, encoder.encode(-0.5));
    evaluator.negate(offset);
    evaluator.add_plain(b, encoder.encode(delta - 0.5));
    if (degree < 1 || npos == 0) {  // no series: c = b / order, times amplitude
        for (size_t i = 0; i < npos; ++i) {
            Ciphertext c(b);
            evaluator.multiply_plain(c, encoder.encode(1.0 / ((double)order)));
            evaluator.multiply(c, amplitude);
            run.push_back(c);
        }
        return;
    }
    CiphertextBatch cos_arg(ctx, E, 2), sin_f(ctx, E, 2), cos_f(ctx, E, 2);
    Ciphertext z;
    for (size_t i = 0; i < npos; ++i)
        for (int j = 1; j <= degree; ++j) {
            const size_t e = (size_t)(j - 1) * npos + i;
            cos_arg.set(e, offset);                                   // Ciphertext cos_arg(offset);
            evaluator.add_plain(offset, encoder.encode((double)i));   // :229, inside the j loop
            encryptor.encrypt(encoder.encode(0.0), z);                // homomorphic_sin's res (:54)
            sin_f.set(e, z);
            encryptor.encrypt(encoder.encode(0.0), z);                // homomorphic_cos's res (:134)
            cos_f.set(e, z);
        }
    // The sine argument b * j pi / order does not depend on the position (fhe_decode.h:224-226): its
    // powers and their weighted sum are the same ciphertexts for every i, so they are formed once per
    // j (`degree` series instead of width * height * degree); only the fresh encryption each
    // homomorphic_sin starts from differs per call and is added per element below.
    CiphertextBatch sin_args(ctx, (size_t)degree, 2);
    for (int j = 1; j <= degree; ++j) {
        const double arg_factor = ((float)j) * M_PI / ((double)order);
        Ciphertext sa(b);
        evaluator.multiply_plain(sa, encoder.encode(arg_factor));     // the same ciphertext for every i
        sin_args.set((size_t)(j - 1), sa);
        CiphertextBatch cj = cos_arg.slice((size_t)(j - 1) * npos, npos);
        evaluator.multiply_plain(cj, encoder.encode(arg_factor));     // in place on the slice
    }
    {
        const CiphertextBatch sin_terms = detail::taylor_terms(sin_args, 1.0, evaluator, encoder);
        CiphertextBatch rep(ctx, E, sin_terms.size());
        for (int j = 1; j <= degree; ++j) {
            const Ciphertext tj = sin_terms.get((size_t)(j - 1));
            for (size_t i = 0; i < npos; ++i) rep.set((size_t)(j - 1) * npos + i, tj);
        }
        evaluator.add_plain(sin_f, encoder.encode(-1.0));              // res + (-1), per element
        evaluator.add(rep, sin_f);                                     // + the shared terms
        sin_f = std::move(rep);
    }
    detail::taylor_3pi2_core(cos_arg, cos_f, -1.0, evaluator, encoder);
    evaluator.multiply(sin_f, cos_f);                                  // term = sin_factor * cos_factor
    CiphertextBatch c(ctx, npos, 2);
    {
        Ciphertext c0(b);
        evaluator.multiply_plain(c0, encoder.encode(1.0 / ((double)order)));
        for (size_t i = 0; i < npos; ++i) c.set(i, c0);
    }
    for (int j = 1; j <= degree; ++j) {
        CiphertextBatch tj = sin_f.slice((size_t)(j - 1) * npos, npos);
        evaluator.multiply_plain(tj, encoder.encode(2.0 / (M_PI * ((float)j))));
        evaluator.add(c, tj);
    }
    CiphertextBatch amp(ctx, npos, amplitude.size());
    for (size_t i = 0; i < npos; ++i) amp.set(i, amplitude);
    evaluator.multiply(c, amp);
    for (size_t i = 0; i < npos; ++i) run.push_back(c.get(i));
}

}  // namespace seal_b200
