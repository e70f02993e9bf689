// fhe_decode_b200.hpp -- the run-length-decode graphs of the reference (homo/fhe_decode.h: the Taylor
// sine / cosine around 3 pi / 2 and the Fourier step approximation built on them) as B200-first
// batched evaluations.
//
// The reference's own functions are NOT restated here: homo/fhe_decode.h compiles against
// include/seal/seal.h (one token patched: homomorphic_cos is declared to return a Ciphertext and never
// returns, fhe_decode.h:128) and runs op by op on the GPU (oracle/Makefile `ref`,
// tests/test_reference_sources_on_gpu.py).  This header is the fast path a server uses instead:
//
//   * approximated_step (fhe_decode.h:202-242) evaluates 2 * width * height * degree Taylor series of
//     15 ciphertext products each, one ciphertext at a time.  Here the cosine series of a run-length
//     pair run as ONE batched series over all (position, j), the sine series once per j (its argument
//     b * j pi / order does not depend on the position, :224-226), and the closing products as one
//     batched call each;
//   * inside a series every distinct product is formed once -- x^2, x^4, x^8 by squaring, x^6 and x^10
//     from x^4 and x^8 (7 products instead of 15: Evaluator::square / multiply are deterministic, so
//     square(shifted_x) is the same ciphertext every time the reference recomputes it) -- and every
//     operand is transformed once (OperandForm: shifted_x enters five products);
//   * with Evaluator::set_relinearization every product is brought back to size 2 at once ("deep
//     multiply + relinearize", BASELINE config 5); without it the sizes grow as in the reference
//     (11 per series, 21 per term, 22 per run element).
//
// Same ciphertexts as the reference's graph for ANY Encryptor: the fresh `encrypt(encode(0.0))` each
// series starts from (:54, :134) is drawn in the reference's call order (sin then cos for j = 1..degree
// inside i = 0..width*height-1) and added per element; the quirks are kept -- both series expand around
// 3 pi / 2 (:57, :137), `offset` receives add_plain(encode(i)) INSIDE the j loop (:229), the (float) casts
// of j.  The decrypting overload (:244-282) needs the SECRET key on the server and is out of scope.
#pragma once
#include <cmath>

#include "seal_b200.hpp"

namespace seal_b200 {

namespace detail {
// the part of a series that depends on the argument only: sum of the five weighted powers of
// (x - 3 pi / 2); sign +1: sine (fhe_decode.h:66-99), -1: cosine (:146-179)
inline CiphertextBatch taylor_terms(const CiphertextBatch &x, double sign, Evaluator &evaluator, FractionalEncoder &encoder)
{
    const Context &ctx = evaluator.context();
    CiphertextBatch shifted(x);
    evaluator.add_plain(shifted, encoder.encode(-3 * M_PI / 2.0));
    const OperandForm fx(ctx, shifted);
    CiphertextBatch p2 = evaluator.multiply(fx, fx);                   // square(x)
    const OperandForm f2(ctx, p2);
    CiphertextBatch p4 = evaluator.multiply(f2, f2);                   // square(square(x))
    const OperandForm f4(ctx, p4);
    CiphertextBatch p8 = evaluator.multiply(f4, f4);
    CiphertextBatch p6 = evaluator.multiply(f4, fx);                   // (x^4 * x) * x
    p6 = evaluator.multiply(OperandForm(ctx, p6), fx);
    CiphertextBatch p10 = evaluator.multiply(OperandForm(ctx, p8), fx);  // (x^8 * x) * x
    p10 = evaluator.multiply(OperandForm(ctx, p10), fx);
    evaluator.multiply_plain(p2, encoder.encode(sign * 0.5));
    evaluator.multiply_plain(p4, encoder.encode(sign * (-1.0 / 24.0)));
    evaluator.multiply_plain(p6, encoder.encode(sign * (1.0 / 720.0)));
    evaluator.multiply_plain(p8, encoder.encode(sign * (-1.0 / 40320.0)));
    evaluator.multiply_plain(p10, encoder.encode(sign * (1.0 / 3628800.0)));
    // ciphertext addition is exact modular addition per coefficient, hence associative: any order of
    // the sum gives the bits of the reference's ((((res + 1) + p2) + p4) + ...) chain
    evaluator.add(p10, p2);
    evaluator.add(p10, p4);
    evaluator.add(p10, p6);
    evaluator.add(p10, p8);
    return p10;
}
}  // namespace detail

// res[i] = series(x[i]): `res` holds the fresh encryptions of 0.0 the series start from on entry
// (one per element, drawn by the caller in the reference's order)
inline void taylor_series(const CiphertextBatch &x, CiphertextBatch &res, double sign, Evaluator &evaluator, FractionalEncoder &encoder)
{
    const CiphertextBatch terms = detail::taylor_terms(x, sign, evaluator, encoder);
    evaluator.add_plain(res, encoder.encode(sign * -1.0));   // 0th order term
    evaluator.add(res, terms);
}
// homomorphic_sin / homomorphic_cos on `x.count()` independent arguments
inline void homomorphic_sin(CiphertextBatch &x, CiphertextBatch &res, Evaluator &evaluator, FractionalEncoder &encoder,
                            Encryptor &encryptor)
{
    encryptor.encrypt(encoder.encode(0.0), res, x.count());
    taylor_series(x, res, 1.0, evaluator, encoder);
}
inline void homomorphic_cos(CiphertextBatch &x, CiphertextBatch &res, Evaluator &evaluator, FractionalEncoder &encoder,
                            Encryptor &encryptor)
{
    encryptor.encrypt(encoder.encode(0.0), res, x.count());
    taylor_series(x, res, -1.0, evaluator, encoder);
}

// approximated_step (fhe_decode.h:202-242) of one run-length pair: `run` receives width * height
// ciphertexts.  Elements of the batched series are stored j-major so that the per-j plaintext factors
// act on contiguous slices.
inline void approximated_step_batched(Ciphertext &amplitude, Ciphertext &index, Ciphertext &count, int order, int degree,
                                      double delta, int width, int height, std::vector<Ciphertext> &run,
                                      Evaluator &evaluator, FractionalEncoder &encoder, Encryptor &encryptor)
{
    const Context &ctx = evaluator.context();
    const size_t npos = (size_t)width * height, E = npos * (size_t)(degree > 0 ? degree : 0);
    Ciphertext b(count);
    evaluator.multiply_plain(b, encoder.encode(0.5));
    Ciphertext offset(index);
    evaluator.add(offset, b);
    evaluator.add_plain(offset, encoder.encode(-0.5));
    evaluator.negate(offset);
    evaluator.add_plain(b, encoder.encode(delta - 0.5));
    CiphertextBatch c(ctx, npos, 2);                                   // c = b / order, per position
    {
        Ciphertext c0(b);
        evaluator.multiply_plain(c0, encoder.encode(1.0 / ((double)order)));
        c.fill(0, npos, c0.device_data());
    }
    if (E) {
        // cos_arg(i, j) is the running `offset`, which receives add_plain(encode(i)) after every use
        // (:229, inside the j loop).  add_plain adds Delta * m coefficient-wise and encode(i) has digits
        // 0 / 1, so the offset of iteration e is the first offset plus ONE short plaintext -- the sum of
        // the encodings added before it -- as long as no coefficient of that sum reaches the upper half
        // of the plaintext range (then the reference's own sequence is replayed instead).
        constexpr uint32_t NC = 32;  // an int has at most 31 binary digits
        CiphertextBatch cos_arg(ctx, E, 2), sin_f, cos_f;
        {
            std::vector<uint64_t> sum(NC, 0), each(E * (size_t)NC, 0);
            const uint64_t thr = (ctx.plain_modulus() + 1) >> 1;
            bool closed_form = true;
            for (size_t i = 0; i < npos && closed_form; ++i) {
                const Plaintext pi = encoder.encode((double)i);
                std::vector<uint32_t> digits;  // exponents of encode(i): its coefficients must all be 1
                for (size_t cidx = 0; cidx < pi.coeffs().size(); ++cidx)
                    if (pi.coeffs()[cidx]) {
                        if (cidx >= NC || pi.coeffs()[cidx] != 1) closed_form = false;
                        else digits.push_back((uint32_t)cidx);
                    }
                for (int j = 1; j <= degree && closed_form; ++j) {
                    std::copy(sum.begin(), sum.end(), each.begin() + ((size_t)(j - 1) * npos + i) * NC);
                    for (uint32_t d : digits)
                        if (++sum[d] >= thr) closed_form = false;
                }
            }
            if (closed_form) {
                cos_arg.fill(0, E, offset.device_data());
                evaluator.add_plain_each(cos_arg, each, NC);
            } else {
                for (size_t i = 0; i < npos; ++i)
                    for (int j = 1; j <= degree; ++j) {
                        cos_arg.set((size_t)(j - 1) * npos + i, offset);
                        evaluator.add_plain(offset, encoder.encode((double)i));
                    }
            }
        }
        // the fresh encryptions the series start from (:54, :134), in the reference's order: sin then cos
        // per (i, j); a deterministic Encryptor yields the same ciphertext every time
        if (encryptor.deterministic()) {
            encryptor.encrypt(encoder.encode(0.0), sin_f, E);
            cos_f = sin_f;
        } else {
            sin_f = CiphertextBatch(ctx, E, 2);
            cos_f = CiphertextBatch(ctx, E, 2);
            Ciphertext z;
            for (size_t i = 0; i < npos; ++i)
                for (int j = 1; j <= degree; ++j) {
                    const size_t e = (size_t)(j - 1) * npos + i;
                    encryptor.encrypt(encoder.encode(0.0), z);
                    sin_f.set(e, z);
                    encryptor.encrypt(encoder.encode(0.0), z);
                    cos_f.set(e, z);
                }
        }
        CiphertextBatch sin_args(ctx, (size_t)degree, 2);
        for (int j = 1; j <= degree; ++j) {
            const double arg_factor = ((float)j) * M_PI / ((double)order);
            Ciphertext sa(b);
            evaluator.multiply_plain(sa, encoder.encode(arg_factor));     // the same ciphertext for every position
            sin_args.set((size_t)(j - 1), sa);
            CiphertextBatch cj = cos_arg.slice((size_t)(j - 1) * npos, npos);
            evaluator.multiply_plain(cj, encoder.encode(arg_factor));     // in place on the slice
        }
        {
            const CiphertextBatch sin_terms = detail::taylor_terms(sin_args, 1.0, evaluator, encoder);
            CiphertextBatch rep(ctx, E, sin_terms.size());
            for (int j = 1; j <= degree; ++j) rep.fill((size_t)(j - 1) * npos, npos, sin_terms.element((size_t)(j - 1)));
            evaluator.add_plain(sin_f, encoder.encode(-1.0));              // start + (-1), per element
            evaluator.add(rep, sin_f);                                     // + the shared terms
            sin_f = std::move(rep);
        }
        taylor_series(cos_arg, cos_f, -1.0, evaluator, encoder);
        CiphertextBatch term = evaluator.multiply(OperandForm(ctx, sin_f), OperandForm(ctx, cos_f));   // sin_factor * cos_factor
        for (int j = 1; j <= degree; ++j) {
            CiphertextBatch tj = term.slice((size_t)(j - 1) * npos, npos);
            evaluator.multiply_plain(tj, encoder.encode(2.0 / (M_PI * ((float)j))));
            evaluator.add(c, tj);
        }
    }
    // c * amplitude: one operand form of the amplitude serves every position
    CiphertextBatch amp(ctx, 1, amplitude.size());
    amp.set(0, amplitude);
    const CiphertextBatch out = evaluator.multiply(OperandForm(ctx, c), OperandForm(ctx, amp), 1, (uint32_t)npos);
    for (size_t i = 0; i < npos; ++i) run.push_back(out.get(i));
}

}  // namespace seal_b200
