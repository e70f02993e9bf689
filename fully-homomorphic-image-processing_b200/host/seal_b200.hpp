// seal_b200.hpp -- seal::-shaped C++ value types over the C-ABI of libfhe_b200.so.
//
// The reference's image-op headers (homo/fhe_image.h, fhe_resize.h, fhe_decode.h) are written
// against seal::Evaluator / Ciphertext / Plaintext / FractionalEncoder.  This
// header gives the same names, argument meaning and error behaviour for the subset those
// headers use on the server side, with the ciphertext data resident in B200 HBM:
//
//   seal::EncryptionParameters / SEALContext   server_jpeg.cpp:74-80   -> Context
//   seal::FractionalEncoder(t, poly, 100, 100, 2)   server_jpeg.cpp:100 -> FractionalEncoder
//   seal::Ciphertext (value type, copied freely)    fhe_image.h:207     -> Ciphertext
//   seal::Evaluator::{add, sub, negate, multiply_plain, add_plain, sub_plain}  (in place on the
//   first argument, second const; std::invalid_argument on mismatch or zero plaintext)
//                                                   fhe_image.h:207-325 -> Evaluator
//   seal::Encryptor (only passed through by fhe_image.h)                -> Encryptor (opaque)
//
// Everything computes on the device through include/fhe_b200.h; there is no CPU path.
// Host language: C++17, no dependency besides the C-ABI header.
#pragma once
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/fhe_b200.h"

namespace seal_b200 {

inline void check(int rc)
{
    if (rc == FHE_OK) return;
    const std::string msg = fhe_last_error();
    if (rc == FHE_ERR_INVALID_ARGUMENT || rc == FHE_ERR_ZERO_PLAIN) throw std::invalid_argument(msg);
    throw std::runtime_error(msg);
}

// EncryptionParameters + SEALContext: ring degree n ("--cmod" in the reference, a naming quirk,
// server_jpeg.cpp:76), coeff_modulus_128(n) unless a chain is given, plain_modulus t.
class Context {
public:
    // cuda_stream: a cudaStream_t the context should work on (null: its own)
    explicit Context(uint32_t n, uint64_t plain_modulus = 1 << 14, std::vector<uint64_t> coeff_modulus = {},
                     int device = 0, void *cuda_stream = nullptr)
        : n_(n), t_(plain_modulus), q_(std::move(coeff_modulus))
    {
        if (q_.empty()) {
            uint64_t buf[16];
            uint32_t k = 0;
            check(fhe_default_coeff_modulus(n, buf, &k));
            q_.assign(buf, buf + k);
        }
        check(fhe_ctx_create(n, (uint32_t)q_.size(), q_.data(), t_, device, cuda_stream, &h_));
    }
    ~Context() { fhe_ctx_destroy(h_); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;

    fhe_ctx *handle() const { return h_; }
    uint32_t poly_degree() const { return n_; }
    uint32_t coeff_mod_count() const { return (uint32_t)q_.size(); }
    const std::vector<uint64_t> &coeff_modulus() const { return q_; }
    uint64_t plain_modulus() const { return t_; }
    size_t ct_words(uint32_t size = 2) const { return (size_t)size * q_.size() * n_; }
    void sync() const { check(fhe_ctx_sync(h_)); }

private:
    uint32_t n_;
    uint64_t t_;
    std::vector<uint64_t> q_;
    fhe_ctx *h_ = nullptr;
};

// seal::Plaintext: n coefficients in [0, t).
class Plaintext {
public:
    Plaintext() = default;
    explicit Plaintext(std::vector<uint64_t> c) : coeffs_(std::move(c)) {}
    const std::vector<uint64_t> &coeffs() const { return coeffs_; }
    bool is_zero() const
    {
        for (uint64_t v : coeffs_)
            if (v) return false;
        return true;
    }

private:
    std::vector<uint64_t> coeffs_;
};

// seal::FractionalEncoder (binary base; the reference's --base override is dead code).
class FractionalEncoder {
public:
    FractionalEncoder(const Context &ctx, uint32_t integer_coeff_count = 100, uint32_t fraction_coeff_count = 100,
                      uint32_t base = 2)
        : n_(ctx.poly_degree()), t_(ctx.plain_modulus()), ic_(integer_coeff_count), fc_(fraction_coeff_count),
          base_(base)
    {
    }
    Plaintext encode(double value) const
    {
        std::vector<uint64_t> c(n_);
        check(fhe_frac_encode(n_, t_, value, ic_, fc_, base_, c.data()));
        return Plaintext(std::move(c));
    }
    // everything an encoding depends on besides the value (cache keys of pre-transformed plaintext sets)
    std::string key() const
    {
        return std::to_string(n_) + "/" + std::to_string(t_) + "/" + std::to_string(ic_) + "/" + std::to_string(fc_) + "/" + std::to_string(base_);
    }

private:
    uint32_t n_;
    uint64_t t_;
    uint32_t ic_, fc_, base_;
};

// Device memory shared by the ciphertexts carved out of it (a block of 64 loaded together stays
// contiguous, so the fused kernels can run on it in place).
class Arena {
public:
    Arena(const Context &ctx, size_t words) : ctx_(&ctx), words_(words)
    {
        void *p = nullptr;
        check(fhe_device_alloc(ctx.handle(), words * sizeof(uint64_t), &p));
        d_ = (uint64_t *)p;
    }
    ~Arena() { fhe_device_free(ctx_->handle(), d_); }
    Arena(const Arena &) = delete;
    Arena &operator=(const Arena &) = delete;
    uint64_t *data() const { return d_; }
    size_t words() const { return words_; }

private:
    const Context *ctx_;
    uint64_t *d_ = nullptr;
    size_t words_;
};

// seal::Ciphertext: a value type.  Copy construction / assignment duplicate the device data
// (the reference copies ciphertexts ~1500 times per block, fhe_image.h:207-284).
class Ciphertext {
public:
    Ciphertext() = default;
    Ciphertext(const Context &ctx, uint32_t size = 2) : ctx_(&ctx), size_(size)
    {
        arena_ = std::make_shared<Arena>(ctx, ctx.ct_words(size));
        d_ = arena_->data();
    }
    // a view of `size` polys at `d` inside a shared arena (see load_block)
    Ciphertext(const Context &ctx, std::shared_ptr<Arena> arena, uint64_t *d, uint32_t size)
        : ctx_(&ctx), arena_(std::move(arena)), d_(d), size_(size), view_(true)
    {
    }
    Ciphertext(const Ciphertext &o) : ctx_(o.ctx_), size_(o.size_)
    {
        if (!o.d_) return;
        arena_ = std::make_shared<Arena>(*ctx_, ctx_->ct_words(size_));
        d_ = arena_->data();
        check(fhe_device_copy(ctx_->handle(), d_, o.d_, ctx_->ct_words(size_) * sizeof(uint64_t)));
    }
    Ciphertext(Ciphertext &&o) noexcept : ctx_(o.ctx_), arena_(std::move(o.arena_)), d_(o.d_), size_(o.size_), view_(o.view_)
    {
        o.d_ = nullptr;
    }
    // Assigning into a view (an element of a block loaded by load_block) overwrites the view's
    // storage, so `data[i] = tmp` keeps the block contiguous; otherwise this is a deep copy.
    Ciphertext &operator=(const Ciphertext &o)
    {
        if (this == &o) return *this;
        if (!o.d_) throw std::invalid_argument("assignment from an empty ciphertext");
        const bool reuse = d_ && ctx_ == o.ctx_ && size_ == o.size_;
        if (!reuse) {
            if (view_) throw std::invalid_argument("cannot resize a ciphertext that lives inside a block");
            ctx_ = o.ctx_;
            size_ = o.size_;
            arena_ = std::make_shared<Arena>(*ctx_, ctx_->ct_words(size_));
            d_ = arena_->data();
        }
        check(fhe_device_copy(ctx_->handle(), d_, o.d_, ctx_->ct_words(size_) * sizeof(uint64_t)));
        return *this;
    }
    Ciphertext &operator=(Ciphertext &&o)
    {
        if (this == &o) return *this;
        if (view_ || o.view_) return *this = static_cast<const Ciphertext &>(o);
        ctx_ = o.ctx_; arena_ = std::move(o.arena_); d_ = o.d_; size_ = o.size_;
        o.d_ = nullptr;
        return *this;
    }

    uint32_t size() const { return size_; }
    uint64_t *device_data() const { return d_; }
    const Context *context() const { return ctx_; }
    bool valid() const { return d_ != nullptr; }

    // host <-> device, layout [poly][limb][n] (drop SEAL 2.3's padding coefficient when importing)
    void upload(const uint64_t *host) { check(fhe_upload(ctx_->handle(), d_, host, ctx_->ct_words(size_) * 8)); }
    void download(uint64_t *host) const { check(fhe_download(ctx_->handle(), host, d_, ctx_->ct_words(size_) * 8)); }

private:
    const Context *ctx_ = nullptr;
    std::shared_ptr<Arena> arena_;
    uint64_t *d_ = nullptr;
    uint32_t size_ = 0;
    bool view_ = false;
};

// `count` size-2 ciphertexts in one contiguous device allocation (what server_jpeg.cpp holds after
// loading the 64 ciphertexts of a block, :115-124).  host may be null (uninitialised storage).
inline std::vector<Ciphertext> load_block(const Context &ctx, const uint64_t *host, size_t count, uint32_t size = 2)
{
    const size_t w = ctx.ct_words(size);
    auto arena = std::make_shared<Arena>(ctx, w * count);
    if (host) check(fhe_upload(ctx.handle(), arena->data(), host, w * count * sizeof(uint64_t)));
    std::vector<Ciphertext> v;
    v.reserve(count);
    for (size_t i = 0; i < count; ++i) v.emplace_back(ctx, arena, arena->data() + i * w, size);
    return v;
}
// true when the ciphertexts are consecutive in one allocation, in order
inline bool is_contiguous(const std::vector<Ciphertext> &v)
{
    if (v.empty() || !v[0].valid()) return false;
    const size_t w = v[0].context()->ct_words(v[0].size());
    for (size_t i = 0; i < v.size(); ++i)
        if (!v[i].valid() || v[i].size() != v[0].size() || v[i].device_data() != v[0].device_data() + i * w) return false;
    return true;
}

// `count` independent ciphertexts of equal size in one allocation, [ct][size][k][n]: the unit the
// device works on efficiently.  The reference evaluates its graphs one seal::Ciphertext at a time
// (one output pixel, one 4x4 neighbourhood); the graphs in fhe_resize_b200.hpp / fhe_decode_b200.hpp
// are templates that run unchanged on a Ciphertext or on a CiphertextBatch -- every Evaluator call
// then covers the whole batch (a row of output pixels) with the same launches.
class CiphertextBatch {
public:
    CiphertextBatch() = default;
    CiphertextBatch(const Context &ctx, size_t count, uint32_t size = 2) : ctx_(&ctx), count_(count), size_(size)
    {
        arena_ = std::make_shared<Arena>(ctx, ctx.ct_words(size) * count);
    }
    CiphertextBatch(const CiphertextBatch &o) : ctx_(o.ctx_), count_(o.count_), size_(o.size_)
    {
        if (!o.arena_) return;  // copying a slice makes an independent batch of just that range
        const size_t words = ctx_->ct_words(size_) * count_;
        arena_ = std::make_shared<Arena>(*ctx_, words);
        check(fhe_device_copy(ctx_->handle(), arena_->data(), o.device_data(), words * sizeof(uint64_t)));
    }
    CiphertextBatch(CiphertextBatch &&) noexcept = default;
    CiphertextBatch &operator=(const CiphertextBatch &o)
    {
        if (this != &o) {
            CiphertextBatch t(o);
            *this = std::move(t);
        }
        return *this;
    }
    CiphertextBatch &operator=(CiphertextBatch &&) noexcept = default;

    size_t count() const { return count_; }
    uint32_t size() const { return size_; }
    bool valid() const { return arena_ != nullptr; }
    const Context *context() const { return ctx_; }
    uint64_t *device_data() const { return arena_ ? arena_->data() + off_ : nullptr; }
    uint64_t *element(size_t i) const { return device_data() + i * ctx_->ct_words(size_); }
    // elements [first, first + count) as a batch SHARING this storage (in-place Evaluator ops on the
    // slice -- multiply_plain, add, negate ... -- act on this batch; size-changing ops detach it)
    CiphertextBatch slice(size_t first, size_t count) const
    {
        if (first + count > count_) throw std::invalid_argument("batch slice out of range");
        CiphertextBatch v;
        v.ctx_ = ctx_;
        v.arena_ = arena_;
        v.count_ = count;
        v.size_ = size_;
        v.off_ = off_ + first * ctx_->ct_words(size_);
        return v;
    }
    // elements [first, first + count) <- copies of the `src_size`-polynomial ciphertext at d_src (one launch)
    void fill(size_t first, size_t count, const uint64_t *d_src)
    {
        if (first + count > count_) throw std::invalid_argument("batch range out of range");
        check(fhe_device_broadcast(ctx_->handle(), element(first), d_src, size_, count));
    }
    // element i <- c (device copy); sizes must agree
    void set(size_t i, const Ciphertext &c)
    {
        if (i >= count_ || !c.valid() || c.size() != size_ || c.context() != ctx_) throw std::invalid_argument("batch element mismatch");
        check(fhe_device_copy(ctx_->handle(), element(i), c.device_data(), ctx_->ct_words(size_) * sizeof(uint64_t)));
    }
    Ciphertext get(size_t i) const
    {
        if (i >= count_) throw std::invalid_argument("batch index out of range");
        Ciphertext c(*ctx_, size_);
        check(fhe_device_copy(ctx_->handle(), c.device_data(), element(i), ctx_->ct_words(size_) * sizeof(uint64_t)));
        return c;
    }
    // the same ciphertexts in storage for `new_size` >= size() polynomials, missing ones zero
    CiphertextBatch widened(uint32_t new_size) const
    {
        if (new_size < size_) throw std::invalid_argument("cannot shrink a ciphertext");
        CiphertextBatch r(*ctx_, count_, new_size);
        const size_t wo = ctx_->ct_words(size_) * sizeof(uint64_t), wn = ctx_->ct_words(new_size) * sizeof(uint64_t);
        check(fhe_device_zero(ctx_->handle(), r.device_data(), wn * count_));
        check(fhe_device_copy_2d(ctx_->handle(), r.device_data(), wn, device_data(), wo, wo, count_));
        return r;
    }
    void upload(const uint64_t *host) { check(fhe_upload(ctx_->handle(), device_data(), host, ctx_->ct_words(size_) * count_ * 8)); }
    void download(uint64_t *host) const { check(fhe_download(ctx_->handle(), host, device_data(), ctx_->ct_words(size_) * count_ * 8)); }

private:
    const Context *ctx_ = nullptr;
    std::shared_ptr<Arena> arena_;
    size_t count_ = 0, off_ = 0;
    uint32_t size_ = 0;
};

// The operand half of Evaluator::multiply done once (include/fhe_b200.h fhe_form_create): the BEHZ
// extension and the forward transforms of `count` ciphertexts.  Immutable; a graph that multiplies by
// the same ciphertexts repeatedly (the powers of the fractions in Cubic, shifted_x in the Taylor
// series) builds the form once and passes it to fhe_multiply_form / fhe_cubic / fhe_linear.
class OperandForm {
public:
    OperandForm(const Context &ctx, const CiphertextBatch &b) : count_(b.count()), size_(b.size())
    {
        if (!b.valid() || b.context() != &ctx) throw std::invalid_argument("encrypted is not valid for encryption parameters");
        check(fhe_form_create(ctx.handle(), b.device_data(), b.size(), b.count(), &h_));
    }
    ~OperandForm() { fhe_form_destroy(h_); }
    OperandForm(const OperandForm &) = delete;
    OperandForm &operator=(const OperandForm &) = delete;
    const fhe_form *handle() const { return h_; }
    size_t count() const { return count_; }
    uint32_t size() const { return size_; }

private:
    fhe_form *h_ = nullptr;
    size_t count_;
    uint32_t size_;
};

// seal::Encryptor.  fhe_image.h only passes it through; fhe_resize.h (:230-266) and fhe_decode.h
// (:54,134) encrypt PUBLIC scalars on the server with it.  Two server-side forms:
//   * Encryptor(fn): a callback that produces the ciphertext (a real public-key encryption made by
//     whoever holds the public key -- the parity tests feed the oracle's encryptions through it);
//   * Encryptor(ctx): the deterministic transparent encryption (Delta*m, 0), which needs no key and
//     is a valid BFV ciphertext of a public value.
// A default-constructed Encryptor throws on use, like a seal::Encryptor without a public key.
class Encryptor {
public:
    using Fn = std::function<void(const Plaintext &, Ciphertext &)>;
    Encryptor() = default;
    explicit Encryptor(Fn fn) : fn_(std::move(fn)) {}
    explicit Encryptor(const Context &ctx) : ctx_(&ctx) {}
    void encrypt(const Plaintext &plain, Ciphertext &destination) const;
    // true: equal plaintexts give equal ciphertexts (the transparent form), so a graph may share them
    bool deterministic() const { return !fn_ && ctx_; }
    // `count` fresh encryptions of the same plaintext (one callback invocation each)
    void encrypt(const Plaintext &plain, CiphertextBatch &destination, size_t count) const
    {
        if (!fn_ && !ctx_) throw std::logic_error("this Encryptor holds no public key and no context");
        Ciphertext c;
        if (deterministic() && count) {  // equal ciphertexts: one encryption, one broadcast
            encrypt(plain, c);
            destination = CiphertextBatch(*c.context(), count, c.size());
            destination.fill(0, count, c.device_data());
            return;
        }
        for (size_t i = 0; i < count; ++i) {
            if (fn_ || i == 0) encrypt(plain, c);
            if (i == 0) destination = CiphertextBatch(*c.context(), count, c.size());
            destination.set(i, c);
        }
    }

private:
    Fn fn_;
    const Context *ctx_ = nullptr;
};

// seal::EvaluationKeys of KeyGenerator::generate_evaluation_keys(dbc, count)
// (tests/parameters.cpp:89-90), coefficient form [count][k][digits][2][k][n] (include/fhe_b200.h).
class EvaluationKeys {
public:
    EvaluationKeys(const Context &ctx, const uint64_t *keys, uint32_t decomposition_bit_count, uint32_t count)
        : dbc_(decomposition_bit_count), count_(count)
    {
        check(fhe_evk_create(ctx.handle(), keys, dbc_, count_, &h_));
    }
    ~EvaluationKeys() { fhe_evk_destroy(h_); }
    EvaluationKeys(const EvaluationKeys &) = delete;
    EvaluationKeys &operator=(const EvaluationKeys &) = delete;
    const fhe_evk *handle() const { return h_; }
    uint32_t decomposition_bit_count() const { return dbc_; }
    uint32_t size() const { return count_; }

private:
    fhe_evk *h_ = nullptr;
    uint32_t dbc_, count_;
};

// seal::Evaluator.  Plaintext operands are transformed once and cached by content (the reference
// re-encodes and re-transforms the same 13 + 64 + 9 constants on every call).
class Evaluator {
public:
    explicit Evaluator(const Context &ctx) : ctx_(&ctx) {}
    ~Evaluator()
    {
        for (auto &kv : cache_) fhe_ptab_destroy(kv.second);
    }
    Evaluator(const Evaluator &) = delete;
    Evaluator &operator=(const Evaluator &) = delete;

    void add(Ciphertext &a, const Ciphertext &b) { binary(a, b, fhe_add); }
    void sub(Ciphertext &a, const Ciphertext &b) { binary(a, b, fhe_sub); }
    void negate(Ciphertext &a)
    {
        require(a);
        check(fhe_negate(ctx_->handle(), a.device_data(), 1, a.size()));
    }
    void multiply_plain(Ciphertext &a, const Plaintext &p)
    {
        require(a);
        if (p.is_zero()) throw std::invalid_argument("plain cannot be zero");
        check(fhe_multiply_plain(ctx_->handle(), a.device_data(), 1, a.size(), table(p), 0));
    }
    void add_plain(Ciphertext &a, const Plaintext &p)
    {
        require(a);
        check(fhe_add_plain(ctx_->handle(), a.device_data(), 1, a.size(), table(p), 0));
    }
    void sub_plain(Ciphertext &a, const Plaintext &p)
    {
        require(a);
        check(fhe_sub_plain(ctx_->handle(), a.device_data(), 1, a.size(), table(p), 0));
    }
    // Evaluator::multiply(a, b): a <- a * b, size a.size() + b.size() - 1 (fhe_resize.h:175-179,
    // 197-198; fhe_decode.h:71-97,235,239).  BEHZ full-RNS scaling on the device.
    void multiply(Ciphertext &a, const Ciphertext &b)
    {
        require(a);
        require(b);
        Ciphertext dest(*ctx_, a.size() + b.size() - 1);
        check(fhe_multiply(ctx_->handle(), a.device_data(), a.size(), b.device_data(), b.size(), dest.device_data(), 1));
        a = std::move(dest);
        if (auto_relin_) relinearize(a, *auto_relin_);
    }
    // Evaluator::square(a) (fhe_resize.h:174; fhe_decode.h:67 ...)
    void square(Ciphertext &a)
    {
        require(a);
        Ciphertext dest(*ctx_, 2 * a.size() - 1);
        check(fhe_square(ctx_->handle(), a.device_data(), a.size(), dest.device_data(), 1));
        a = std::move(dest);
        if (auto_relin_) relinearize(a, *auto_relin_);
    }
    // Relinearisation policy (an EXTENSION: homo/ never relinearises, so its ciphertexts grow to
    // size 6 in the resize graphs and 22 in the decode graphs).  With keys set, every ct x ct
    // product (multiply, square; single and batched) is brought back to size 2 immediately -- the
    // insertion point BASELINE config 5 ("deep-multiply + relinearize") asks for -- so the op
    // libraries of fhe_resize_b200.hpp / fhe_decode_b200.hpp run unchanged on size-2 operands.
    // The keys need evk.size() >= 1 (only size 3 -> 2 occurs).  nullptr restores the reference's
    // behaviour.  Results decrypt to the same plaintext but are different ciphertexts.
    void set_relinearization(const EvaluationKeys *evk) { auto_relin_ = evk; }
    const EvaluationKeys *relinearization() const { return auto_relin_; }
    // Evaluator::relinearize(a, evk): back to size 2 (north_star; no call site in homo/)
    void relinearize(Ciphertext &a, const EvaluationKeys &evk)
    {
        require(a);
        if (a.size() <= 2) return;
        Ciphertext dest(*ctx_, 2);
        check(fhe_relinearize(ctx_->handle(), a.device_data(), a.size(), evk.handle(), dest.device_data(), 1));
        a = std::move(dest);
    }
    // ---- the same operations on a CiphertextBatch: one call, `count` independent ciphertexts ----
    void add(CiphertextBatch &a, const CiphertextBatch &b) { binary_batch(a, b, fhe_add); }
    void sub(CiphertextBatch &a, const CiphertextBatch &b) { binary_batch(a, b, fhe_sub); }
    void negate(CiphertextBatch &a)
    {
        require(a);
        check(fhe_negate(ctx_->handle(), a.device_data(), a.count(), a.size()));
    }
    void multiply_plain(CiphertextBatch &a, const Plaintext &p)
    {
        require(a);
        if (p.is_zero()) throw std::invalid_argument("plain cannot be zero");
        check(fhe_multiply_plain(ctx_->handle(), a.device_data(), a.count(), a.size(), table(p), 0));
    }
    void add_plain(CiphertextBatch &a, const Plaintext &p)
    {
        require(a);
        check(fhe_add_plain(ctx_->handle(), a.device_data(), a.count(), a.size(), table(p), 0));
    }
    void sub_plain(CiphertextBatch &a, const Plaintext &p)
    {
        require(a);
        check(fhe_sub_plain(ctx_->handle(), a.device_data(), a.count(), a.size(), table(p), 0));
    }
    // add_plain with a different short plaintext per element: plains[i * ncoef + c] = coefficient c of the
    // plaintext of element i (all higher coefficients zero)
    void add_plain_each(CiphertextBatch &a, const std::vector<uint64_t> &plains, uint32_t ncoef)
    {
        require(a);
        if (plains.size() != a.count() * (size_t)ncoef) throw std::invalid_argument("one short plaintext per element expected");
        check(fhe_add_plain_each(ctx_->handle(), a.device_data(), a.count(), a.size(), plains.data(), ncoef));
    }
    void multiply(CiphertextBatch &a, const CiphertextBatch &b)
    {
        require(a);
        require(b);
        if (a.count() != b.count()) throw std::invalid_argument("batch counts differ");
        CiphertextBatch dest(*ctx_, a.count(), a.size() + b.size() - 1);
        check(fhe_multiply(ctx_->handle(), a.device_data(), a.size(), b.device_data(), b.size(), dest.device_data(), a.count()));
        a = std::move(dest);
        if (auto_relin_) relinearize(a, *auto_relin_);
    }
    void square(CiphertextBatch &a)
    {
        require(a);
        CiphertextBatch dest(*ctx_, a.count(), 2 * a.size() - 1);
        check(fhe_square(ctx_->handle(), a.device_data(), a.size(), dest.device_data(), a.count()));
        a = std::move(dest);
        if (auto_relin_) relinearize(a, *auto_relin_);
    }
    // product i = ciphertext i / a_div of `a` times ciphertext i / b_div of `b`, from operand forms
    // (fhe_multiply_form): `n_ct` products (default: one per ciphertext of `a`); a form times itself is
    // Evaluator::square.  Follows the relinearisation policy like multiply.
    CiphertextBatch multiply(const OperandForm &a, const OperandForm &b, uint32_t a_div = 1, uint32_t b_div = 1, size_t n_ct = 0)
    {
        if (!n_ct) n_ct = a.count() * a_div;
        CiphertextBatch dest(*ctx_, n_ct, a.size() + b.size() - 1);
        check(fhe_multiply_form(ctx_->handle(), a.handle(), a_div, b.handle(), b_div, dest.device_data(), n_ct));
        if (auto_relin_) relinearize(dest, *auto_relin_);
        return dest;
    }
    void relinearize(CiphertextBatch &a, const EvaluationKeys &evk)
    {
        require(a);
        if (a.size() <= 2) return;
        CiphertextBatch dest(*ctx_, a.count(), 2);
        check(fhe_relinearize(ctx_->handle(), a.device_data(), a.size(), evk.handle(), dest.device_data(), a.count()));
        a = std::move(dest);
    }
    const Context &context() const { return *ctx_; }
    // per-evaluator cache slot for the fused-graph plans of fhe_image_b200.hpp
    std::shared_ptr<void> &plan_slot(const std::string &key) { return plans_[key]; }

private:
    void require(const Ciphertext &a) const
    {
        if (!a.valid() || a.context() != ctx_) throw std::invalid_argument("encrypted is not valid for encryption parameters");
    }
    void require(const CiphertextBatch &a) const
    {
        if (!a.valid() || a.context() != ctx_) throw std::invalid_argument("encrypted is not valid for encryption parameters");
    }
    // unequal sizes: the smaller operand is re-strided into zero-padded storage of the larger size,
    // then the equal-size kernel runs (exactly SEAL's "missing polynomials are zero")
    template <class F>
    void binary_batch(CiphertextBatch &a, const CiphertextBatch &b, F fn)
    {
        require(a);
        require(b);
        if (a.count() != b.count()) throw std::invalid_argument("batch counts differ");
        if (a.size() < b.size()) a = a.widened(b.size());
        if (b.size() < a.size()) {
            const CiphertextBatch bw = b.widened(a.size());
            check(fn(ctx_->handle(), a.device_data(), bw.device_data(), a.count(), a.size()));
            return;  // bw is released in stream order (fhe_device_free)
        }
        check(fn(ctx_->handle(), a.device_data(), b.device_data(), a.count(), a.size()));
    }
    template <class F>
    void binary(Ciphertext &a, const Ciphertext &b, F fn)
    {
        require(a);
        require(b);
        // SEAL allows unequal sizes: the result has the larger size, missing polynomials count as
        // zero (fhe_resize.h:182-185 adds sizes 4, 3 and 2).
        if (b.size() > a.size()) {
            const size_t wa = ctx_->ct_words(a.size()), wb = ctx_->ct_words(b.size());
            Ciphertext r(*ctx_, b.size());
            check(fhe_device_copy(ctx_->handle(), r.device_data(), a.device_data(), wa * sizeof(uint64_t)));
            check(fhe_device_copy(ctx_->handle(), r.device_data() + wa, b.device_data() + wa, (wb - wa) * sizeof(uint64_t)));
            if ((void *)fn == (void *)fhe_sub) check(fhe_negate(ctx_->handle(), r.device_data() + wa, 1, b.size() - a.size()));
            check(fn(ctx_->handle(), r.device_data(), b.device_data(), 1, a.size()));
            a = std::move(r);
            return;
        }
        check(fn(ctx_->handle(), a.device_data(), b.device_data(), 1, b.size()));
    }
    fhe_ptab *table(const Plaintext &p)
    {
        if (p.coeffs().size() != ctx_->poly_degree()) throw std::invalid_argument("plain is not valid for encryption parameters");
        std::string key((const char *)p.coeffs().data(), p.coeffs().size() * sizeof(uint64_t));
        auto it = cache_.find(key);
        if (it != cache_.end()) return it->second;
        fhe_ptab *t = nullptr;
        check(fhe_ptab_create(ctx_->handle(), p.coeffs().data(), 1, &t));
        cache_.emplace(std::move(key), t);
        return t;
    }
    const Context *ctx_;
    const EvaluationKeys *auto_relin_ = nullptr;
    std::unordered_map<std::string, fhe_ptab *> cache_;
    std::unordered_map<std::string, std::shared_ptr<void>> plans_;
};

inline void Encryptor::encrypt(const Plaintext &plain, Ciphertext &destination) const
{
    if (fn_) {
        fn_(plain, destination);
        return;
    }
    if (!ctx_) throw std::logic_error("this Encryptor holds no public key and no context");
    if (plain.coeffs().size() != ctx_->poly_degree()) throw std::invalid_argument("plain is not valid for encryption parameters");
    Ciphertext z(*ctx_, 2);
    if (plain.is_zero()) check(fhe_device_zero(ctx_->handle(), z.device_data(), ctx_->ct_words(2) * sizeof(uint64_t)));
    else check(fhe_encrypt_transparent(ctx_->handle(), plain.coeffs().data(), 1, z.device_data()));
    destination = std::move(z);
}

}  // namespace seal_b200
