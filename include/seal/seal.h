// seal/seal.h -- source-compatibility header: the subset of the SEAL 2.3 C++ API that the reference's
// server side is written against, implemented on libfhe_b200 (B200, sm_100a).
//
// With `-I <this repo>/include` in front of SEAL's own include path, the reference's image-op headers
// and server programs compile UNCHANGED and run their ciphertext arithmetic on the GPU:
//
//   homo/fhe_image.h   (encrypted_dct :196-288, quantize_fhe :294-305, rgb_to_ycc_fhe :310-325)
//   homo/fhe_resize.h  (Cubic :143-189, Linear :191-204, Sample* :222-305, ResizeImage :308-392)
//   homo/fhe_decode.h  (homomorphic_sin :48-120, homomorphic_cos :128-200, approximated_step :202-242)
//   homo/server_jpeg.cpp, homo/server_resize.cpp, homo/server_decode.cpp
//
// `make -C oracle ref` does exactly that with the sources where they lie under /root/reference and
// tests/test_reference_sources_on_gpu.py runs the resulting programs against the oracle.
//
// What is here: EncryptionParameters / SEALContext / coeff_modulus_128 / SmallModulus / BigPoly /
// BigUInt (the accessors print_parameters uses, fhe_image.h:493-506), Plaintext, Ciphertext (value
// type; load / save in the wire format of host/seal_wire_b200.hpp), FractionalEncoder (encode /
// decode), Evaluator (in-place add, sub, negate, multiply_plain, add_plain, sub_plain, multiply,
// square, relinearize + the three-argument forms), EvaluationKeys, PublicKey / SecretKey (load /
// save), Encryptor, Decryptor.
//
// Deliberate differences from SEAL, all on the key-holding side the reference itself calls client
// side (README.md:9):
//   * Encryptor::encrypt (the server encrypts PUBLIC scalars with it: fhe_resize.h:230-266,
//     fhe_decode.h:54,134, server_decode.cpp:121-124) produces the transparent encryption
//     (Delta * m, 0): a valid BFV ciphertext of a public value, deterministic, no key needed;
//   * Decryptor (the servers construct one "FOR DEBUGGING ONLY", server_jpeg.cpp:96-97; only the
//     decrypting overload of approximated_step uses it, fhe_decode.h:244-282) throws
//     std::logic_error on use: secret-key operations are not part of the server library;
//   * KeyGenerator, IntegerEncoder, PolyCRTBuilder, Chooser*, Galois: never used by homo/ servers.
//   * Ciphertext::load needs the ring: it binds to the live SEALContext whose (n, k) match the record.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "../../fully-homomorphic-image-processing_b200/host/seal_b200.hpp"
#include "../../fully-homomorphic-image-processing_b200/host/seal_wire_b200.hpp"

namespace seal {

// SEAL 2.3's own headers make namespace std visible to whoever says `using namespace seal`, and the
// reference depends on it (`chrono::duration<double, milli>` at fhe_image.h:286, fhe_resize.h:187).
using namespace std;

// ---- parameters -----------------------------------------------------------------------------
class SmallModulus {
public:
    SmallModulus(std::uint64_t value = 0) : value_(value) {}
    std::uint64_t value() const { return value_; }
    int bit_count() const { return value_ ? 64 - __builtin_clzll(value_) : 0; }
    bool is_zero() const { return value_ == 0; }
    bool operator==(const SmallModulus &o) const { return value_ == o.value_; }
    bool operator!=(const SmallModulus &o) const { return value_ != o.value_; }

private:
    std::uint64_t value_;
};

// the polynomial modulus x^n + 1 (SEAL's BigPoly; only this shape is valid for the scheme)
class BigPoly {
public:
    BigPoly() = default;
    BigPoly(const std::string &s) { parse(s); }
    BigPoly(const char *s) { parse(s); }
    int coeff_count() const { return n_ ? (int)n_ + 1 : 0; }
    int significant_coeff_count() const { return coeff_count(); }
    std::uint32_t degree() const { return n_; }
    std::string to_string() const { return n_ ? "1x^" + std::to_string(n_) + " + 1" : "0"; }

private:
    void parse(const std::string &s)
    {
        unsigned n = 0;
        if (std::sscanf(s.c_str(), "1x^%u + 1", &n) != 1 || n < 2 || (n & (n - 1)))
            throw std::invalid_argument("poly_modulus must be 1x^(power of two) + 1");
        n_ = n;
    }
    std::uint32_t n_ = 0;
};

// unsigned big integer, little-endian words (SEAL's BigUInt; here only the product of the primes)
class BigUInt {
public:
    BigUInt() = default;
    explicit BigUInt(std::vector<std::uint64_t> words) : w_(std::move(words)) {}
    int significant_bit_count() const
    {
        for (size_t i = w_.size(); i-- > 0;)
            if (w_[i]) return (int)(i * 64) + 64 - __builtin_clzll(w_[i]);
        return 0;
    }
    const std::vector<std::uint64_t> &words() const { return w_; }

private:
    std::vector<std::uint64_t> w_;
};

// SEAL 2.3 defaults for 128-bit security (fhe_b200.h: fhe_default_coeff_modulus)
inline std::vector<SmallModulus> coeff_modulus_128(int poly_modulus_degree)
{
    std::uint64_t q[16];
    std::uint32_t k = 0;
    seal_b200::check(fhe_default_coeff_modulus((std::uint32_t)poly_modulus_degree, q, &k));
    return std::vector<SmallModulus>(q, q + k);
}

class EncryptionParameters {
public:
    void set_poly_modulus(const BigPoly &p) { poly_ = p; }
    void set_poly_modulus(const std::string &p) { poly_ = BigPoly(p); }
    void set_poly_modulus(const char *p) { poly_ = BigPoly(p); }
    void set_coeff_modulus(const std::vector<SmallModulus> &q) { coeff_ = q; }
    void set_plain_modulus(std::uint64_t t) { plain_ = SmallModulus(t); }
    void set_plain_modulus(const SmallModulus &t) { plain_ = t; }
    void set_noise_standard_deviation(double s) { sigma_ = s; }
    const BigPoly &poly_modulus() const { return poly_; }
    const std::vector<SmallModulus> &coeff_modulus() const { return coeff_; }
    const SmallModulus &plain_modulus() const { return plain_; }
    double noise_standard_deviation() const { return sigma_; }

private:
    BigPoly poly_;
    std::vector<SmallModulus> coeff_;
    SmallModulus plain_;
    double sigma_ = 3.19;
};

class SEALContext;
namespace detail {
inline std::vector<const SEALContext *> &live_contexts()
{
    static std::vector<const SEALContext *> v;
    return v;
}
inline std::mutex &live_mutex()
{
    static std::mutex m;
    return m;
}
inline seal_b200::hash_block_type &last_hash()
{
    static seal_b200::hash_block_type h{};
    return h;
}
}  // namespace detail

// Validates the parameters and creates the device context (throws when no B200 is usable: there is
// no CPU path).
class SEALContext {
public:
    explicit SEALContext(const EncryptionParameters &parms, int device = 0) : parms_(parms)
    {
        if (!parms.poly_modulus().degree() || parms.coeff_modulus().empty() || parms.plain_modulus().is_zero())
            throw std::invalid_argument("encryption parameters are not set");
        std::vector<std::uint64_t> q;
        for (const SmallModulus &m : parms.coeff_modulus()) q.push_back(m.value());
        ctx_ = std::make_shared<seal_b200::Context>(parms.poly_modulus().degree(), parms.plain_modulus().value(), q, device);
        // product of the primes
        std::vector<std::uint64_t> w(q.size() + 1, 0);
        w[0] = 1;
        for (std::uint64_t m : q) {
            unsigned __int128 carry = 0;
            for (auto &x : w) {
                const unsigned __int128 v = (unsigned __int128)x * m + carry;
                x = (std::uint64_t)v;
                carry = v >> 64;
            }
        }
        total_ = BigUInt(std::move(w));
        std::lock_guard<std::mutex> g(detail::live_mutex());
        detail::live_contexts().push_back(this);
    }
    ~SEALContext()
    {
        std::lock_guard<std::mutex> g(detail::live_mutex());
        auto &v = detail::live_contexts();
        v.erase(std::remove(v.begin(), v.end(), this), v.end());
    }
    SEALContext(const SEALContext &) = delete;
    SEALContext &operator=(const SEALContext &) = delete;

    const EncryptionParameters &parms() const { return parms_; }
    const BigPoly &poly_modulus() const { return parms_.poly_modulus(); }
    const std::vector<SmallModulus> &coeff_modulus() const { return parms_.coeff_modulus(); }
    const SmallModulus &plain_modulus() const { return parms_.plain_modulus(); }
    const BigUInt &total_coeff_modulus() const { return total_; }
    double noise_standard_deviation() const { return parms_.noise_standard_deviation(); }
    bool parameters_set() const { return true; }
    const seal_b200::Context &device_context() const { return *ctx_; }

private:
    EncryptionParameters parms_;
    std::shared_ptr<seal_b200::Context> ctx_;
    BigUInt total_;
};

// ---- values ---------------------------------------------------------------------------------
using Plaintext = seal_b200::Plaintext;

// seal::Ciphertext: a value type whose data lives in HBM.  load / save: host/seal_wire_b200.hpp.
class Ciphertext : public seal_b200::Ciphertext {
public:
    using seal_b200::Ciphertext::Ciphertext;
    Ciphertext() = default;
    Ciphertext(const seal_b200::Ciphertext &c) : seal_b200::Ciphertext(c) {}
    Ciphertext(seal_b200::Ciphertext &&c) : seal_b200::Ciphertext(std::move(c)) {}

    void load(std::istream &stream)
    {
        // the record names (n + 1, k); bind to the live context of that shape (most recent first)
        const std::istream::pos_type at = stream.tellg();
        std::uint64_t hash[4];
        std::int32_t meta[4];
        stream.read((char *)hash, sizeof(hash));
        stream.read((char *)meta, sizeof(meta));
        if (!stream) throw std::invalid_argument("ciphertext stream: truncated header");
        stream.seekg(at);
        const seal_b200::Context *ctx = nullptr;
        {
            std::lock_guard<std::mutex> g(detail::live_mutex());
            const auto &v = detail::live_contexts();
            for (size_t i = v.size(); i-- > 0 && !ctx;) {
                const seal_b200::Context &c = v[i]->device_context();
                if ((std::int32_t)c.poly_degree() + 1 == meta[2] && (std::int32_t)c.coeff_mod_count() == meta[3]) ctx = &c;
            }
        }
        if (!ctx) throw std::invalid_argument("ciphertext is not valid for any live SEALContext");
        seal_b200::WireHeader hdr;
        seal_b200::Ciphertext c;
        if (!seal_b200::load_ciphertext(stream, *ctx, c, hdr)) throw std::invalid_argument("ciphertext stream ended");
        detail::last_hash() = hdr.hash;
        static_cast<seal_b200::Ciphertext &>(*this) = std::move(c);
    }
    void save(std::ostream &stream) const
    {
        if (!valid()) throw std::invalid_argument("cannot save an empty ciphertext");
        seal_b200::save_ciphertext(stream, *this, detail::last_hash());
    }
};

// seal::FractionalEncoder(plain_modulus, poly_modulus, integer_coeff_count, fraction_coeff_count, base)
// (server_jpeg.cpp:100).  Host only.
class FractionalEncoder {
public:
    FractionalEncoder(const SmallModulus &plain_modulus, const BigPoly &poly_modulus, int integer_coeff_count,
                      int fraction_coeff_count, std::uint64_t base = 2)
        : n_(poly_modulus.degree()), t_(plain_modulus.value()), ic_((std::uint32_t)integer_coeff_count),
          fc_((std::uint32_t)fraction_coeff_count), base_((std::uint32_t)base)
    {
        if (integer_coeff_count <= 0 || fraction_coeff_count <= 0 || (std::uint64_t)ic_ + fc_ > n_)
            throw std::invalid_argument("integer / fraction coefficient counts do not fit the polynomial modulus");
    }
    Plaintext encode(double value) const
    {
        std::vector<std::uint64_t> c(n_);
        seal_b200::check(fhe_frac_encode(n_, t_, value, ic_, fc_, base_, c.data()));
        return Plaintext(std::move(c));
    }
    // SEAL's integer overloads (Cubic passes encode(3), approximated_step encode(i)): integer part only
    template <class T, typename std::enable_if<std::is_integral<T>::value, int>::type = 0>
    Plaintext encode(T value) const
    {
        return encode((double)value);
    }
    // integer part: coefficients below n - fraction_coeff_count, weight base^i; fractional part: the top
    // fraction_coeff_count coefficients, coefficient n-1-i has weight -base^-(i+1) (x^n = -1); every
    // coefficient is read as its centred representative mod t
    double decode(const Plaintext &plain) const
    {
        const std::vector<std::uint64_t> &c = plain.coeffs();
        if (c.size() > n_) throw std::invalid_argument("plain is not valid for encryption parameters");
        double acc = 0;
        const double b = (double)base_;
        for (size_t i = 0; i < c.size(); ++i) {
            if (!c[i]) continue;
            if (c[i] >= t_) throw std::invalid_argument("plain is not valid for encryption parameters");
            const double v = c[i] > t_ / 2 ? -(double)(t_ - c[i]) : (double)c[i];
            if (i < n_ - fc_) acc += v * std::pow(b, (double)i);
            else acc -= v * std::pow(b, -(double)(n_ - i));
        }
        return acc;
    }
    std::uint64_t base() const { return base_; }

private:
    std::uint32_t n_;
    std::uint64_t t_;
    std::uint32_t ic_, fc_, base_;
};

// ---- keys -----------------------------------------------------------------------------------
// The servers load both key files (server_jpeg.cpp:84-91) but never need their contents here: the
// bytes are kept verbatim so that save() echoes them.
class KeyBlob {
public:
    void load(std::istream &stream)
    {
        bytes_.assign(std::istreambuf_iterator<char>(stream), std::istreambuf_iterator<char>());
    }
    void save(std::ostream &stream) const { stream.write(bytes_.data(), (std::streamsize)bytes_.size()); }
    const std::vector<char> &bytes() const { return bytes_; }

private:
    std::vector<char> bytes_;
};
class PublicKey : public KeyBlob {};
class SecretKey : public KeyBlob {};

// Evaluation keys of KeyGenerator::generate_evaluation_keys(decomposition_bit_count, count)
// (tests/parameters.cpp:89-90).  Stream format (this library's own; SEAL's NTT-form key layout is
// not reproducible without its sources): "B200EVK1", u32 dbc, u32 count, u32 k, u32 n, then the keys
// in COEFFICIENT form [count][k][digits][2][k][n] u64 (include/fhe_b200.h fhe_evk_create).
class EvaluationKeys {
public:
    EvaluationKeys() = default;
    void load(std::istream &stream)
    {
        char magic[8];
        std::uint32_t h[4];
        stream.read(magic, 8);
        stream.read((char *)h, sizeof(h));
        if (!stream || std::string(magic, 8) != "B200EVK1") throw std::invalid_argument("not an evaluation key stream of this library");
        dbc_ = h[0]; count_ = h[1]; k_ = h[2]; n_ = h[3];
        const std::uint32_t digits = max_digits();
        words_.resize((size_t)count_ * k_ * digits * 2 * k_ * n_);
        stream.read((char *)words_.data(), (std::streamsize)(words_.size() * sizeof(std::uint64_t)));
        if (!stream) throw std::invalid_argument("evaluation key stream: truncated");
        dev_.reset();
    }
    void save(std::ostream &stream) const
    {
        const std::uint32_t h[4] = {dbc_, count_, k_, n_};
        stream.write("B200EVK1", 8);
        stream.write((const char *)h, sizeof(h));
        stream.write((const char *)words_.data(), (std::streamsize)(words_.size() * sizeof(std::uint64_t)));
    }
    // from raw coefficient-form keys (what a key holder hands over)
    void assign(const SEALContext &context, const std::uint64_t *keys, int decomposition_bit_count, int count)
    {
        const seal_b200::Context &c = context.device_context();
        dbc_ = (std::uint32_t)decomposition_bit_count; count_ = (std::uint32_t)count;
        k_ = c.coeff_mod_count(); n_ = c.poly_degree();
        const std::uint32_t digits = fhe_evk_digits(c.handle(), dbc_);
        words_.assign(keys, keys + (size_t)count_ * k_ * digits * 2 * k_ * n_);
        dev_.reset();
    }
    int decomposition_bit_count() const { return (int)dbc_; }
    int size() const { return (int)count_; }
    const seal_b200::EvaluationKeys &device(const seal_b200::Context &c) const
    {
        if (!dev_) {
            if (words_.empty() || c.coeff_mod_count() != k_ || c.poly_degree() != n_)
                throw std::invalid_argument("evaluation keys are not valid for encryption parameters");
            dev_ = std::make_shared<seal_b200::EvaluationKeys>(c, words_.data(), dbc_, count_);
        }
        return *dev_;
    }

private:
    std::uint32_t max_digits() const { return dbc_ ? (60 + dbc_ - 1) / dbc_ : 0; }
    std::uint32_t dbc_ = 0, count_ = 0, k_ = 0, n_ = 0;
    std::vector<std::uint64_t> words_;
    mutable std::shared_ptr<seal_b200::EvaluationKeys> dev_;
};

// ---- operations -----------------------------------------------------------------------------
class Encryptor {
public:
    Encryptor(const SEALContext &context, const PublicKey &) : enc_(context.device_context()) {}
    void encrypt(const Plaintext &plain, Ciphertext &destination) const { enc_.encrypt(plain, destination); }
    operator seal_b200::Encryptor &() { return enc_; }

private:
    seal_b200::Encryptor enc_;
};

class Decryptor {
public:
    Decryptor(const SEALContext &, const SecretKey &) {}
    void decrypt(const Ciphertext &, Plaintext &) const { refuse(); }
    int invariant_noise_budget(const Ciphertext &) const
    {
        refuse();
        return 0;
    }

private:
    [[noreturn]] static void refuse()
    {
        throw std::logic_error("libfhe_b200 is the server side: secret-key operations (decrypt, noise budget) are not provided");
    }
};

class Evaluator : public seal_b200::Evaluator {
public:
    using Base = seal_b200::Evaluator;
    explicit Evaluator(const SEALContext &context) : Base(context.device_context()) {}
    using Base::add;
    using Base::add_plain;
    using Base::multiply;
    using Base::multiply_plain;
    using Base::negate;
    using Base::relinearize;
    using Base::square;
    using Base::sub;
    using Base::sub_plain;
    void relinearize(Ciphertext &a, const EvaluationKeys &evk) { Base::relinearize(a, evk.device(context())); }
    // SEAL's three-argument forms
    void add(const Ciphertext &a, const Ciphertext &b, Ciphertext &dest) { dest = a; Base::add(dest, b); }
    void sub(const Ciphertext &a, const Ciphertext &b, Ciphertext &dest) { dest = a; Base::sub(dest, b); }
    void negate(const Ciphertext &a, Ciphertext &dest) { dest = a; Base::negate(dest); }
    void multiply(const Ciphertext &a, const Ciphertext &b, Ciphertext &dest) { dest = a; Base::multiply(dest, b); }
    void square(const Ciphertext &a, Ciphertext &dest) { dest = a; Base::square(dest); }
    void multiply_plain(const Ciphertext &a, const Plaintext &p, Ciphertext &dest) { dest = a; Base::multiply_plain(dest, p); }
    void add_plain(const Ciphertext &a, const Plaintext &p, Ciphertext &dest) { dest = a; Base::add_plain(dest, p); }
    void sub_plain(const Ciphertext &a, const Plaintext &p, Ciphertext &dest) { dest = a; Base::sub_plain(dest, p); }
    void relinearize(const Ciphertext &a, const EvaluationKeys &evk, Ciphertext &dest) { dest = a; relinearize(dest, evk); }
};

}  // namespace seal
