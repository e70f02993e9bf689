// seal_wire_b200.hpp -- seal::Ciphertext::save / load wire format (SEAL 2.3), SURVEY.md 8f-N2.
//
// The reference's client and server exchange ciphertexts as files of concatenated
// seal::Ciphertext::save records with no framing (client_jpeg.cpp:140-165 writes, server_jpeg.cpp:117
// reads, :146-153 writes back).  SEAL is absent from /root/reference, so the layout below is the
// published 2.3 layout as recalled in SURVEY.md Appendix A.8 -- UNVERIFIED against real SEAL files:
//
//   hash block     4 x uint64   (SHA-3 of the encryption parameters in SEAL; opaque here)
//   size_capacity  int32
//   size           int32        number of polynomials
//   poly_coeff_count int32      n + 1  (SEAL 2.3 keeps a zero top coefficient per limb)
//   coeff_mod_count  int32      k
//   data           size * k * (n + 1) uint64, [poly][limb][coeff], little endian
//
// The device layout drops the padding coefficient ([poly][limb][n]); these helpers convert.  The
// hash block cannot be recomputed without SEAL's hash; save() writes the block given by the caller
// (by default the one seen on the last load, so a server echoes its client's block) and load() can
// check it against an expected value.
#pragma once
#include <array>
#include <cstdint>
#include <cstring>
#include <istream>
#include <ostream>
#include <stdexcept>
#include <vector>

#include "seal_b200.hpp"

namespace seal_b200 {

using hash_block_type = std::array<uint64_t, 4>;

// Header variant.  The layout above is recalled, not checked against a file written by SEAL (absent here); the
// field most likely to differ between 2.3.x releases is size_capacity.  wire_has_capacity() = false selects
// records WITHOUT it (hash block, then size, poly_coeff_count, coeff_mod_count: 44 header bytes, data of
// exactly `size` polynomials); the servers expose it as `--wire nocapacity`.  Process-wide, set before any I/O.
inline bool &wire_has_capacity()
{
    static bool v = true;
    return v;
}
inline size_t wire_header_bytes() { return 4 * sizeof(uint64_t) + (wire_has_capacity() ? 4 : 3) * sizeof(int32_t); }

struct WireHeader {
    hash_block_type hash{};
    int32_t size_capacity = 0, size = 0, poly_coeff_count = 0, coeff_mod_count = 0;
};

// Reads one record; `host` receives size * k * n words in device layout.  Returns false at a clean
// end of stream (no bytes left), throws std::invalid_argument on a truncated or mismatching record.
inline bool load_ciphertext(std::istream &in, const Context &ctx, std::vector<uint64_t> &host, WireHeader &hdr)
{
    in.read((char *)hdr.hash.data(), sizeof(uint64_t) * 4);
    if (in.gcount() == 0 && in.eof()) return false;
    int32_t meta[4];
    const bool cap = wire_has_capacity();
    in.read((char *)(meta + (cap ? 0 : 1)), (cap ? 4 : 3) * sizeof(int32_t));
    if (!in) throw std::invalid_argument("ciphertext stream: truncated header");
    if (!cap) meta[0] = meta[1];
    hdr.size_capacity = meta[0];
    hdr.size = meta[1];
    hdr.poly_coeff_count = meta[2];
    hdr.coeff_mod_count = meta[3];
    const uint32_t n = ctx.poly_degree(), k = ctx.coeff_mod_count();
    if (hdr.size < 2 || hdr.size > hdr.size_capacity || hdr.poly_coeff_count != (int32_t)n + 1 || hdr.coeff_mod_count != (int32_t)k)
        throw std::invalid_argument("ciphertext is not valid for encryption parameters");
    host.resize((size_t)hdr.size * k * n);
    std::vector<uint64_t> row(n + 1);
    for (size_t p = 0; p < (size_t)hdr.size * k; ++p) {
        in.read((char *)row.data(), (std::streamsize)(row.size() * sizeof(uint64_t)));
        if (!in) throw std::invalid_argument("ciphertext stream: truncated data");
        if (row[n] != 0) throw std::invalid_argument("ciphertext stream: non-zero padding coefficient");
        for (uint32_t i = 0; i < n; ++i)
            if (row[i] >= ctx.coeff_modulus()[p % k]) throw std::invalid_argument("ciphertext coefficient is not reduced");
        std::copy(row.begin(), row.begin() + n, host.begin() + p * n);
    }
    // SEAL also stores the (unused) capacity tail
    const size_t tail = (size_t)(hdr.size_capacity - hdr.size) * k * (n + 1) * sizeof(uint64_t);
    if (tail) in.ignore((std::streamsize)tail);
    return true;
}

inline void save_ciphertext(std::ostream &out, const Context &ctx, const uint64_t *host, uint32_t size, const hash_block_type &hash)
{
    const uint32_t n = ctx.poly_degree(), k = ctx.coeff_mod_count();
    out.write((const char *)hash.data(), sizeof(uint64_t) * 4);
    const int32_t meta[4] = {(int32_t)size, (int32_t)size, (int32_t)n + 1, (int32_t)k};
    const bool cap = wire_has_capacity();
    out.write((const char *)(meta + (cap ? 0 : 1)), (cap ? 4 : 3) * sizeof(int32_t));
    const uint64_t zero = 0;
    for (size_t p = 0; p < (size_t)size * k; ++p) {
        out.write((const char *)(host + p * n), (std::streamsize)(n * sizeof(uint64_t)));
        out.write((const char *)&zero, sizeof(zero));
    }
    if (!out) throw std::runtime_error("ciphertext stream: write failed");
}

// The same conversions on records held in memory (server/stream_b200.hpp reads and writes fixed-size records of a
// stream with several threads: pread / pwrite at computed offsets).
// bytes of a record that stores `capacity` polynomials
inline size_t record_bytes(const Context &ctx, uint32_t capacity)
{
    return wire_header_bytes() + (size_t)capacity * ctx.coeff_mod_count() * (ctx.poly_degree() + 1) * sizeof(uint64_t);
}
// header of a record (its first wire_header_bytes() bytes); throws like load_ciphertext on a mismatch with the context
inline WireHeader parse_record_header(const unsigned char *rec, const Context &ctx)
{
    WireHeader hdr;
    std::memcpy(hdr.hash.data(), rec, 32);
    int32_t meta[4];
    const bool cap = wire_has_capacity();
    std::memcpy(meta + (cap ? 0 : 1), rec + 32, (cap ? 4 : 3) * sizeof(int32_t));
    if (!cap) meta[0] = meta[1];
    hdr.size_capacity = meta[0];
    hdr.size = meta[1];
    hdr.poly_coeff_count = meta[2];
    hdr.coeff_mod_count = meta[3];
    if (hdr.size < 2 || hdr.size > hdr.size_capacity || hdr.poly_coeff_count != (int32_t)ctx.poly_degree() + 1 ||
        hdr.coeff_mod_count != (int32_t)ctx.coeff_mod_count())
        throw std::invalid_argument("ciphertext is not valid for encryption parameters");
    return hdr;
}
// rec: a whole record of record_bytes(ctx, capacity) bytes; dst receives size * k * n words in device layout
inline WireHeader parse_record(const unsigned char *rec, const Context &ctx, uint64_t *dst)
{
    const WireHeader hdr = parse_record_header(rec, ctx);
    const uint32_t n = ctx.poly_degree(), k = ctx.coeff_mod_count();
    const unsigned char *p = rec + wire_header_bytes();
    for (size_t l = 0; l < (size_t)hdr.size * k; ++l, p += (size_t)(n + 1) * sizeof(uint64_t)) {
        uint64_t *d = dst + l * n;
        std::memcpy(d, p, (size_t)n * sizeof(uint64_t));
        uint64_t pad;
        std::memcpy(&pad, p + (size_t)n * sizeof(uint64_t), sizeof(pad));
        const uint64_t q = ctx.coeff_modulus()[l % k];
        uint64_t bad = pad;
        for (uint32_t i = 0; i < n; ++i) bad |= (uint64_t)(d[i] >= q);
        if (bad) throw std::invalid_argument(pad ? "ciphertext stream: non-zero padding coefficient" : "ciphertext coefficient is not reduced");
    }
    return hdr;
}
// rec: record_bytes(ctx, size) bytes
inline void format_record(unsigned char *rec, const Context &ctx, const uint64_t *host, uint32_t size, const hash_block_type &hash)
{
    const uint32_t n = ctx.poly_degree(), k = ctx.coeff_mod_count();
    std::memcpy(rec, hash.data(), 32);
    const int32_t meta[4] = {(int32_t)size, (int32_t)size, (int32_t)n + 1, (int32_t)k};
    const bool cap = wire_has_capacity();
    std::memcpy(rec + 32, meta + (cap ? 0 : 1), (cap ? 4 : 3) * sizeof(int32_t));
    unsigned char *p = rec + wire_header_bytes();
    const uint64_t zero = 0;
    for (size_t l = 0; l < (size_t)size * k; ++l, p += (size_t)(n + 1) * sizeof(uint64_t)) {
        std::memcpy(p, host + l * n, (size_t)n * sizeof(uint64_t));
        std::memcpy(p + (size_t)n * sizeof(uint64_t), &zero, sizeof(zero));
    }
}

// Ciphertext-level convenience (one H2D / D2H per record)
inline bool load_ciphertext(std::istream &in, const Context &ctx, Ciphertext &dst, WireHeader &hdr)
{
    std::vector<uint64_t> host;
    if (!load_ciphertext(in, ctx, host, hdr)) return false;
    Ciphertext c(ctx, (uint32_t)hdr.size);
    c.upload(host.data());
    dst = std::move(c);
    return true;
}
inline void save_ciphertext(std::ostream &out, const Ciphertext &c, const hash_block_type &hash)
{
    std::vector<uint64_t> host(c.context()->ct_words(c.size()));
    c.download(host.data());
    save_ciphertext(out, *c.context(), host.data(), c.size(), hash);
}

}  // namespace seal_b200
