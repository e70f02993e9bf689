// jpeg_entropy_b200.hpp -- the client-side tail of the JPEG pipeline: what
// client_jpeg --recieve does with the decrypted, decoded, rounded DCT coefficients the server
// produced (homo/client_jpeg.cpp:221-291): quantisation tables for the file header, baseline JFIF
// headers, Huffman entropy coding of each 8x8 block (homo/fhe_image.h:331-397: writeBits, calcBits,
// processBlock) and the RMS comparator (fhe_image.h:508-521).  CPU only by the reference's own
// design (the secret key and the plaintext never leave the client); nothing here touches the GPU.
//
// Written from the JPEG standard (ITU-T T.81): the Huffman code tables are BUILT from the Annex K.3
// BITS / HUFFVAL specifications by the Annex C procedure instead of being tabulated, and the very
// same specifications are emitted in the DHT segment, so a file is decodable by construction.
// Reference quirks kept behind parameters: the reference writes header tables scaled by
// QUALITY = 0 (fhe_image.h:32), i.e. all ones, while the server divides by YQT / UVQT -- pass
// quality = 100 to get header tables equal to the divisors actually applied.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <vector>

namespace jpeg_b200 {

// natural (row-major) index -> zig-zag position (T.81 Figure 5; fhe_image.h:34 s_ZigZag)
inline const uint8_t *zigzag_position()
{
    static uint8_t zz[64];
    static bool init = false;
    if (!init) {
        int r = 0, c = 0;
        for (int pos = 0; pos < 64; ++pos) {
            zz[8 * r + c] = (uint8_t)pos;
            if ((r + c) % 2 == 0) {            // moving up-right
                if (c == 7) ++r; else if (r == 0) ++c; else { --r; ++c; }
            } else {                           // moving down-left
                if (r == 7) ++c; else if (c == 0) ++r; else { ++r; --c; }
            }
        }
        init = true;
    }
    return zz;
}

// header quantisation table in zig-zag order (client_jpeg.cpp:222-229): (q * quality + 50) / 100,
// clamped to [1, 255]
inline void header_quant_table(const int *natural64, int quality, uint8_t out_zigzag[64])
{
    const uint8_t *zz = zigzag_position();
    for (int i = 0; i < 64; ++i) {
        int v = (natural64[i] * quality + 50) / 100;
        out_zigzag[zz[i]] = (uint8_t)(v < 1 ? 1 : v > 255 ? 255 : v);
    }
}

// ---- Huffman tables: T.81 Annex K.3 specifications, Annex C code generation ----------------------
struct HuffSpec {
    uint8_t bits[16];            // number of codes of length 1..16
    std::vector<uint8_t> vals;   // symbols in order of increasing code length
};
struct HuffCodes {
    uint16_t code[256] = {}, len[256] = {};
};

inline const HuffSpec &spec_dc_luminance()
{
    static const HuffSpec s{{0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0}, {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11}};
    return s;
}
inline const HuffSpec &spec_dc_chrominance()
{
    static const HuffSpec s{{0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0}, {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11}};
    return s;
}
inline const HuffSpec &spec_ac_luminance()
{
    static const HuffSpec s{
        {0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d},
        {0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71,
         0x14, 0x32, 0x81, 0x91, 0xa1, 0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72,
         0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37,
         0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59,
         0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83,
         0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3,
         0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3,
         0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2,
         0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa}};
    return s;
}
inline const HuffSpec &spec_ac_chrominance()
{
    static const HuffSpec s{
        {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77},
        {0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22,
         0x32, 0x81, 0x08, 0x14, 0x42, 0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1,
         0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x35, 0x36,
         0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58,
         0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a,
         0x82, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a,
         0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba,
         0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda,
         0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa}};
    return s;
}

// Annex C: codes of each length are consecutive integers, the first code of length l+1 is twice
// (last code of length l, plus one)
inline HuffCodes build_codes(const HuffSpec &s)
{
    size_t total = 0;
    for (int l = 0; l < 16; ++l) total += s.bits[l];
    if (total != s.vals.size() || total > 256) throw std::logic_error("inconsistent Huffman specification");
    HuffCodes h;
    unsigned code = 0;
    size_t k = 0;
    for (int l = 1; l <= 16; ++l) {
        for (int i = 0; i < s.bits[l - 1]; ++i, ++k) {
            h.code[s.vals[k]] = (uint16_t)code++;
            h.len[s.vals[k]] = (uint16_t)l;
        }
        code <<= 1;
    }
    return h;
}

// ---- the writer ----------------------------------------------------------------------------------
class JpegWriter {
public:
    // quant tables in natural order (YQT / UVQT, fhe_image.h:99-100); quality as client_jpeg.cpp:224
    JpegWriter(std::FILE *fp, int width, int height, const int *yqt, const int *uvqt, int quality)
        : fp_(fp), ydc_(build_codes(spec_dc_luminance())), yac_(build_codes(spec_ac_luminance())),
          cdc_(build_codes(spec_dc_chrominance())), cac_(build_codes(spec_ac_chrominance()))
    {
        if (!fp || width <= 0 || height <= 0 || width > 65535 || height > 65535 || width % 8 || height % 8)
            throw std::invalid_argument("JpegWriter: bad file or dimensions (multiples of 8, as the reference requires)");
        uint8_t yt[64], ct[64];
        header_quant_table(yqt, quality, yt);
        header_quant_table(uvqt, quality, ct);
        // SOI, APP0 (JFIF 1.1, no density), DQT with both tables
        put({0xFF, 0xD8, 0xFF, 0xE0, 0, 0x10, 'J', 'F', 'I', 'F', 0, 1, 1, 0, 0, 1, 0, 1, 0, 0, 0xFF, 0xDB, 0, 0x84, 0});
        std::fwrite(yt, 1, 64, fp_);
        std::putc(1, fp_);
        std::fwrite(ct, 1, 64, fp_);
        // SOF0: 8 bit, three components, no subsampling; component 1 -> table 0, 2 and 3 -> table 1
        put({0xFF, 0xC0, 0, 0x11, 8, height >> 8, height & 0xFF, width >> 8, width & 0xFF, 3, 1, 0x11, 0, 2, 0x11, 1, 3, 0x11, 1});
        // DHT: four tables in one segment (length 2 + 4 * 17 + 12 + 162 + 12 + 162 = 0x01A2)
        put({0xFF, 0xC4, 0x01, 0xA2});
        dht(0x00, spec_dc_luminance());
        dht(0x10, spec_ac_luminance());
        dht(0x01, spec_dc_chrominance());
        dht(0x11, spec_ac_chrominance());
        // SOS: Y uses tables 0/0, Cb and Cr 1/1; spectral selection 0..63
        put({0xFF, 0xDA, 0, 0xC, 3, 1, 0, 2, 0x11, 3, 0x11, 0, 0x3F, 0});
    }

    // One MCU = one 8x8 block per component, coefficients already quantised and rounded, indexed by
    // ZIG-ZAG position (the reference fills block_zz[k][s_ZigZag[j]], client_jpeg.cpp:270).
    void write_mcu(const int y[64], const int cb[64], const int cr[64])
    {
        dc_[0] = block(y, dc_[0], ydc_, yac_);
        dc_[1] = block(cb, dc_[1], cdc_, cac_);
        dc_[2] = block(cr, dc_[2], cdc_, cac_);
    }

    // pad the last byte with ones and write EOI (client_jpeg.cpp:280-284)
    void finish()
    {
        bits(0x7F, 7);
        put({0xFF, 0xD9});
        std::fflush(fp_);
    }

private:
    std::FILE *fp_;
    HuffCodes ydc_, yac_, cdc_, cac_;
    int dc_[3] = {0, 0, 0};
    uint32_t buf_ = 0;
    int cnt_ = 0;

    void put(std::initializer_list<int> bytes) { for (int b : bytes) std::putc(b & 0xFF, fp_); }
    void dht(int id, const HuffSpec &s)
    {
        std::putc(id, fp_);
        std::fwrite(s.bits, 1, 16, fp_);
        std::fwrite(s.vals.data(), 1, s.vals.size(), fp_);
    }
    // append `len` bits; a 0xFF data byte is followed by a stuffed zero (T.81 B.1.1.5)
    void bits(unsigned code, int len)
    {
        cnt_ += len;
        buf_ |= code << (24 - cnt_);
        while (cnt_ >= 8) {
            const int c = (buf_ >> 16) & 255;
            std::putc(c, fp_);
            if (c == 255) std::putc(0, fp_);
            buf_ <<= 8;
            cnt_ -= 8;
        }
    }
    // magnitude category and the additional bits of a non-zero value (T.81 F.1.2.1: negative values
    // as the low bits of value - 1)
    static void category(int val, unsigned &extra, int &size)
    {
        int mag = val < 0 ? -val : val;
        size = 0;
        while (mag) { ++size; mag >>= 1; }
        extra = (unsigned)(val < 0 ? val - 1 : val) & ((1u << size) - 1);
    }
    int block(const int du[64], int prev_dc, const HuffCodes &dc, const HuffCodes &ac)
    {
        const int diff = du[0] - prev_dc;
        if (diff == 0) bits(dc.code[0], dc.len[0]);
        else {
            unsigned extra; int size;
            category(diff, extra, size);
            bits(dc.code[size], dc.len[size]);
            bits(extra, size);
        }
        int last = 63;
        while (last > 0 && du[last] == 0) --last;
        if (last == 0) { bits(ac.code[0x00], ac.len[0x00]); return du[0]; }   // EOB straight away
        for (int i = 1; i <= last; ++i) {
            int run = 0;
            while (du[i] == 0 && i <= last) { ++i; ++run; }
            for (; run >= 16; run -= 16) bits(ac.code[0xF0], ac.len[0xF0]);      // ZRL
            unsigned extra; int size;
            category(du[i], extra, size);
            if (size > 10) throw std::range_error("JPEG AC coefficient out of range");
            bits(ac.code[(run << 4) + size], ac.len[(run << 4) + size]);
            bits(extra, size);
        }
        if (last != 63) bits(ac.code[0x00], ac.len[0x00]);
        return du[0];
    }
};

// sqrt(mean squared difference) over all samples (compare_jpeg_jojpeg, fhe_image.h:514-519)
inline double rms_error(const uint8_t *a, const uint8_t *b, size_t count)
{
    double acc = 0;
    for (size_t i = 0; i < count; ++i) {
        const double d = (double)a[i] - (double)b[i];
        acc += d * d;
    }
    return count ? std::sqrt(acc / (double)count) : 0.0;
}

// decoded value -> integer coefficient exactly as client_jpeg.cpp:270 rounds it (single-precision
// round half away from zero)
inline int round_coefficient(double v) { return (int)(v < 0 ? std::ceil((float)v - 0.5f) : std::floor((float)v + 0.5f)); }

}  // namespace jpeg_b200
