// fhe_resize_b200.hpp -- the resize image-op library of the reference (homo/fhe_resize.h) on B200.
//
// Same call surface as the reference, evaluated through seal_b200::Evaluator on the device:
//
//   Cubic  (result, A, B, C, D, t, evaluator, encoder, encryptor)      homo/fhe_resize.h:143-189
//   Linear (result, A, B, t, evaluator, encoder, encryptor)            homo/fhe_resize.h:191-204
//   SImageData, GetPixelClamped                                       homo/fhe_resize.h:207-220
//   SampleLinear / SampleBicubic (ret, image, x, y, ...)              homo/fhe_resize.h:222-305
//   ResizeImage (source, ow, oh, destImage, dw, dh, inter, ...)       homo/fhe_resize.h:308-392
//
// The graphs are the reference's own, op for op, INCLUDING its quirks:
// Cubic's `t3` is t*t, not t^3 (:175); u = x / (dw - 1) * ow - 0.5 and int() truncation toward
// zero for the integer part with floor() for the fraction (:228-234, :351, :382); the fractions are
// encrypted on the server (:230-266) -- through whatever seal_b200::Encryptor the caller passes.
// ct x ct products grow the ciphertexts (no relinearisation in the reference): size 4 after the
// first Cubic level, size 6 after the second; sizes 3 / 4 for Linear.
// Like the reference, Cubic and Linear print their wall time in ms followed by ',' to stdout.
//
// The only departure is the ciphertext SOURCE of ResizeImage: the reference reads
// seal::Ciphertext::load from a file stream (:313-342); here a callable yields the next ciphertext
// of the RGB-interleaved row-major stream (client_resize.cpp:141-154), so the caller chooses the
// wire format (a raw u64 reader is provided; SEAL's own serialisation is SURVEY.md 8f-N2).
#pragma once
#include <algorithm>
#include <chrono>
#include <cmath>
#include <fstream>
#include <iostream>
#include <string>

#include "seal_b200.hpp"

namespace seal_b200 {

constexpr int BILINEAR = 0;  // fhe_resize.h:14
constexpr int BICUBIC = 1;   // fhe_resize.h:15

namespace detail {
// The bodies of Cubic and Linear, written once for any ciphertext type the Evaluator overloads
// (one seal-style Ciphertext, or a CiphertextBatch = the same graph on a whole row of pixels).
// fhe_resize.h:174-175: t2 = square(t); "t3" = multiply(t, t) (sic: t*t, not t^3)
template <class CT>
inline void cubic_powers(const CT &t, CT &t2, CT &t3, Evaluator &evaluator)
{
    t2 = t;
    evaluator.square(t2);
    t3 = t;
    evaluator.multiply(t3, t);
}
template <class CT>
inline void cubic_body(CT &result, CT &A, CT &B, CT &C, CT &D, CT &t, Evaluator &evaluator, FractionalEncoder &encoder,
                       const std::pair<CT, CT> *powers = nullptr)
{
    CT a, b, c, d;
    CT boaz1(A);
    CT boaz2(B); evaluator.multiply_plain(boaz2, encoder.encode(3));
    evaluator.sub(boaz2, boaz1);
    CT boaz3(C); evaluator.multiply_plain(boaz3, encoder.encode(3));
    evaluator.sub(boaz2, boaz3);
    CT boaz4(D);
    evaluator.add(boaz2, boaz4);
    a = boaz2;

    CT boaz5(A); evaluator.multiply_plain(boaz5, encoder.encode(2));
    CT boaz6(B); evaluator.multiply_plain(boaz6, encoder.encode(5));
    evaluator.sub(boaz5, boaz6);
    CT boaz7(C); evaluator.multiply_plain(boaz7, encoder.encode(4));
    evaluator.add(boaz5, boaz7);
    CT boaz8(D);
    evaluator.sub(boaz5, boaz8);
    b = boaz5;

    CT boaz9(A);
    CT boaz10(C);
    evaluator.sub(boaz10, boaz9);
    c = boaz10;

    d = B;

    CT t2, t3;
    if (powers) {  // square(t) and t*t are functions of t alone: a caller may compute them once per t
        t2 = powers->first;
        t3 = powers->second;
    } else {
        cubic_powers(t, t2, t3, evaluator);
    }

    evaluator.multiply(a, t3);
    evaluator.multiply(b, t2);
    evaluator.multiply(c, t);

    evaluator.add(a, b);
    evaluator.add(a, c);
    evaluator.multiply_plain(a, encoder.encode(0.5));
    evaluator.add(a, d);
    result = a;
}
template <class CT>
inline void linear_body(CT &result, CT &A, CT &B, CT &t, Evaluator &evaluator, FractionalEncoder &encoder)
{
    CT boaz1(t); evaluator.negate(boaz1); evaluator.add_plain(boaz1, encoder.encode(1.0));
    evaluator.multiply(boaz1, A);
    CT boaz2(B); evaluator.multiply(boaz2, t);
    evaluator.add(boaz1, boaz2);
    result = boaz1;
}
struct CsvMs {  // the reference prints the wall time of each call in ms followed by ','
    const Context &ctx;
    std::chrono::steady_clock::time_point start = std::chrono::steady_clock::now();
    explicit CsvMs(const Context &c) : ctx(c) {}
    ~CsvMs()
    {
        fhe_ctx_sync(ctx.handle());
        auto diff = std::chrono::steady_clock::now() - start;
        std::cout << std::chrono::duration<double, std::milli>(diff).count() << ',';
    }
};
}  // namespace detail

inline void Cubic(Ciphertext &result, Ciphertext &A, Ciphertext &B, Ciphertext &C, Ciphertext &D, Ciphertext &t,
                  Evaluator &evaluator, FractionalEncoder &encoder, Encryptor &encryptor)
{
    (void)encryptor;
    detail::CsvMs timer(evaluator.context());
    detail::cubic_body(result, A, B, C, D, t, evaluator, encoder);
}
inline void Linear(Ciphertext &result, Ciphertext &A, Ciphertext &B, Ciphertext &t, Evaluator &evaluator,
                   FractionalEncoder &encoder, Encryptor &encryptor)
{
    (void)encryptor;
    detail::CsvMs timer(evaluator.context());
    detail::linear_body(result, A, B, t, evaluator, encoder);
}
// the same graphs on `count` independent inputs at once (no timing output: not a reference call)
inline void Cubic(CiphertextBatch &result, CiphertextBatch &A, CiphertextBatch &B, CiphertextBatch &C, CiphertextBatch &D,
                  CiphertextBatch &t, Evaluator &evaluator, FractionalEncoder &encoder,
                  const std::pair<CiphertextBatch, CiphertextBatch> *powers = nullptr)
{
    detail::cubic_body(result, A, B, C, D, t, evaluator, encoder, powers);
}
inline void Linear(CiphertextBatch &result, CiphertextBatch &A, CiphertextBatch &B, CiphertextBatch &t, Evaluator &evaluator,
                   FractionalEncoder &encoder)
{
    detail::linear_body(result, A, B, t, evaluator, encoder);
}

struct SImageData {
    int width;
    int height;
    int start;
    std::vector<std::vector<Ciphertext>> pixels;
};

inline std::vector<Ciphertext> GetPixelClamped(const SImageData &image, int x, int y)
{
    x = std::min(std::max(x, 0), image.width - 1);
    y = std::min(std::max(y, 0), image.height - 1);
    return image.pixels.at((size_t)(y * image.width + x - image.start));
}

inline void SampleLinear(std::vector<Ciphertext> &ret, const SImageData &image, float x, float y, Evaluator &evaluator,
                         FractionalEncoder &encoder, Encryptor &encryptor)
{
    int xint = int(x);
    Ciphertext xfract;
    encryptor.encrypt(encoder.encode(x - std::floor(x)), xfract);
    int yint = int(y);
    Ciphertext yfract;
    encryptor.encrypt(encoder.encode(y - std::floor(y)), yfract);

    auto p00 = GetPixelClamped(image, xint + 0, yint + 0);
    auto p10 = GetPixelClamped(image, xint + 1, yint + 0);
    auto p01 = GetPixelClamped(image, xint + 0, yint + 1);
    auto p11 = GetPixelClamped(image, xint + 1, yint + 1);

    Ciphertext col0, col1;
    for (int i = 0; i < 3; ++i) {
        Linear(col0, p00[i], p10[i], xfract, evaluator, encoder, encryptor);
        Linear(col1, p01[i], p11[i], xfract, evaluator, encoder, encryptor);
        Linear(ret[i], col0, col1, yfract, evaluator, encoder, encryptor);
    }
}

inline void SampleBicubic(std::vector<Ciphertext> &ret, const SImageData &image, float x, float y, Evaluator &evaluator,
                          FractionalEncoder &encoder, Encryptor &encryptor)
{
    int xint = int(x);
    Ciphertext xfract;
    encryptor.encrypt(encoder.encode(x - std::floor(x)), xfract);
    int yint = int(y);
    Ciphertext yfract;
    encryptor.encrypt(encoder.encode(y - std::floor(y)), yfract);

    std::vector<Ciphertext> p[4][4];  // p[column offset][row offset]
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) p[c][r] = GetPixelClamped(image, xint - 1 + c, yint - 1 + r);

    Ciphertext col0, col1, col2, col3;
    for (int i = 0; i < 3; ++i) {
        Cubic(col0, p[0][0][i], p[1][0][i], p[2][0][i], p[3][0][i], xfract, evaluator, encoder, encryptor);
        Cubic(col1, p[0][1][i], p[1][1][i], p[2][1][i], p[3][1][i], xfract, evaluator, encoder, encryptor);
        Cubic(col2, p[0][2][i], p[1][2][i], p[2][2][i], p[3][2][i], xfract, evaluator, encoder, encryptor);
        Cubic(col3, p[0][3][i], p[1][3][i], p[2][3][i], p[3][3][i], xfract, evaluator, encoder, encryptor);
        Cubic(ret[i], col0, col1, col2, col3, yfract, evaluator, encoder, encryptor);
    }
}

// next ciphertext of the input stream (replaces `c.load(infile)`, fhe_resize.h:333-338)
using CiphertextSource = std::function<void(Ciphertext &)>;

// raw reader: each ciphertext is 2*k*n little-endian u64 in [poly][limb][n] order, no framing
inline CiphertextSource raw_file_source(const Context &ctx, const std::string &path)
{
    auto in = std::make_shared<std::ifstream>(path, std::ios::binary);
    if (!*in) throw std::invalid_argument("cannot open " + path);
    auto buf = std::make_shared<std::vector<uint64_t>>(ctx.ct_words(2));
    const Context *c = &ctx;
    return [in, buf, c](Ciphertext &dst) {
        in->read((char *)buf->data(), (std::streamsize)(buf->size() * sizeof(uint64_t)));
        if (!*in) throw std::invalid_argument("ciphertext stream ended early");
        Ciphertext x(*c, 2);
        x.upload(buf->data());
        dst = std::move(x);
    };
}

inline void ResizeImage(const CiphertextSource &load, int original_width, int original_height, SImageData &destImage,
                        int dest_width, int dest_height, int inter, Evaluator &evaluator, FractionalEncoder &encoder,
                        Encryptor &encryptor)
{
    SImageData srcImage;
    srcImage.width = original_width;
    srcImage.height = original_height;
    srcImage.start = 0;

    int init_rows = 0;
    Ciphertext c;
    if (inter == BILINEAR) init_rows = 2;
    else if (inter == BICUBIC) init_rows = 4;
    else throw std::invalid_argument("unknown interpolation");
    auto load_pixel = [&]() {
        std::vector<Ciphertext> pixel;
        for (int j = 0; j < 3; ++j) {
            load(c);
            pixel.push_back(c);
        }
        return pixel;
    };
    for (int i = 0; i < original_width * init_rows; i++) srcImage.pixels.push_back(load_pixel());

    destImage.width = dest_width;
    destImage.height = dest_height;
    destImage.start = 0;
    destImage.pixels.clear();

    for (int y = 0; y < destImage.height; ++y) {
        float v = float(y) / float(destImage.height - 1) * float(srcImage.height) - 0.5;
        int new_start = std::min(int(v) - init_rows / 2 + 1, srcImage.height - init_rows) * srcImage.width;
        if (new_start > srcImage.start) {
            for (int i = srcImage.start + init_rows * srcImage.width; i < new_start; i++)
                for (int j = 0; j < 3; j++) load(c);  // rows skipped entirely
            std::vector<std::vector<Ciphertext>> new_pixels;
            int present = 0;
            for (int i = new_start - srcImage.start; i < init_rows * srcImage.width; i++) {
                new_pixels.push_back(std::move(srcImage.pixels[(size_t)i]));
                present++;
            }
            for (int i = present; i < init_rows * srcImage.width; i++) new_pixels.push_back(load_pixel());
            srcImage.start = new_start;
            srcImage.pixels = std::move(new_pixels);
        }
        for (int x = 0; x < destImage.width; ++x) {
            float u = float(x) / float(destImage.width - 1) * float(srcImage.width) - 0.5;
            std::vector<Ciphertext> sample(3);
            if (inter == BILINEAR) SampleLinear(sample, srcImage, u, v, evaluator, encoder, encryptor);
            else SampleBicubic(sample, srcImage, u, v, evaluator, encoder, encryptor);
            destImage.pixels.push_back(std::move(sample));
        }
    }
}

// the reference's signature: reads the raw stream from a file
inline void ResizeImage(const std::string &infile_str, int original_width, int original_height, SImageData &destImage,
                        int dest_width, int dest_height, int inter, Evaluator &evaluator, FractionalEncoder &encoder,
                        Encryptor &encryptor)
{
    ResizeImage(raw_file_source(evaluator.context(), infile_str), original_width, original_height, destImage, dest_width,
                dest_height, inter, evaluator, encoder, encryptor);
}

// ---------------------------------------------------------------------------------------------
// B200-first variant of ResizeImage + SampleBicubic / SampleLinear: identical arithmetic and
// identical results (tests/cpp/test_resize_decode.cpp compares them ciphertext by ciphertext), but
// one output ROW is sampled at a time: the 16 (4) neighbour pixels of every output pixel and
// channel are gathered into CiphertextBatches of dest_width * 3 elements and each Evaluator call of
// the Cubic / Linear graph covers the whole row.  The two fractions are encrypted once per output
// pixel, in the reference's call order (x then y), and shared by the three channels.
inline void ResizeImageBatched(const CiphertextSource &load, int original_width, int original_height, SImageData &destImage,
                               int dest_width, int dest_height, int inter, Evaluator &evaluator, FractionalEncoder &encoder,
                               Encryptor &encryptor)
{
    const Context &ctx = evaluator.context();
    SImageData srcImage;
    srcImage.width = original_width;
    srcImage.height = original_height;
    srcImage.start = 0;
    const int init_rows = inter == BILINEAR ? 2 : 4, taps = init_rows;
    if (inter != BILINEAR && inter != BICUBIC) throw std::invalid_argument("unknown interpolation");
    Ciphertext c;
    auto load_pixel = [&]() {
        std::vector<Ciphertext> pixel;
        for (int j = 0; j < 3; ++j) {
            load(c);
            pixel.push_back(c);
        }
        return pixel;
    };
    for (int i = 0; i < original_width * init_rows; i++) srcImage.pixels.push_back(load_pixel());
    destImage.width = dest_width;
    destImage.height = dest_height;
    destImage.start = 0;
    destImage.pixels.clear();
    const size_t count = (size_t)dest_width * 3;
    for (int y = 0; y < dest_height; ++y) {
        float v = float(y) / float(dest_height - 1) * float(srcImage.height) - 0.5;
        int new_start = std::min(int(v) - init_rows / 2 + 1, srcImage.height - init_rows) * srcImage.width;
        if (new_start > srcImage.start) {
            for (int i = srcImage.start + init_rows * srcImage.width; i < new_start; i++)
                for (int j = 0; j < 3; j++) load(c);
            std::vector<std::vector<Ciphertext>> new_pixels;
            int present = 0;
            for (int i = new_start - srcImage.start; i < init_rows * srcImage.width; i++) {
                new_pixels.push_back(std::move(srcImage.pixels[(size_t)i]));
                present++;
            }
            for (int i = present; i < init_rows * srcImage.width; i++) new_pixels.push_back(load_pixel());
            srcImage.start = new_start;
            srcImage.pixels = std::move(new_pixels);
        }
        // gather: tap[r][cidx] holds, for every (x, channel), the neighbour (xint + off + cidx, yint + off + r)
        std::vector<std::vector<CiphertextBatch>> tap((size_t)taps, std::vector<CiphertextBatch>());
        for (auto &row : tap)
            for (int cidx = 0; cidx < taps; ++cidx) row.emplace_back(ctx, count, 2);
        // one encrypted fraction pair per output pixel; the three channels share it
        CiphertextBatch xf1(ctx, (size_t)dest_width, 2), yf1(ctx, (size_t)dest_width, 2);
        const int off = inter == BICUBIC ? -1 : 0;
        const int yint = int(v);
        for (int x = 0; x < dest_width; ++x) {
            float u = float(x) / float(dest_width - 1) * float(srcImage.width) - 0.5;
            const int xint = int(u);
            Ciphertext xf, yf;
            encryptor.encrypt(encoder.encode(u - std::floor(u)), xf);
            encryptor.encrypt(encoder.encode(v - std::floor(v)), yf);
            xf1.set((size_t)x, xf);
            yf1.set((size_t)x, yf);
            for (int r = 0; r < taps; ++r)
                for (int cidx = 0; cidx < taps; ++cidx) {
                    int px = std::min(std::max(xint + off + cidx, 0), srcImage.width - 1);
                    int py = std::min(std::max(yint + off + r, 0), srcImage.height - 1);
                    const std::vector<Ciphertext> &pixel = srcImage.pixels.at((size_t)(py * srcImage.width + px - srcImage.start));
                    for (int ch = 0; ch < 3; ++ch) tap[(size_t)r][(size_t)cidx].set((size_t)x * 3 + ch, pixel[(size_t)ch]);
                }
        }
        auto per_channel = [&](const CiphertextBatch &one) {  // element x -> elements 3x, 3x+1, 3x+2
            CiphertextBatch r(ctx, count, one.size());
            const size_t w = ctx.ct_words(one.size()) * sizeof(uint64_t);
            for (int ch = 0; ch < 3; ++ch)
                check(fhe_device_copy_2d(ctx.handle(), r.device_data() + (size_t)ch * ctx.ct_words(one.size()), 3 * w,
                                         one.device_data(), w, w, one.count()));
            return r;
        };
        CiphertextBatch xfract = per_channel(xf1), yfract = per_channel(yf1);
        CiphertextBatch out;
        if (inter == BICUBIC) {
            // the powers of the fractions are the same ciphertexts in every Cubic that uses them (4 rows x 3
            // channels for x, 3 channels for y): computed once per output pixel, bit-identical
            CiphertextBatch x2, x3, y2, y3;
            detail::cubic_powers(xf1, x2, x3, evaluator);
            detail::cubic_powers(yf1, y2, y3, evaluator);
            const std::pair<CiphertextBatch, CiphertextBatch> xp(per_channel(x2), per_channel(x3)), yp(per_channel(y2), per_channel(y3));
            std::vector<CiphertextBatch> col(4);
            for (int r = 0; r < 4; ++r)
                Cubic(col[(size_t)r], tap[(size_t)r][0], tap[(size_t)r][1], tap[(size_t)r][2], tap[(size_t)r][3], xfract, evaluator, encoder, &xp);
            Cubic(out, col[0], col[1], col[2], col[3], yfract, evaluator, encoder, &yp);
        } else {
            CiphertextBatch col0, col1;
            Linear(col0, tap[0][0], tap[0][1], xfract, evaluator, encoder);
            Linear(col1, tap[1][0], tap[1][1], xfract, evaluator, encoder);
            Linear(out, col0, col1, yfract, evaluator, encoder);
        }
        for (int x = 0; x < dest_width; ++x) {
            std::vector<Ciphertext> sample;
            for (int ch = 0; ch < 3; ++ch) sample.push_back(out.get((size_t)x * 3 + ch));
            destImage.pixels.push_back(std::move(sample));
        }
        // Nothing above waits for the device, so the host could queue many rows ahead; the
        // temporaries of every queued row would then be live in the stream-ordered pool at once
        // (gigabytes at size 6) and the pool would keep growing through the driver.  One
        // synchronisation per output row bounds that to a row's worth.
        check(fhe_ctx_sync(ctx.handle()));
    }
}

}  // namespace seal_b200
