// fhe_resize_b200.hpp -- the pixel loop of the reference's resize server (homo/fhe_resize.h:308-392:
// ResizeImage over SampleBicubic / SampleLinear over Cubic / Linear) as a B200-first row engine.
//
// The reference's own op graphs are NOT restated here: homo/fhe_resize.h compiles unchanged against
// include/seal/seal.h and runs op by op on the GPU (oracle/Makefile `ref`,
// tests/test_reference_sources_on_gpu.py).  This header is the fast path a server uses instead:
//
//   * one output ROW at a time: the 15 Cubic (9 Linear) calls of every output pixel become five
//     (three) calls of the fused device graphs fhe_cubic / fhe_linear (include/fhe_b200.h), each over
//     the whole row x 3 channels, reading the 4 x 4 (2 x 2) neighbourhoods through pointer tables;
//   * the powers of the fractions are formed once and kept as operand forms: square(t) and
//     multiply(t, t) are functions of t alone (and the same ciphertext: fhe_resize.h:174-175), the
//     three channels of a pixel share them, and with a deterministic Encryptor (the key-less
//     transparent one) the x fractions do not depend on the row and the y fraction not on the column;
//   * with a deterministic Encryptor the horizontal pass of a SOURCE row is the same ciphertext for
//     every output row that touches it, so it is computed once per source row (two output rows share
//     it at a 2x reduction);
//   * source rows stream in and output rows stream out (RowSource / RowSink): device memory holds a
//     window, not the image -- 512 x 512 x 3 ciphertexts of n = 8192 are 384 GiB.
//
// Every ciphertext produced equals the reference's bit for bit (tests/cpp/test_resize_decode.cpp
// against the oracle's op-by-op graph; with a callback Encryptor the encryptions are drawn in the
// reference's order, x then y per output pixel, and nothing is shared across pixels).
// The coordinates follow fhe_resize.h:351,382 (float arithmetic, int() truncation, floor() for the
// fraction, clamped neighbourhoods :215-220).
#pragma once
#include <algorithm>
#include <cmath>
#include <fstream>
#include <map>
#include <string>

#include "seal_b200.hpp"

namespace seal_b200 {

constexpr int BILINEAR = 0;  // fhe_resize.h:14
constexpr int BICUBIC = 1;   // fhe_resize.h:15

// the five plaintexts of Cubic (fhe_resize.h:151-183) from the caller's encoder
class ResizePlan {
public:
    ResizePlan(const Context &ctx, const FractionalEncoder &encoder)
    {
        std::vector<uint64_t> plains;
        for (double v : {3.0, 2.0, 5.0, 4.0, 0.5}) {
            const Plaintext p = encoder.encode(v);
            plains.insert(plains.end(), p.coeffs().begin(), p.coeffs().end());
        }
        check(fhe_resize_plan_create(ctx.handle(), plains.data(), &h_));
    }
    ~ResizePlan() { fhe_resize_plan_destroy(h_); }
    ResizePlan(const ResizePlan &) = delete;
    ResizePlan &operator=(const ResizePlan &) = delete;
    const fhe_resize_plan *handle() const { return h_; }

private:
    fhe_resize_plan *h_ = nullptr;
};

// next ciphertext of the input stream (what `c.load(infile)` yields, fhe_resize.h:333-338)
using CiphertextSource = std::function<void(Ciphertext &)>;
// Source row y of the RGB-interleaved row-major stream (client_resize.cpp:141-154): returns the row's
// original_width * 3 size-2 ciphertexts when `needed` (a fresh batch, or a slice of storage the caller
// keeps resident), skips the row of the stream otherwise.  Rows are requested in increasing order,
// each once.
using RowSource = std::function<CiphertextBatch(int y, bool needed)>;
// Output row y: dest_width * 3 ciphertexts (size 6 bicubic, 4 bilinear), RGB per pixel
using RowSink = std::function<void(int y, CiphertextBatch &row)>;

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
// a RowSource over a ciphertext-at-a-time source
inline RowSource rows_of(const CiphertextSource &load, int original_width)
{
    return [load, original_width](int, bool needed) {
        Ciphertext c;
        CiphertextBatch row;
        for (int i = 0; i < original_width * 3; ++i) {
            load(c);
            if (needed && i == 0) row = CiphertextBatch(*c.context(), (size_t)original_width * 3, 2);
            if (needed) row.set((size_t)i, c);
        }
        return row;
    };
}

class RowResizer {
public:
    RowResizer(Evaluator &evaluator, FractionalEncoder &encoder, Encryptor &encryptor, int original_width, int original_height,
               int dest_width, int dest_height, int inter)
        : ev_(evaluator), enc_(encoder), cr_(encryptor), ctx_(evaluator.context()), ow_(original_width), oh_(original_height),
          dw_(dest_width), dh_(dest_height), inter_(inter), taps_(inter == BICUBIC ? 4 : 2), off_(inter == BICUBIC ? -1 : 0)
    {
        if (inter != BILINEAR && inter != BICUBIC) throw std::invalid_argument("unknown interpolation");
        if (ow_ < 1 || oh_ < 1 || dw_ < 1 || dh_ < 1) throw std::invalid_argument("bad image size");
        if (evaluator.relinearization())
            throw std::invalid_argument("the fused resize graphs follow the reference, which never relinearises");
        if (inter == BICUBIC) plan_ = std::make_unique<ResizePlan>(ctx_, encoder);
    }

    // output rows [y_begin, y_end) (a band: ranks of a multi-GPU run take disjoint bands and feed each
    // the source rows it touches)
    void run(const RowSource &source, const RowSink &sink, int y_begin = 0, int y_end = -1)
    {
        if (y_end < 0) y_end = dh_;
        const size_t count = (size_t)dw_ * 3;
        const bool det = cr_.deterministic();
        std::vector<int> xint((size_t)dw_);
        std::vector<double> xfr((size_t)dw_);
        for (int x = 0; x < dw_; ++x) {
            const float u = float(x) / float(dw_ - 1) * float(ow_) - 0.5;
            xint[(size_t)x] = int(u);
            xfr[(size_t)x] = u - std::floor(u);
        }
        Fraction fx;  // the x fractions of a row: shared by all rows when the Encryptor is deterministic
        if (det) {
            CiphertextBatch xf(ctx_, (size_t)dw_, 2);
            Ciphertext c;
            for (int x = 0; x < dw_; ++x) {
                cr_.encrypt(enc_.encode(xfr[(size_t)x]), c);
                xf.set((size_t)x, c);
            }
            fx = fraction(xf);
        }
        int next_src = 0;
        auto taps_of = [&](int y, int rows[4], int &yint, double &yfr) {
            const float v = float(y) / float(dh_ - 1) * float(oh_) - 0.5;
            yint = int(v);
            yfr = v - std::floor(v);
            for (int r = 0; r < taps_; ++r) rows[r] = std::min(std::max(yint + off_ + r, 0), oh_ - 1);
        };
        // request from the source every row up to the last one output row y touches
        auto load_for = [&](int y) {
            int rows[4] = {0, 0, 0, 0}, yi;
            double yf;
            taps_of(y, rows, yi, yf);
            for (int r = 0; r < taps_; ++r)
                while (next_src <= rows[r]) {
                    bool needed = false;
                    for (int q = 0; q < taps_; ++q) needed |= rows[q] == next_src;
                    CiphertextBatch row = source(next_src, needed);
                    if (needed) {
                        if (!row.valid() || row.count() != (size_t)ow_ * 3 || row.size() != 2) throw std::invalid_argument("source row has the wrong shape");
                        src_.emplace(next_src, std::move(row));
                    }
                    ++next_src;
                }
        };
        load_for(y_begin);
        CiphertextBatch alive[2];  // the last two output rows: a sink may still be downloading them (fhe_download_async)
        for (int y = y_begin; y < y_end; ++y) {
            int rows[4] = {0, 0, 0, 0}, yint = 0;
            double yfr = 0;
            taps_of(y, rows, yint, yfr);
            // window: drop what no later row needs
            const int lowest = *std::min_element(rows, rows + taps_);
            for (auto it = src_.begin(); it != src_.end();) it = it->first < lowest ? src_.erase(it) : std::next(it);
            for (auto it = hor_.begin(); it != hor_.end();) it = it->first < lowest ? hor_.erase(it) : std::next(it);
            // The rows of THIS output row were requested one iteration ago: a source that uploads
            // asynchronously (fhe_upload_async) has had the previous row's kernels to hide the copy behind.
            // Wait for them on the stream, then request the next output row's source rows before queueing
            // this row's kernels.  A sink may be downloading asynchronously: the row before last is about to
            // be released (alive[] below), so all downloads but the most recent one must have landed.
            check(fhe_transfer_fence(ctx_.handle(), 1));
            if (y + 1 < y_end) load_for(y + 1);
            for (int r = 0; r < taps_; ++r)
                if (!src_.count(rows[r])) throw std::logic_error("source row left the window (bands must be fed from their first row)");
            Fraction fy;
            if (!det) {  // the reference's order: x then y for every output pixel (fhe_resize.h:230-234)
                CiphertextBatch xf(ctx_, (size_t)dw_, 2), yf(ctx_, (size_t)dw_, 2);
                Ciphertext c;
                for (int x = 0; x < dw_; ++x) {
                    cr_.encrypt(enc_.encode(xfr[(size_t)x]), c);
                    xf.set((size_t)x, c);
                    cr_.encrypt(enc_.encode(yfr), c);
                    yf.set((size_t)x, c);
                }
                fx = fraction(xf);
                fy = fraction(yf);
                hor_.clear();
            } else {
                CiphertextBatch yf(ctx_, 1, 2);
                Ciphertext c;
                cr_.encrypt(enc_.encode(yfr), c);
                yf.set(0, c);
                fy = fraction(yf);
            }
            // horizontal pass of every source row this output row touches
            for (int r = 0; r < taps_; ++r) {
                if (hor_.count(rows[r])) continue;
                const CiphertextBatch &row = src_.at(rows[r]);
                std::vector<const uint64_t *> tp(count * (size_t)taps_);
                for (int x = 0; x < dw_; ++x)
                    for (int c = 0; c < taps_; ++c) {
                        const int px = std::min(std::max(xint[(size_t)x] + off_ + c, 0), ow_ - 1);
                        for (int ch = 0; ch < 3; ++ch) tp[((size_t)x * 3 + ch) * taps_ + c] = row.element((size_t)px * 3 + ch);
                    }
                hor_.emplace(rows[r], graph(tp, 2, fx, 3, count));
            }
            // vertical pass
            std::vector<const uint64_t *> tp(count * (size_t)taps_);
            for (size_t e = 0; e < count; ++e)
                for (int r = 0; r < taps_; ++r) tp[e * taps_ + r] = hor_.at(rows[r]).element(e);
            CiphertextBatch out = graph(tp, hor_.begin()->second.size(), fy, det ? (uint32_t)count : 3u, count);
            sink(y, out);
            alive[y & 1] = std::move(out);
        }
        check(fhe_transfer_fence(ctx_.handle(), 0));  // asynchronous downloads of the last rows have landed
        src_.clear();
        hor_.clear();
    }

private:
    // a batch of encrypted fractions with everything the graphs multiply by, as operand forms
    struct Fraction {
        std::shared_ptr<OperandForm> t, t2, one_minus_t;
    };
    Fraction fraction(const CiphertextBatch &t)
    {
        Fraction f;
        f.t = std::make_shared<OperandForm>(ctx_, t);
        if (inter_ == BICUBIC) {
            // t2 = square(t); "t3" = multiply(t, t) (fhe_resize.h:174-175) is the same ciphertext
            CiphertextBatch t2(ctx_, t.count(), 3);
            check(fhe_multiply_form(ctx_.handle(), f.t->handle(), 1, f.t->handle(), 1, t2.device_data(), t.count()));
            f.t2 = std::make_shared<OperandForm>(ctx_, t2);
        } else {
            CiphertextBatch m(t);  // 1 - t: negate, add_plain(encode(1.0)) (fhe_resize.h:196)
            ev_.negate(m);
            ev_.add_plain(m, enc_.encode(1.0));
            f.one_minus_t = std::make_shared<OperandForm>(ctx_, m);
        }
        return f;
    }
    CiphertextBatch graph(const std::vector<const uint64_t *> &taps, uint32_t size_in, const Fraction &f, uint32_t t_div, size_t n_ct)
    {
        CiphertextBatch dest(ctx_, n_ct, size_in + (inter_ == BICUBIC ? 2u : 1u));
        if (inter_ == BICUBIC)
            check(fhe_cubic(ctx_.handle(), plan_->handle(), taps.data(), size_in, f.t->handle(), f.t2->handle(), f.t2->handle(), t_div,
                            dest.device_data(), n_ct));
        else
            check(fhe_linear(ctx_.handle(), taps.data(), size_in, f.one_minus_t->handle(), f.t->handle(), t_div, dest.device_data(), n_ct));
        return dest;
    }

    Evaluator &ev_;
    FractionalEncoder &enc_;
    Encryptor &cr_;
    const Context &ctx_;
    int ow_, oh_, dw_, dh_, inter_, taps_, off_;
    std::unique_ptr<ResizePlan> plan_;
    std::map<int, CiphertextBatch> src_, hor_;  // source rows in the window; their horizontal passes
};

struct SImageData {  // fhe_resize.h:207-213
    int width;
    int height;
    int start;
    std::vector<std::vector<Ciphertext>> pixels;
};

// ResizeImage's signature over the row engine, for images that fit in device memory
inline void ResizeImageBatched(const CiphertextSource &load, int original_width, int original_height, SImageData &destImage,
                               int dest_width, int dest_height, int inter, Evaluator &evaluator, FractionalEncoder &encoder,
                               Encryptor &encryptor)
{
    destImage.width = dest_width;
    destImage.height = dest_height;
    destImage.start = 0;
    destImage.pixels.clear();
    RowResizer rr(evaluator, encoder, encryptor, original_width, original_height, dest_width, dest_height, inter);
    rr.run(rows_of(load, original_width), [&](int, CiphertextBatch &row) {
        for (int x = 0; x < dest_width; ++x) {
            std::vector<Ciphertext> px;
            for (int ch = 0; ch < 3; ++ch) px.push_back(row.get((size_t)x * 3 + ch));
            destImage.pixels.push_back(std::move(px));
        }
    });
}

}  // namespace seal_b200
