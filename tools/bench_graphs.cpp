// bench_graphs.cpp -- in-process driver of the C++ host layer for bench.py's `resize` and `decode`
// workloads (BASELINE.json configs 3 and 5).  Built as tools/libbench_graphs.so (C ABI, loaded with
// ctypes) so that bench.py owns the barriers, the CUDA-event timing on its stream and the JSON, while
// the code being measured is the product's own host layer: RowResizer (host/fhe_resize_b200.hpp)
// and approximated_step_batched (host/fhe_decode_b200.hpp) over libfhe_b200.so.
//
// resize: output rows [y0, y1) of a bicubic / bilinear resize of an ow x oh RGB image of synthetic
//   ciphertexts (uniform residues, seeded like the oracle's generator), fractions encrypted by the
//   key-less transparent Encryptor.  Device mode: the source rows of the band are resident in HBM
//   before a step starts and the output rows stay on the device (only a checksum leaves it).  Host
//   mode: the source rows are in pinned host memory and every output row is copied back to pinned
//   host memory inside the step -- the server's real data path without the file I/O.
// decode: one run-length pair through approximated_step_batched (positions x degree Taylor series),
//   optionally with every product relinearised (synthetic evaluation keys: timing only).
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../fully-homomorphic-image-processing_b200/host/fhe_decode_b200.hpp"
#include "../fully-homomorphic-image-processing_b200/host/fhe_resize_b200.hpp"

using namespace seal_b200;

namespace {
thread_local std::string g_err;

struct ResizeBench {
    std::unique_ptr<Context> ctx;
    std::unique_ptr<Evaluator> ev;
    std::unique_ptr<FractionalEncoder> enc;
    std::unique_ptr<Encryptor> cr;
    int ow, oh, dw, dh, inter;
    bool host_mode;
    int y0 = 0, y1 = 0, first_src = 0, n_src = 0;
    CiphertextBatch resident;         // device mode: source rows first_src .. first_src + n_src - 1
    uint64_t *h_src = nullptr;        // host mode: the same rows, pinned
    uint64_t *h_out = nullptr;        // host mode: one output row, pinned
    size_t row_words = 0, out_row_words = 0;
    uint64_t last_checksum = 0;
};

struct DecodeBench {
    std::unique_ptr<Context> ctx;
    std::unique_ptr<Evaluator> ev;
    std::unique_ptr<FractionalEncoder> enc;
    std::unique_ptr<Encryptor> cr;
    std::unique_ptr<EvaluationKeys> evk;
    Ciphertext amplitude, index, count;
    int npos, degree, order;
    uint32_t out_size = 0;
};

// source rows a band touches: taps of its first and last output row (fhe_resize.h:351, clamped)
void band_rows(int oh, int dh, int inter, int y0, int y1, int &first, int &last)
{
    const int off = inter == BICUBIC ? -1 : 0, taps = inter == BICUBIC ? 4 : 2;
    first = oh;
    last = 0;
    for (int y = y0; y < y1; ++y) {
        const float v = float(y) / float(dh - 1) * float(oh) - 0.5;
        for (int r = 0; r < taps; ++r) {
            const int py = std::min(std::max(int(v) + off + r, 0), oh - 1);
            first = std::min(first, py);
            last = std::max(last, py);
        }
    }
}
}  // namespace

extern "C" {

const char *bench_graphs_error(void) { return g_err.c_str(); }

void *bres_create(int device, void *stream, uint32_t n, int ow, int oh, int dw, int dh, int inter, int host_mode, int y0, int y1)
{
    try {
        auto b = std::make_unique<ResizeBench>();
        b->ctx = std::make_unique<Context>(n, 1 << 14, std::vector<uint64_t>{}, device, stream);
        b->ev = std::make_unique<Evaluator>(*b->ctx);
        b->enc = std::make_unique<FractionalEncoder>(*b->ctx, 100, 100);
        b->cr = std::make_unique<Encryptor>(*b->ctx);
        b->ow = ow; b->oh = oh; b->dw = dw; b->dh = dh; b->inter = inter; b->host_mode = host_mode != 0;
        b->y0 = y0; b->y1 = y1;
        int last = 0;
        band_rows(oh, dh, inter, y0, y1, b->first_src, last);
        b->n_src = last - b->first_src + 1;
        b->row_words = b->ctx->ct_words(2) * (size_t)ow * 3;
        b->out_row_words = b->ctx->ct_words(inter == BICUBIC ? 6 : 4) * (size_t)dw * 3;
        // the band's source rows: ciphertext index = (row * ow + x) * 3 + channel of the whole image
        b->resident = CiphertextBatch(*b->ctx, (size_t)b->n_src * ow * 3, 2);
        check(fhe_fill_uniform(b->ctx->handle(), b->resident.device_data(), (uint64_t)b->first_src * ow * 3, (uint64_t)b->n_src * ow * 3, 2, 0xC0F3));
        if (b->host_mode) {
            void *p = nullptr;
            check(fhe_host_alloc(b->row_words * b->n_src * sizeof(uint64_t), &p));
            b->h_src = (uint64_t *)p;
            check(fhe_host_alloc(b->out_row_words * sizeof(uint64_t), &p));
            b->h_out = (uint64_t *)p;
            b->resident.download(b->h_src);
            b->resident = CiphertextBatch();
        }
        b->ctx->sync();
        return b.release();
    } catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}

// one step: the whole band.  checksum: order-sensitive checksum of every output row (device mode: of
// the device rows; host mode: of the rows as they arrived in host memory, first word of each row only
// hashed on the host to keep the timed region free of CPU work).
int bres_step(void *h, unsigned long long *checksum, unsigned long long *launches)
{
    ResizeBench *b = (ResizeBench *)h;
    try {
        const uint64_t l0 = fhe_launch_count(b->ctx->handle());
        uint64_t cs = 0;
        RowResizer rr(*b->ev, *b->enc, *b->cr, b->ow, b->oh, b->dw, b->dh, b->inter);
        rr.run(
            [&](int y, bool needed) {
                CiphertextBatch row;
                if (!needed || y < b->first_src || y >= b->first_src + b->n_src) return row;
                if (b->host_mode) {
                    row = CiphertextBatch(*b->ctx, (size_t)b->ow * 3, 2);
                    row.upload(b->h_src + (size_t)(y - b->first_src) * b->row_words);
                } else {
                    row = b->resident.slice((size_t)(y - b->first_src) * b->ow * 3, (size_t)b->ow * 3);
                }
                return row;
            },
            [&](int, CiphertextBatch &row) {
                if (b->host_mode) {
                    row.download(b->h_out);
                    cs = cs * 0x9E3779B97F4A7C15ull + b->h_out[0] + b->h_out[b->out_row_words - 1];
                } else {
                    uint64_t c = 0;
                    check(fhe_checksum(b->ctx->handle(), row.device_data(), b->out_row_words, &c));
                    cs = cs * 0x9E3779B97F4A7C15ull + c;
                }
            },
            b->y0, b->y1);
        b->last_checksum = cs;
        if (checksum) *checksum = cs;
        if (launches) *launches = fhe_launch_count(b->ctx->handle()) - l0;
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        return 1;
    }
}
// copies of one output ciphertext and of its 16 (4) x 3 source taps for the oracle spot check:
// output pixel (x, y) of the band, channel ch -> out (size 6 / 4), and the source rows the band holds
// to `src` (n_src rows x ow x 3 ciphertexts, size 2) when src is not null
int bres_spot(void *h, int x, int y, int ch, uint64_t *out, uint64_t *src, int *first_src, int *n_src)
{
    ResizeBench *b = (ResizeBench *)h;
    try {
        if (first_src) *first_src = b->first_src;
        if (n_src) *n_src = b->n_src;
        if (src) {
            if (b->host_mode) std::memcpy(src, b->h_src, b->row_words * b->n_src * sizeof(uint64_t));
            else b->resident.download(src);
        }
        if (out) {
            RowResizer rr(*b->ev, *b->enc, *b->cr, b->ow, b->oh, b->dw, b->dh, b->inter);
            rr.run(
                [&](int yy, bool needed) {
                    CiphertextBatch row;
                    if (!needed || yy < b->first_src || yy >= b->first_src + b->n_src) return row;
                    if (b->host_mode) {
                        row = CiphertextBatch(*b->ctx, (size_t)b->ow * 3, 2);
                        row.upload(b->h_src + (size_t)(yy - b->first_src) * b->row_words);
                    } else {
                        row = b->resident.slice((size_t)(yy - b->first_src) * b->ow * 3, (size_t)b->ow * 3);
                    }
                    return row;
                },
                [&](int yy, CiphertextBatch &row) {
                    if (yy == y) row.get((size_t)x * 3 + ch).download(out);
                },
                y, y + 1);
        }
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        return 1;
    }
}
void bres_destroy(void *h)
{
    ResizeBench *b = (ResizeBench *)h;
    if (!b) return;
    if (b->h_src) fhe_host_free(b->h_src);
    if (b->h_out) fhe_host_free(b->h_out);
    delete b;
}

void *bdec_create(int device, void *stream, uint32_t n, int npos, int degree, int order, uint32_t dbc)
{
    try {
        auto b = std::make_unique<DecodeBench>();
        b->ctx = std::make_unique<Context>(n, 1 << 14, std::vector<uint64_t>{}, device, stream);
        b->ev = std::make_unique<Evaluator>(*b->ctx);
        b->enc = std::make_unique<FractionalEncoder>(*b->ctx, 100, 100);
        b->cr = std::make_unique<Encryptor>(*b->ctx);
        b->npos = npos; b->degree = degree; b->order = order;
        if (dbc) {
            // timing only: any residues below every q_j serve as keys
            const uint32_t k = b->ctx->coeff_mod_count(), digits = fhe_evk_digits(b->ctx->handle(), dbc);
            std::vector<uint64_t> kw((size_t)k * digits * 2 * k * n);
            uint64_t z = 0x5EED;
            for (auto &w : kw) { z = z * 6364136223846793005ull + 1442695040888963407ull; w = z >> 12; }
            b->evk = std::make_unique<EvaluationKeys>(*b->ctx, kw.data(), dbc, 1);
            b->ev->set_relinearization(b->evk.get());
        }
        b->amplitude = Ciphertext(*b->ctx, 2); b->index = Ciphertext(*b->ctx, 2); b->count = Ciphertext(*b->ctx, 2);
        check(fhe_fill_uniform(b->ctx->handle(), b->amplitude.device_data(), 0, 1, 2, 0xDEC));
        check(fhe_fill_uniform(b->ctx->handle(), b->index.device_data(), 1, 1, 2, 0xDEC));
        check(fhe_fill_uniform(b->ctx->handle(), b->count.device_data(), 2, 1, 2, 0xDEC));
        b->ctx->sync();
        return b.release();
    } catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}
// one run-length pair; first_out (optional): the first output ciphertext (out_size polynomials)
int bdec_step(void *h, unsigned long long *launches, unsigned int *out_size, uint64_t *first_out)
{
    DecodeBench *b = (DecodeBench *)h;
    try {
        const uint64_t l0 = fhe_launch_count(b->ctx->handle());
        std::vector<Ciphertext> run;
        approximated_step_batched(b->amplitude, b->index, b->count, b->order, b->degree, 0.5, b->npos, 1, run, *b->ev, *b->enc, *b->cr);
        b->out_size = run.at(0).size();
        if (launches) *launches = fhe_launch_count(b->ctx->handle()) - l0;
        if (out_size) *out_size = b->out_size;
        if (first_out) run[0].download(first_out);
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        return 1;
    }
}
// the three inputs (size 2 each), for the oracle spot check
int bdec_inputs(void *h, uint64_t *amplitude, uint64_t *index, uint64_t *count)
{
    DecodeBench *b = (DecodeBench *)h;
    try {
        b->amplitude.download(amplitude);
        b->index.download(index);
        b->count.download(count);
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        return 1;
    }
}
void bdec_destroy(void *h) { delete (DecodeBench *)h; }

}  // extern "C"
