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
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include <cuda_runtime_api.h>

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
    int y0 = 0, y1 = 0, first_src = 0, n_src = 0, n_host = 0;
    CiphertextBatch resident;         // device mode: source rows first_src .. first_src + n_src - 1
    uint64_t *h_src = nullptr;        // host mode: the same rows, pinned
    uint64_t *h_out = nullptr;        // host mode: two output rows, pinned (double buffer of the asynchronous downloads)
    size_t row_words = 0, out_row_words = 0;
    uint64_t last_checksum = 0;
    cudaStream_t stream = nullptr;    // the stream the context works on (BRES_TRACE)
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
    bool host_mode = false;
    uint64_t *h_in = nullptr, *h_out = nullptr;
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

// host_rows (host mode): pinned source rows to hold, 0 = every row the band touches; with fewer, source row y
// is read from slot (y - first) % host_rows -- the data is synthetic, so a long band (whose fill and drain
// are amortised as they are over a whole image) need not pin tens of GB; the spot check reads the same slots
void *bres_create(int device, void *stream, uint32_t n, int ow, int oh, int dw, int dh, int inter, int host_mode, int y0, int y1, int host_rows)
{
    try {
        auto b = std::make_unique<ResizeBench>();
        b->ctx = std::make_unique<Context>(n, 1 << 14, std::vector<uint64_t>{}, device, stream);
        b->stream = (cudaStream_t)stream;
        b->ev = std::make_unique<Evaluator>(*b->ctx);
        b->enc = std::make_unique<FractionalEncoder>(*b->ctx, 100, 100);
        b->cr = std::make_unique<Encryptor>(*b->ctx);
        b->ow = ow; b->oh = oh; b->dw = dw; b->dh = dh; b->inter = inter; b->host_mode = host_mode != 0;
        b->y0 = y0; b->y1 = y1;
        int last = 0;
        band_rows(oh, dh, inter, y0, y1, b->first_src, last);
        b->n_src = last - b->first_src + 1;
        b->n_host = host_mode && host_rows > 0 && host_rows < b->n_src ? host_rows : b->n_src;
        b->row_words = b->ctx->ct_words(2) * (size_t)ow * 3;
        b->out_row_words = b->ctx->ct_words(inter == BICUBIC ? 6 : 4) * (size_t)dw * 3;
        // the band's source rows: ciphertext index = (row * ow + x) * 3 + channel of the whole image
        b->resident = CiphertextBatch(*b->ctx, (size_t)b->n_host * ow * 3, 2);
        check(fhe_fill_uniform(b->ctx->handle(), b->resident.device_data(), (uint64_t)b->first_src * ow * 3, (uint64_t)b->n_host * ow * 3, 2, 0xC0F3));
        if (b->host_mode) {
            void *p = nullptr;
            check(fhe_host_alloc(b->row_words * b->n_host * sizeof(uint64_t), &p));
            b->h_src = (uint64_t *)p;
            check(fhe_host_alloc(2 * b->out_row_words * sizeof(uint64_t), &p));
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

// one step: the whole band.  checksum: device mode: of the band's last output row (computed on the
// device, the step's one synchronisation); host mode: a hash of the first and last word of the last two
// rows as they arrived in host memory (no CPU pass over the data inside the timed region).
int bres_step(void *h, unsigned long long *checksum, unsigned long long *launches)
{
    ResizeBench *b = (ResizeBench *)h;
    try {
        const uint64_t l0 = fhe_launch_count(b->ctx->handle());
        uint64_t cs = 0;
        // BRES_TRACE=1 (diagnostic, stderr): per output row, how long the stream waited for the row's
        // uploads after the previous row's kernels ("stall") and how long its own kernels ran ("busy").
        // The engine requests row y + 1's sources right before it queues row y's kernels, after making the
        // stream wait for row y's uploads: an event recorded then fires when row y can start.
        const bool trace = std::getenv("BRES_TRACE") != nullptr && b->stream;
        std::vector<cudaEvent_t> t_start, t_end;  // t_start[k]: the last mark recorded while k rows had been emitted
        auto mark = [&](cudaEvent_t &e) {
            if (!e) cudaEventCreate(&e);
            cudaEventRecord(e, b->stream);
        };
        RowResizer rr(*b->ev, *b->enc, *b->cr, b->ow, b->oh, b->dw, b->dh, b->inter);
        rr.run(
            [&](int y, bool needed) {
                if (trace) {  // all marks of one batch of requests fire together: keep the last one per emitted-row count
                    if (t_start.size() <= t_end.size()) t_start.resize(t_end.size() + 1, nullptr);
                    mark(t_start[t_end.size()]);
                }
                CiphertextBatch row;
                if (!needed || y < b->first_src || y >= b->first_src + b->n_src) return row;
                if (b->host_mode) {  // pinned host -> device on the copy stream, behind the previous row's kernels
                    row = CiphertextBatch(*b->ctx, (size_t)b->ow * 3, 2);
                    check(fhe_upload_async(b->ctx->handle(), row.device_data(), b->h_src + (size_t)((y - b->first_src) % b->n_host) * b->row_words,
                                           b->row_words * sizeof(uint64_t)));
                } else {
                    row = b->resident.slice((size_t)(y - b->first_src) * b->ow * 3, (size_t)b->ow * 3);
                }
                return row;
            },
            [&](int y, CiphertextBatch &row) {
                if (trace) {
                    t_end.push_back(nullptr);
                    mark(t_end.back());
                }
                if (b->host_mode) {  // device -> pinned host on the other copy stream, beside the next row's kernels
                    check(fhe_download_async(b->ctx->handle(), b->h_out + (size_t)(y & 1) * b->out_row_words, row.device_data(),
                                             b->out_row_words * sizeof(uint64_t)));
                } else if (y == b->y1 - 1) {  // one synchronisation per step: the checksum of the last row
                    check(fhe_checksum(b->ctx->handle(), row.device_data(), b->out_row_words, &cs));
                }
            },
            b->y0, b->y1);
        if (b->host_mode)  // run() returned after the last download landed: hash the two rows still in the buffers
            for (int s2 = 0; s2 < 2; ++s2) cs = cs * 0x9E3779B97F4A7C15ull + b->h_out[(size_t)s2 * b->out_row_words] + b->h_out[(size_t)(s2 + 1) * b->out_row_words - 1];
        if (trace) {
            cudaStreamSynchronize(b->stream);
            for (size_t i = 0; i < t_end.size(); ++i) {
                float busy = 0, stall = 0;
                if (i < t_start.size() && t_start[i]) cudaEventElapsedTime(&busy, t_start[i], t_end[i]);
                if (i >= 1 && i < t_start.size() && t_start[i]) cudaEventElapsedTime(&stall, t_end[i - 1], t_start[i]);
                std::fprintf(stderr, "[bres trace %s] row %d: stall %.1f ms, busy %.1f ms\n", b->host_mode ? "host" : "device", b->y0 + (int)i, stall, busy);
            }
            for (auto e : t_start) if (e) cudaEventDestroy(e);
            for (auto e : t_end) cudaEventDestroy(e);
        }
        b->last_checksum = cs;
        if (checksum) *checksum = cs;
        if (launches) *launches = fhe_launch_count(b->ctx->handle()) - l0;
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        return 1;
    }
}
// Oracle spot check of one output value of the band: output pixel (x, y), channel ch -> `out` (size 6 /
// 4); `taps`: its taps x taps source ciphertexts (size 2, row-major: tap row r, tap column c); the two
// fractions the server encrypts for it.  Runs the engine on the single row y.
int bres_spot(void *h, int x, int y, int ch, uint64_t *out, uint64_t *taps, double *xfrac, double *yfrac)
{
    ResizeBench *b = (ResizeBench *)h;
    try {
        const int off = b->inter == BICUBIC ? -1 : 0, nt = b->inter == BICUBIC ? 4 : 2;
        const float u = float(x) / float(b->dw - 1) * float(b->ow) - 0.5, v = float(y) / float(b->dh - 1) * float(b->oh) - 0.5;
        *xfrac = u - std::floor(u);
        *yfrac = v - std::floor(v);
        const size_t W = b->ctx->ct_words(2);
        std::vector<uint64_t> row_host;
        auto source = [&](int yy, bool needed) {
            CiphertextBatch row;
            if (!needed || yy < b->first_src || yy >= b->first_src + b->n_src) return row;
            if (b->host_mode) {
                row = CiphertextBatch(*b->ctx, (size_t)b->ow * 3, 2);
                row.upload(b->h_src + (size_t)((yy - b->first_src) % b->n_host) * b->row_words);
            } else {
                row = b->resident.slice((size_t)(yy - b->first_src) * b->ow * 3, (size_t)b->ow * 3);
            }
            return row;
        };
        for (int r = 0; r < nt; ++r) {
            const int py = std::min(std::max(int(v) + off + r, 0), b->oh - 1);
            CiphertextBatch row = source(py, true);
            if (!row.valid()) throw std::runtime_error("spot row outside the band");
            for (int c = 0; c < nt; ++c) {
                const int px = std::min(std::max(int(u) + off + c, 0), b->ow - 1);
                check(fhe_download(b->ctx->handle(), taps + ((size_t)r * nt + c) * W, row.element((size_t)px * 3 + ch), W * sizeof(uint64_t)));
            }
        }
        RowResizer rr(*b->ev, *b->enc, *b->cr, b->ow, b->oh, b->dw, b->dh, b->inter);
        rr.run(source, [&](int yy, CiphertextBatch &row) {
            if (yy == y) row.get((size_t)x * 3 + ch).download(out);
        }, y, y + 1);
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        return 1;
    }
}
void *bres_ctx(void *h) { return ((ResizeBench *)h)->ctx->handle(); }
void bres_destroy(void *h)
{
    ResizeBench *b = (ResizeBench *)h;
    if (!b) return;
    if (b->h_src) fhe_host_free(b->h_src);
    if (b->h_out) fhe_host_free(b->h_out);
    delete b;
}

// keys: evaluation keys in coefficient form (fhe_evk_create layout) or null (no relinearisation: the
// reference's graph); host_mode: the three inputs come from pinned host memory and all outputs go back
// to pinned host memory inside the step
void *bdec_create(int device, void *stream, uint32_t n, int npos, int degree, int order, uint32_t dbc, const uint64_t *keys,
                  int host_mode, uint64_t first_ct)
{
    try {
        auto b = std::make_unique<DecodeBench>();
        b->ctx = std::make_unique<Context>(n, 1 << 14, std::vector<uint64_t>{}, device, stream);
        b->ev = std::make_unique<Evaluator>(*b->ctx);
        b->enc = std::make_unique<FractionalEncoder>(*b->ctx, 100, 100);
        b->cr = std::make_unique<Encryptor>(*b->ctx);
        b->npos = npos; b->degree = degree; b->order = order; b->host_mode = host_mode != 0;
        if (dbc && keys) {
            b->evk = std::make_unique<EvaluationKeys>(*b->ctx, keys, dbc, 1);
            b->ev->set_relinearization(b->evk.get());
        }
        b->amplitude = Ciphertext(*b->ctx, 2); b->index = Ciphertext(*b->ctx, 2); b->count = Ciphertext(*b->ctx, 2);
        check(fhe_fill_uniform(b->ctx->handle(), b->amplitude.device_data(), first_ct + 0, 1, 2, 0xDEC));
        check(fhe_fill_uniform(b->ctx->handle(), b->index.device_data(), first_ct + 1, 1, 2, 0xDEC));
        check(fhe_fill_uniform(b->ctx->handle(), b->count.device_data(), first_ct + 2, 1, 2, 0xDEC));
        if (b->host_mode) {
            const size_t W = b->ctx->ct_words(2);
            void *p = nullptr;
            check(fhe_host_alloc(3 * W * sizeof(uint64_t), &p));
            b->h_in = (uint64_t *)p;
            b->amplitude.download(b->h_in);
            b->index.download(b->h_in + W);
            b->count.download(b->h_in + 2 * W);
            check(fhe_host_alloc((size_t)npos * b->ctx->ct_words(b->evk ? 2 : 22) * sizeof(uint64_t), &p));
            b->h_out = (uint64_t *)p;
        }
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
        const size_t W = b->ctx->ct_words(2);
        if (b->host_mode) {
            b->amplitude.upload(b->h_in);
            b->index.upload(b->h_in + W);
            b->count.upload(b->h_in + 2 * W);
        }
        std::vector<Ciphertext> run;
        approximated_step_batched(b->amplitude, b->index, b->count, b->order, b->degree, 0.5, b->npos, 1, run, *b->ev, *b->enc, *b->cr);
        b->out_size = run.at(0).size();
        if (b->host_mode)
            for (size_t i = 0; i < run.size(); ++i) run[i].download(b->h_out + i * b->ctx->ct_words(b->out_size));
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
void *bdec_ctx(void *h) { return ((DecodeBench *)h)->ctx->handle(); }
void bdec_destroy(void *h)
{
    DecodeBench *b = (DecodeBench *)h;
    if (!b) return;
    if (b->h_in) fhe_host_free(b->h_in);
    if (b->h_out) fhe_host_free(b->h_out);
    delete b;
}

}  // extern "C"
