// server_jpeg_b200.cpp -- the reference's homo/server_jpeg.cpp re-hosted on libfhe_b200
// (SURVEY.md 8f-N3).  Same place in the client/server split, same files and flags:
//
//   ./keys/params.txt            "W H C pmod" written by client_jpeg --send (client_jpeg.cpp:88-94)
//   --file / -f   (default ./image/nothingpersonnel.txt)   per 8x8 tile: R64, G64, B64 ciphertexts
//   --outfile / -o (default ./image/zoop.txt)               per tile: Y64, Cb64, Cr64 (:146-153)
//   --ncoeff --fcoeff --cmod --pmod --base                  as server_jpeg.cpp:44-52 (--base is parsed
//                                                           and, like there (:66), has no effect: base 2)
//   stdout: "RGBYCC,<64 ms values>,\nDCT,<3 ms values>,\n" per tile (server_jpeg.cpp:126-137) -- the CSV
//   benchmark/analyze.py parses.  The values are the fused kernels' device times split evenly over
//   the calls they replace (the colour conversion is the epilogue of the forward transform).
//
// Unlike the reference (one tile at a time, one Evaluator call at a time, "three blocks to not get
// OOM", :109-111) tiles are streamed through the device in chunks, as a pipeline (stream_b200.hpp): a reader
// thread parses the next chunk of the file, ONE fhe_jpeg_rgb_tiles call per chunk (forward NTT + colour
// conversion, pointwise DCT per channel, inverse NTT) with the neighbouring chunks' copies beside it, a
// writer thread serialises the results.  The server never loads the secret key (the reference does, "FOR
// DEBUGGING ONLY", :97-98) and needs no public key on this path.
// Extra flags: --quantize (also apply quantize_fhe with the JPEG luminance / chrominance tables,
// fhe_image.h:99-100,294-305 -- the reference's main never calls it), --chunk N (tiles per call).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <fstream>
#include <iostream>
#include <string>

#include "../host/fhe_image_b200.hpp"
#include "../host/seal_wire_b200.hpp"
#include "stream_b200.hpp"

using namespace seal_b200;

static const double kYQT[64] = {16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57,
                                69, 56, 14, 17, 22, 29, 51, 87, 80, 62, 18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64,
                                81, 104, 113, 92, 49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
static const double kUVQT[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99,
                                 99, 99, 47, 66, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
                                 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};
static const double kColor[9] = {0.299, 0.587, 0.114, -0.168736, 0.331264, 0.5, 0.5, 0.418688, 0.081312};

static void usage(const char *argv0)
{
    std::printf("%s [-f|--file in] [-o|--outfile out] [--ncoeff N] [--fcoeff N] [--cmod n] [--pmod t] [--base b]\n"
                "        [--quantize] [--chunk tiles] [--wire capacity|nocapacity] [-v|--verbose] [--help]\n"
                "  ciphertext stream layout: host/seal_wire_b200.hpp (SEAL 2.3 layout from recall, UNVERIFIED against real SEAL files);\n"
                "  --wire nocapacity reads and writes records whose header has no size_capacity field\n", argv0);
}

int main(int argc, const char **argv)
{
    std::string ctext_infile("./image/nothingpersonnel.txt"), ctext_outfile("./image/zoop.txt");
    bool verbose = false, quantize = false;
    int n_number_coeffs = 100, n_fractional_coeffs = 100, n_poly_base = 2;  // fhe_image.h:22-26
    int plain_modulus = 1 << 14, coeff_modulus = 8192;                       // fhe_image.h:26-27
    int chunk_tiles = 64;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto val = [&](const char *name) -> std::string {
            if (i + 1 >= argc) {
                std::cout << "error parsing options: Option '" << name << "' is missing an argument" << std::endl;
                std::exit(1);
            }
            return argv[++i];
        };
        if (a == "-f" || a == "--file") ctext_infile = val("file");
        else if (a == "-o" || a == "--outfile") ctext_outfile = val("outfile");
        else if (a == "-v" || a == "--verbose") verbose = true;
        else if (a == "--ncoeff") n_number_coeffs = std::atoi(val("ncoeff").c_str());
        else if (a == "--fcoeff") n_fractional_coeffs = std::atoi(val("fcoeff").c_str());
        else if (a == "--cmod") coeff_modulus = std::atoi(val("cmod").c_str());
        else if (a == "--pmod") plain_modulus = std::atoi(val("pmod").c_str());
        else if (a == "--base") n_poly_base = std::atoi(val("base").c_str());
        else if (a == "--quantize") quantize = true;
        else if (a == "--chunk") chunk_tiles = std::atoi(val("chunk").c_str());
        else if (a == "--wire") {
            const std::string v = val("wire");
            if (v != "capacity" && v != "nocapacity") {
                std::cout << "error parsing options: --wire takes capacity or nocapacity" << std::endl;
                return 1;
            }
            wire_has_capacity() = v == "capacity";
        }
        else if (a == "--help") {
            usage(argv[0]);
            return 0;
        } else {
            std::cout << "error parsing options: Option '" << a << "' does not exist" << std::endl;
            return 1;
        }
    }
    (void)n_poly_base;  // the reference's override is dead code (server_jpeg.cpp:66): base stays 2
    if (chunk_tiles < 1) chunk_tiles = 1;

    int WIDTH = 0, HEIGHT = 0, channels = 0;
    {
        std::ifstream paramfile("./keys/params.txt");
        paramfile >> WIDTH >> HEIGHT >> channels;
        if (!paramfile || channels != 3) {
            std::fprintf(stderr, "server_jpeg_b200: ./keys/params.txt missing or channels != 3\n");
            return 1;
        }
    }
    try {
        Context context((uint32_t)coeff_modulus, (uint64_t)plain_modulus);
        FractionalEncoder encoder(context, (uint32_t)n_number_coeffs, (uint32_t)n_fractional_coeffs, 2);
        const uint32_t n = context.poly_degree();
        auto stack = [&](const double *vals, int count, bool reciprocal) {
            std::vector<uint64_t> out((size_t)count * n);
            for (int i = 0; i < count; ++i) {
                Plaintext p = encoder.encode(reciprocal ? 1 / vals[i] : vals[i]);
                std::memcpy(out.data() + (size_t)i * n, p.coeffs().data(), n * sizeof(uint64_t));
            }
            return out;
        };
        const std::vector<uint64_t> color = stack(kColor, 9, false), dct = stack(kDctConstants, 12, false);
        const std::vector<uint64_t> off = encoder.encode(128.0).coeffs(), eighth = encoder.encode(0.125).coeffs();
        std::vector<uint64_t> qy, qc;
        if (quantize) {
            qy = stack(kYQT, 64, true);
            qc = stack(kUVQT, 64, true);
        }
        fhe_jpeg_plan *plan = nullptr;
        check(fhe_jpeg_plan_create(context.handle(), color.data(), off.data(), dct.data(), eighth.data(),
                                   quantize ? qy.data() : nullptr, quantize ? qc.data() : nullptr, &plan));

        const int num_blocks = WIDTH * HEIGHT / 64;  // server_jpeg.cpp:102 (tiles: R, G and B block each)
        struct Fd {
            int fd;
            ~Fd() { if (fd >= 0) ::close(fd); }
        } myfile{::open(ctext_infile.c_str(), O_RDONLY)}, outfile{::open(ctext_outfile.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644)};
        if (myfile.fd < 0 || outfile.fd < 0) {
            std::fprintf(stderr, "server_jpeg_b200: cannot open %s or %s\n", ctext_infile.c_str(), ctext_outfile.c_str());
            return 1;
        }
        const size_t W = context.ct_words(2), tile_words = 192 * W;
        const size_t chunk_bytes = (size_t)chunk_tiles * tile_words * 8;
        // Pipeline over chunks of tiles (stream_b200.hpp): the reader thread parses chunk c + 1 into pinned memory
        // while chunk c is on the device and the writer thread serialises chunk c - 2; on the device, the upload of
        // chunk c + 1 and the download of chunk c - 1 run beside the kernels of chunk c (three device buffers).
        // The records of a stream have one size (the client writes them with one set of parameters): the first header
        // gives it, and the helper threads read / write record i at i * record size, several at a time.
        unsigned char head[48];
        if (!stream_b200::pread_exact(myfile.fd, head, wire_header_bytes(), 0)) throw std::invalid_argument("ciphertext stream ended early");
        const WireHeader first = parse_record_header(head, context);
        if (first.size != 2) throw std::invalid_argument("input ciphertexts must have size 2");
        const hash_block_type hash = first.hash;
        const size_t in_rec = record_bytes(context, (uint32_t)first.size_capacity), out_rec = record_bytes(context, 2);
        const unsigned io = stream_b200::io_threads();
        stream_b200::Reader reader(3, (size_t)chunk_tiles * tile_words, [&](long u, stream_b200::Slot &slot) {
            const long t0 = u * (long)chunk_tiles;
            if (t0 >= num_blocks) return false;
            const int nt = (int)std::min<long>(chunk_tiles, num_blocks - t0);
            stream_b200::parallel_ranges((size_t)nt * 192, io, [&](size_t b, size_t e) {
                std::vector<unsigned char> rec(in_rec);
                for (size_t i = b; i < e; ++i) {
                    if (!stream_b200::pread_exact(myfile.fd, rec.data(), in_rec, ((uint64_t)t0 * 192 + i) * in_rec))
                        throw std::invalid_argument("ciphertext stream ended early");
                    const WireHeader hdr = parse_record(rec.data(), context, slot.host + i * W);
                    if (hdr.size != 2 || hdr.size_capacity != first.size_capacity)
                        throw std::invalid_argument("ciphertext stream: records of different sizes");
                }
            });
            slot.count = (size_t)nt;
            return true;
        });
        stream_b200::Writer writer(3, (size_t)chunk_tiles * tile_words, [&](stream_b200::Slot &slot) {
            stream_b200::parallel_ranges(slot.count * 192, io, [&](size_t b, size_t e) {
                std::vector<unsigned char> rec(out_rec);
                for (size_t i = b; i < e; ++i) {
                    format_record(rec.data(), context, slot.host + i * W, 2, hash);
                    stream_b200::pwrite_exact(outfile.fd, rec.data(), out_rec, ((uint64_t)slot.unit * chunk_tiles * 192 + i) * out_rec);
                }
            });
        });
        void *dbuf[3] = {nullptr, nullptr, nullptr};
        for (auto &d : dbuf) check(fhe_device_alloc(context.handle(), chunk_bytes, &d));
        check(fhe_set_stage_timing(context.handle(), 0));
        std::deque<stream_b200::Slot *> landing;  // output buffers whose downloads are in flight, oldest first
        stream_b200::Slot *cur = reader.next();
        if (cur) check(fhe_upload_async(context.handle(), dbuf[0], cur->host, cur->count * tile_words * 8));
        for (long c = 0; cur; ++c) {
            const int nt = (int)cur->count;
            const auto s0 = std::chrono::steady_clock::now();
            // the stream waits for the upload of chunk c; every download but the newest has landed
            check(fhe_transfer_fence(context.handle(), 1));
            while (landing.size() > 1) {
                writer.submit(landing.front());
                landing.pop_front();
            }
            stream_b200::Slot *nxt = reader.next();  // upload of chunk c + 1 beside the kernels of chunk c; its buffer
            if (nxt)                                 // held chunk c - 2, whose download has landed
                check(fhe_upload_async(context.handle(), dbuf[(c + 1) % 3], nxt->host, nxt->count * tile_words * 8));
            void *d = dbuf[c % 3];
            check(fhe_profile_begin(context.handle()));
            check(fhe_jpeg_rgb_tiles(context.handle(), plan, (const uint64_t *)d, (uint64_t *)d, (uint64_t)nt));
            // CSV contract (benchmark/analyze.py: per-label mean of the values): MEASURED device times of this
            // chunk's kernels -- RGBYCC: the colour-conversion pass per pixel; DCT: forward transforms +
            // pointwise DCT (+quantize) + inverse transforms per channel block (the transforms serve both
            // steps and are counted under DCT).  Chunk averages, copies excluded (-v prints the wall time).
            char names[512];
            float kms[16];
            uint64_t kcount[16];
            uint32_t nk = 0;
            check(fhe_profile_end(context.handle(), names, sizeof names, kms, kcount, 16, &nk));  // the chunk's kernels are done
            reader.recycle(cur);  // and with them the upload they waited for
            double color_ms = 0, dct_ms = 0;
            {
                const char *p = names;
                for (uint32_t i = 0; i < nk; ++i) {
                    const char *e = std::strchr(p, '\n');
                    const std::string name(p, e ? (size_t)(e - p) : std::strlen(p));
                    (name == "k_rgb2ycc_ntt" ? color_ms : dct_ms) += kms[i];
                    p = e ? e + 1 : p + name.size();
                }
            }
            stream_b200::Slot *o = writer.acquire();
            o->count = (size_t)nt;
            o->unit = c;
            check(fhe_download_async(context.handle(), o->host, d, (size_t)nt * tile_words * 8));
            landing.push_back(o);
            const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - s0).count();
            for (int t = 0; t < nt; ++t) {
                std::cout << "RGBYCC,";
                for (int i = 0; i < 64; ++i) std::cout << color_ms / (64.0 * nt) << ',';
                std::cout << std::endl << "DCT,";
                for (int i = 0; i < 3; ++i) std::cout << dct_ms / (3.0 * nt) << ',';
                std::cout << std::endl;
            }
            if (verbose) std::fprintf(stderr, "tiles %ld..%ld: %.3f ms\n", c * chunk_tiles, c * chunk_tiles + nt - 1, ms);
            cur = nxt;
        }
        check(fhe_transfer_fence(context.handle(), 0));
        for (stream_b200::Slot *o : landing) writer.submit(o);
        writer.finish();
        for (auto &d : dbuf) fhe_device_free(context.handle(), d);
        fhe_jpeg_plan_destroy(plan);
        return 0;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "server_jpeg_b200: %s\n", e.what());
        return 1;
    }
}
