// server_resize_b200.cpp -- the reference's homo/server_resize.cpp re-hosted on libfhe_b200
// (SURVEY.md 8f-N3): same files, flags and stdout label.
//
//   ./keys/params.txt   "W H C ..." (client_resize --send)           server_resize.cpp:15-23
//   --file/-f, --outfile/-o, --bicubic, --width/-w, --height/-h, --ncoeff, --fcoeff, --cmod, --pmod,
//   --base, --verbose/-v, --help                                       server_resize.cpp:42-87
//   input: RGB-interleaved row-major ciphertext stream (client_resize.cpp:141-154); output: the
//   resized pixels, 3 ciphertexts each (size 6 bicubic / 4 bilinear), server_resize.cpp:141-146
//   stdout: "Cubic,<ms>,<ms>,...\n" or "Linear,..." (server_resize.cpp:124-139)
//
// The pixel loop runs on the row engine of host/fhe_resize_b200.hpp (RowResizer: fused Cubic / Linear
// device graphs, fraction powers as operand forms, horizontal passes shared between output rows);
// source rows stream in from the input file and output rows stream out, so device memory holds a
// window of rows, not the image.  The ciphertexts written are the reference's bit for bit (its own
// homo/server_resize.cpp compiles unchanged against include/seal/seal.h: oracle/_ref/server_resize;
// tests compare both with the oracle).  The public fractions are encrypted with the key-less
// transparent Encryptor: this server loads neither the public key nor the secret key the reference
// reads "FOR DEBUGGING ONLY" (server_resize.cpp:99-116).
// stdout: the label, then ONE measured value per output row -- the pipeline's period for the row in ms
// divided by the Cubic / Linear calls the reference would have made for it (15 / 9 per pixel), repeated
// once per such call so that benchmark/analyze.py's per-label mean stays a per-call time.  The values
// are row averages, not individual call timings (the calls of a row run as five fused launches).
#include <chrono>
#include <cstdio>
#include <cstring>
#include <deque>
#include <fstream>
#include <sstream>

#include "../host/fhe_resize_b200.hpp"
#include "../host/seal_wire_b200.hpp"
#include "cli_b200.hpp"
#include "stream_b200.hpp"

using namespace seal_b200;

int main(int argc, const char **argv)
{
    cli_b200::Spec spec;
    spec.alias = {{"-f", "file"}, {"-o", "outfile"}, {"-v", "verbose"}, {"-w", "width"}, {"-h", "height"}};
    spec.valued = {"file", "outfile", "ncoeff", "fcoeff", "cmod", "pmod", "width", "height", "base", "wire"};
    spec.switches = {"bicubic", "verbose", "help"};
    const cli_b200::Parsed opt = cli_b200::parse(argc, argv, spec);
    if (opt.has("wire")) {  // header variant of the ciphertext records (host/seal_wire_b200.hpp)
        const std::string v = opt.str("wire", "capacity");
        if (v != "capacity" && v != "nocapacity") {
            std::cout << "error parsing options: --wire takes capacity or nocapacity" << std::endl;
            return 1;
        }
        wire_has_capacity() = v == "capacity";
    }
    if (opt.has("help")) {
        std::printf("%s [-f in] [-o out] [--bicubic] -w W -h H [--ncoeff N] [--fcoeff N] [--cmod n] [--pmod t] [--base b] "
                    "[-v]\n  ciphertext stream layout: host/seal_wire_b200.hpp (SEAL 2.3 layout from recall, UNVERIFIED against real SEAL files);\n  --wire nocapacity: records whose header has no size_capacity field\n", argv[0]);
        return 0;
    }
    int original_width = 0, original_height = 0, channels = 0;
    {
        std::ifstream paramfile("./keys/params.txt");
        paramfile >> original_width >> original_height >> channels;
        if (!paramfile || channels != 3 || original_width == 0 || original_height == 0) {
            std::fprintf(stderr, "server_resize_b200: ./keys/params.txt missing or invalid\n");
            return 1;
        }
    }
    if (!opt.has("width") || !opt.has("height")) {
        std::cout << "Please enter the width/height of the resized image..." << std::endl;
        std::cout << "\t Use --help to see the options." << std::endl;
        return 1;
    }
    const std::string ctext_infile = opt.str("file", "./image/nothingpersonnel.txt");
    const std::string ctext_outfile = opt.str("outfile", "./image/zoop.txt");
    const bool bicubic = opt.has("bicubic");
    const int resized_width = opt.num("width", 0), resized_height = opt.num("height", 0);
    try {
        Context context((uint32_t)opt.num("cmod", 8192), (uint64_t)opt.num("pmod", 1 << 14));
        Evaluator evaluator(context);
        FractionalEncoder encoder(context, (uint32_t)opt.num("ncoeff", 100), (uint32_t)opt.num("fcoeff", 100), 2);
        Encryptor encryptor(context);
        struct Fd {
            int fd;
            ~Fd() { if (fd >= 0) ::close(fd); }
        } infile{::open(ctext_infile.c_str(), O_RDONLY)}, outfile{::open(ctext_outfile.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644)};
        if (infile.fd < 0 || outfile.fd < 0) {
            std::fprintf(stderr, "server_resize_b200: cannot open %s or %s\n", ctext_infile.c_str(), ctext_outfile.c_str());
            return 1;
        }
        // Pipeline over rows (stream_b200.hpp): a reader thread parses the next source rows of the file into pinned
        // memory, the row engine uploads them one output row ahead of the kernels (fhe_upload_async), output rows
        // come back asynchronously (fhe_download_async) and a writer thread serialises them.
        const size_t W2 = context.ct_words(2), in_row_words = (size_t)original_width * 3 * W2;
        const uint32_t out_size = bicubic ? 6 : 4;
        const size_t out_row_words = (size_t)resized_width * 3 * context.ct_words(out_size);
        // records of one size: record i of the stream sits at i * record size; the helper threads read / write several at a time
        unsigned char head[48];
        if (!stream_b200::pread_exact(infile.fd, head, wire_header_bytes(), 0)) throw std::invalid_argument("ciphertext stream ended early");
        const WireHeader first = parse_record_header(head, context);
        if (first.size != 2) throw std::invalid_argument("input ciphertexts must have size 2");
        const hash_block_type hash = first.hash;
        const size_t in_rec = record_bytes(context, (uint32_t)first.size_capacity), out_rec = record_bytes(context, out_size);
        const size_t in_per_row = (size_t)original_width * 3, out_per_row = (size_t)resized_width * 3, Wout = context.ct_words(out_size);
        const unsigned io = stream_b200::io_threads();
        stream_b200::Reader reader(6, in_row_words, [&](long y, stream_b200::Slot &slot) {
            if (y >= original_height) return false;
            stream_b200::parallel_ranges(in_per_row, io, [&](size_t b, size_t e) {
                std::vector<unsigned char> rec(in_rec);
                for (size_t i = b; i < e; ++i) {
                    if (!stream_b200::pread_exact(infile.fd, rec.data(), in_rec, ((uint64_t)y * in_per_row + i) * in_rec))
                        throw std::invalid_argument("ciphertext stream ended early");
                    const WireHeader hdr = parse_record(rec.data(), context, slot.host + i * W2);
                    if (hdr.size != 2 || hdr.size_capacity != first.size_capacity)
                        throw std::invalid_argument("ciphertext stream: records of different sizes");
                }
            });
            slot.count = in_per_row;
            return true;
        });
        stream_b200::Writer writer(3, out_row_words, [&](stream_b200::Slot &slot) {
            stream_b200::parallel_ranges(slot.count, io, [&](size_t b, size_t e) {
                std::vector<unsigned char> rec(out_rec);
                for (size_t i = b; i < e; ++i) {
                    format_record(rec.data(), context, slot.host + i * Wout, out_size, hash);
                    stream_b200::pwrite_exact(outfile.fd, rec.data(), out_rec, ((uint64_t)slot.unit * out_per_row + i) * out_rec);
                }
            });
        });
        std::deque<stream_b200::Slot *> uploading, landing;  // pinned buffers with a transfer in flight, oldest first
        RowSource source = [&](int y, bool needed) {
            stream_b200::Slot *s = reader.next();
            if (!s || s->unit != y) throw std::invalid_argument("ciphertext stream ended early");
            CiphertextBatch row;
            if (!needed) {
                reader.recycle(s);
                return row;
            }
            row = CiphertextBatch(context, (size_t)original_width * 3, 2);
            check(fhe_upload_async(context.handle(), row.device_data(), s->host, in_row_words * sizeof(uint64_t)));
            uploading.push_back(s);
            // at most four uploads in flight (two output rows' worth): older buffers go back to the reader
            check(fhe_upload_fence(context.handle(), 4));
            while (uploading.size() > 4) {
                reader.recycle(uploading.front());
                uploading.pop_front();
            }
            return row;
        };
        std::cout << (bicubic ? "Cubic," : "Linear,");
        const int inter = bicubic ? BICUBIC : BILINEAR;
        const size_t calls_per_row = (size_t)resized_width * (bicubic ? 15 : 9);
        RowResizer resizer(evaluator, encoder, encryptor, original_width, original_height, resized_width, resized_height, inter);
        auto t0 = std::chrono::steady_clock::now();
        resizer.run(source, [&](int y, CiphertextBatch &row) {
            // the engine has just waited for every download but the newest (fhe_transfer_fence(ctx, 1)): the host
            // runs at most two rows ahead of the device, so the time between two calls is the pipeline's period
            const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            t0 = std::chrono::steady_clock::now();
            for (size_t i = 0; i < calls_per_row; ++i) std::cout << ms / calls_per_row << ',';
            while (landing.size() > 1) {
                writer.submit(landing.front());
                landing.pop_front();
            }
            if (row.size() != out_size) throw std::logic_error("unexpected output ciphertext size");
            stream_b200::Slot *o = writer.acquire();
            o->count = row.count();
            o->size = row.size();
            o->unit = y;
            check(fhe_download_async(context.handle(), o->host, row.device_data(), out_row_words * sizeof(uint64_t)));
            landing.push_back(o);
        });
        for (stream_b200::Slot *o : landing) writer.submit(o);  // run() returned after the last download landed
        writer.finish();
        std::cout << std::endl;
        return 0;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "server_resize_b200: %s\n", e.what());
        return 1;
    }
}
