"""server_jpeg_b200 (fully-homomorphic-image-processing_b200/server/server_jpeg_b200.cpp): the
reference's homo/server_jpeg.cpp re-hosted on libfhe_b200 -- same files (./keys/params.txt, ciphertext
stream in, stream out), same flags, same stdout CSV labels -- plus the SEAL-2.3-style wire format of
host/seal_wire_b200.hpp (unverified against real SEAL, SEAL is absent).
CPU: builds, rejects bad flags like cxxopts does, fails loudly without a device.
GPU: a client-side stream made with the oracle's encryptor goes through the binary; the output
stream equals the oracle's op-by-op rgb_to_ycc_fhe + encrypted_dct bit for bit and decrypts to the
plaintext pipeline's DCT coefficients."""
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import PKG_NAME, ROOT

SRC = os.path.join(ROOT, PKG_NAME, "server", "server_jpeg_b200.cpp")
EXE = os.path.join(ROOT, PKG_NAME, "server", "server_jpeg_b200")
HASH = (0x1111111111111111, 0x2222222222222222, 0x3333333333333333, 0x4444444444444444)


def build_exe(pkg):
    pkg.build()
    libdir = os.path.join(ROOT, PKG_NAME)
    deps = [SRC, os.path.join(ROOT, "include", "fhe_b200.h")] + [
        os.path.join(libdir, "host", f) for f in ("seal_b200.hpp", "fhe_image_b200.hpp", "seal_wire_b200.hpp")] + [
        os.path.join(libdir, "server", "stream_b200.hpp")]
    if not os.path.exists(EXE) or any(os.path.getmtime(d) > os.path.getmtime(EXE) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-Werror", "-pthread", "-o", EXE, SRC, "-L" + libdir,
                               "-lfhe_b200", "-Wl,-rpath," + libdir])
    return EXE


def wire_save(f, ct, size, k, n):
    """seal::Ciphertext::save record (host/seal_wire_b200.hpp): hash, 4 x int32, (n+1)-padded limbs."""
    f.write(struct.pack("<4Q", *HASH))
    f.write(struct.pack("<4i", size, size, n + 1, k))
    rows = np.zeros((size * k, n + 1), dtype="<u8")
    rows[:, :n] = ct.reshape(size * k, n)
    f.write(rows.tobytes())


def wire_load_all(path, k, n):
    out = []
    with open(path, "rb") as f:
        while True:
            head = f.read(32)
            if not head:
                break
            assert struct.unpack("<4Q", head) == HASH  # the server echoes its client's hash block
            cap, size, pcc, cmc = struct.unpack("<4i", f.read(16))
            assert (cap, pcc, cmc) == (size, n + 1, k)
            rows = np.frombuffer(f.read(size * k * (n + 1) * 8), dtype="<u8").reshape(size * k, n + 1)
            assert (rows[:, n] == 0).all()
            out.append(np.ascontiguousarray(rows[:, :n]).reshape(-1).astype(np.uint64))
    return out


def test_server_builds_parses_flags_and_has_no_cpu_fallback(pkg, tmp_path):
    exe = build_exe(pkg)
    r = subprocess.run([exe, "--help"], capture_output=True, text=True)
    assert r.returncode == 0 and "--cmod" in r.stdout and "--outfile" in r.stdout
    r = subprocess.run([exe, "--nonsense"], capture_output=True, text=True)
    assert r.returncode == 1 and "error parsing options" in r.stdout
    r = subprocess.run([exe], capture_output=True, text=True, cwd=tmp_path)     # no ./keys/params.txt
    assert r.returncode == 1 and "params.txt" in r.stderr
    import torch
    if torch.cuda.is_available():
        return
    os.makedirs(tmp_path / "keys")
    (tmp_path / "keys" / "params.txt").write_text("8 8 3 16384\n")
    r = subprocess.run([exe, "--cmod", "4096"], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 1 and "no usable CUDA device" in r.stderr


def plain_tile(rgb):
    """The plaintext pipeline of the reference for one tile: rgb -> ycc (fhe_image.h:317-319), then
    dct() (fhe_image.h:400-484) per channel."""
    from test_oracle import plain_dct
    r, g, b = (rgb[:, c].astype(float) for c in range(3))
    y = 0.299 * r + 0.587 * g + 0.114 * b - 128.0
    u = -0.168736 * r - 0.331264 * g + 0.5 * b
    v = 0.5 * r - 0.418688 * g - 0.081312 * b
    return [plain_dct(list(ch)) for ch in (y, u, v)]


@pytest.mark.gpu
def test_server_jpeg_end_to_end_on_gpu(pkg, oracle_mod, tmp_path):
    exe = build_exe(pkg)
    n, width, height = 4096, 16, 8          # two 8x8 RGB tiles
    q = oracle_mod.default_coeff_modulus(n)
    o = oracle_mod.Oracle(n, q, 1 << 14)
    k, W = o.k, o.ct_words(2)
    sk, pk = o.keygen(2026)
    rng = np.random.default_rng(5)
    image = rng.integers(0, 256, size=(height, width, 3))
    os.makedirs(tmp_path / "keys")
    os.makedirs(tmp_path / "image")
    (tmp_path / "keys" / "params.txt").write_text(f"{width} {height} 3 {1 << 14}\n")
    # client_jpeg --send: split into 8x8 blocks (fhe_image.h:108-124), per tile R64, G64, B64
    tiles, stream = [], []
    for by in range(height // 8):
        for bx in range(width // 8):
            blk = image[8 * by:8 * by + 8, 8 * bx:8 * bx + 8].reshape(64, 3)
            tiles.append(blk)
            for c in range(3):
                for p in range(64):
                    stream.append(o.encrypt(pk, o.encode(float(blk[p, c])), len(stream) + 1))
    with open(tmp_path / "image" / "nothingpersonnel.txt", "wb") as f:
        for ct in stream:
            wire_save(f, ct, 2, k, n)
    r = subprocess.run([exe, "--cmod", str(n), "--pmod", str(1 << 14), "--chunk", "1"], capture_output=True, text=True,
                       cwd=tmp_path, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = r.stdout.strip().splitlines()
    assert len(lines) == 2 * len(tiles)
    assert lines[0].startswith("RGBYCC,") and len(lines[0].rstrip(",").split(",")) == 65
    assert lines[1].startswith("DCT,") and len(lines[1].rstrip(",").split(",")) == 4
    got = wire_load_all(tmp_path / "image" / "zoop.txt", k, n)
    assert len(got) == len(stream)
    # oracle: the reference's loop op by op
    for t, blk in enumerate(tiles):
        chans = [np.concatenate(stream[(t * 3 + c) * 64:(t * 3 + c + 1) * 64]) for c in range(3)]
        for p in range(64):
            r_, g_, b_ = (np.ascontiguousarray(ch[p * W:(p + 1) * W]) for ch in chans)
            o.rgb_to_ycc_fhe(r_, g_, b_)
            for ch, x in zip(chans, (r_, g_, b_)):
                ch[p * W:(p + 1) * W] = x
        plain = plain_tile(blk)
        for c in range(3):
            o.encrypted_dct(chans[c])
            for p in range(64):
                out = got[(t * 3 + c) * 64 + p]
                assert (out == chans[c][p * W:(p + 1) * W]).all(), (t, c, p)
            for p in (0, 1, 9, 63):
                val = o.decode(o.decrypt(sk, got[(t * 3 + c) * 64 + p], 2))
                assert abs(val - plain[c][p]) < 1e-3, (t, c, p, val, plain[c][p])
    # client_jpeg --recieve (client_jpeg.cpp:255-291): decrypt + decode every ciphertext of the
    # stream, round, entropy-code with QUALITY = 0 header tables; the file must decode (PIL) to the
    # picture that was encrypted
    from test_client_jpeg import decode_rgb, pack
    values = np.array([o.decode(o.decrypt(sk, ct, 2)) for ct in got]).reshape(len(tiles), 3, 64)
    rgb = decode_rgb(pack(values, width, height, 0, str(tmp_path / "fhe.jpg")))
    rms = float(np.sqrt(np.mean((rgb.astype(float) - image.astype(float)) ** 2)))
    assert rms < 1.5, rms


@pytest.mark.gpu
def test_wire_header_variant_without_size_capacity(pkg, oracle_mod, tmp_path):
    """the record layout is recalled, not checked against SEAL; the field most likely to differ is
    size_capacity.  `--wire nocapacity` reads and writes records without it: same ciphertexts, 4 bytes less each."""
    exe = build_exe(pkg)
    n, width, height = 2048, 8, 8           # one tile
    q = oracle_mod.default_coeff_modulus(n)
    o = oracle_mod.Oracle(n, q, 1 << 14)
    k = o.k
    rng = np.random.default_rng(11)
    cts = [rng.integers(0, min(q), size=2 * k * n, dtype=np.uint64) for _ in range(192)]
    outs = {}
    for variant in ("capacity", "nocapacity"):
        d = tmp_path / variant
        os.makedirs(d / "keys")
        os.makedirs(d / "image")
        (d / "keys" / "params.txt").write_text(f"{width} {height} 3 {1 << 14}\n")
        with open(d / "image" / "nothingpersonnel.txt", "wb") as f:
            for ct in cts:
                f.write(struct.pack("<4Q", *HASH))
                f.write(struct.pack("<4i", 2, 2, n + 1, k) if variant == "capacity" else struct.pack("<3i", 2, n + 1, k))
                rows = np.zeros((2 * k, n + 1), dtype="<u8")
                rows[:, :n] = ct.reshape(2 * k, n)
                f.write(rows.tobytes())
        r = subprocess.run([exe, "--cmod", str(n), "--pmod", str(1 << 14), "--wire", variant], capture_output=True, text=True,
                           cwd=d, timeout=600)
        assert r.returncode == 0, r.stdout + r.stderr
        raw = (d / "image" / "zoop.txt").read_bytes()
        head = 32 + (16 if variant == "capacity" else 12)
        rec = head + 2 * k * (n + 1) * 8
        assert len(raw) == 192 * rec
        recs = []
        for i in range(192):
            b = raw[i * rec:(i + 1) * rec]
            assert struct.unpack("<4Q", b[:32]) == HASH
            meta = struct.unpack("<4i" if variant == "capacity" else "<3i", b[32:head])
            assert meta == ((2, 2, n + 1, k) if variant == "capacity" else (2, n + 1, k))
            recs.append(np.frombuffer(b[head:], dtype="<u8").reshape(2 * k, n + 1))
        outs[variant] = recs
    for a, b in zip(outs["capacity"], outs["nocapacity"]):
        assert (a == b).all()
    # a stream of the other variant is refused, not mis-parsed
    r = subprocess.run([exe, "--cmod", str(n), "--pmod", str(1 << 14), "--wire", "nocapacity"], capture_output=True, text=True,
                       cwd=tmp_path / "capacity", timeout=600)
    assert r.returncode != 0


@pytest.mark.gpu
def test_published_rms_error_through_the_encrypted_pipeline(pkg, oracle_mod, tmp_path):
    """image/boazbarak.jpg (48 x 48, committed fixture) -> client-side encryption of every pixel
    (oracle harness) -> server_jpeg_b200 on the GPU -> decryption -> this repo's JPEG writer ->
    the reference's decoder and comparator: RMSError 1.71767, the number the reference publishes
    for its own pipeline (benchmark/results.txt:37-42)."""
    from test_client_jpeg import boazbarak, check_published_rms
    exe = build_exe(pkg)
    image, _, _ = boazbarak()
    height, width, _ = image.shape
    n = 4096
    q = oracle_mod.default_coeff_modulus(n)
    o = oracle_mod.Oracle(n, q, 1 << 14)
    k = o.k
    sk, pk = o.keygen(48)
    os.makedirs(tmp_path / "keys")
    os.makedirs(tmp_path / "image")
    (tmp_path / "keys" / "params.txt").write_text(f"{width} {height} 3 {1 << 14}\n")
    seed = 1
    with open(tmp_path / "image" / "nothingpersonnel.txt", "wb") as f:
        for by in range(height // 8):
            for bx in range(width // 8):
                blk = image[8 * by:8 * by + 8, 8 * bx:8 * bx + 8].reshape(64, 3)
                for c in range(3):
                    for p in range(64):
                        wire_save(f, o.encrypt(pk, o.encode(float(blk[p, c])), seed), 2, k, n)
                        seed += 1
    r = subprocess.run([exe, "--cmod", str(n), "--pmod", str(1 << 14)], capture_output=True, text=True,
                       cwd=tmp_path, timeout=1800)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    got = wire_load_all(tmp_path / "image" / "zoop.txt", k, n)
    tiles = (height // 8) * (width // 8)
    assert len(got) == tiles * 192
    values = np.array([o.decode(o.decrypt(sk, ct, 2)) for ct in got]).reshape(tiles, 3, 64)
    check_published_rms(values, tmp_path)
